/*
 * zb_oracle.c -- CPU ORACLE for the zebra query hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is a plain-C restatement of the reference algorithm (emmyoh/zebra, read-only copy at
 * /root/reference).  It exists to CHECK the CUDA path and to provide the reported CPU baseline.  It is
 * never linked into, imported by, or called from the product (zebra_b200/): only tests/, bench.py's
 * cpu_baseline / --impl reference legs and __graft_entry__.smoke() may use it.
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures, cannot be built here (no Rust
 * toolchain), and its arithmetic lives in the un-vendored crate simsimd (^6.2.3, Cargo.toml:27).  The
 * simsimd kernels are restated from their published algorithm (AVX-512 "skylake" f32 kernels); the
 * accumulation order chosen here ("skylake-16", see below) is the contract the CUDA kernels reproduce
 * bit-for-bit.  The known-answer tests in tests/ are hand-derived from the reference's source.
 *
 * Reference lines followed (all under /root/reference/src):
 *   point_is_above        database/index/lsh.rs:39-43
 *   tree_result           database/index/lsh.rs:290-348   (quirks Q1 count cascade, Q2 leaf truncation)
 *   search                database/index/lsh.rs:544-565
 *   insert (descent)      database/index/lsh.rs:350-382
 *   build_a_tree          database/index/lsh.rs:250-267
 *   build_hyperplane      database/index/lsh.rs:192-248   (plane arithmetic :222-225, :174-190)
 *   remove                database/index/lsh.rs:473-503   (divergence D1: tombstone = intent of :488-490)
 *   Cosine/L2Squared/L2   distance.rs:19-32, :38-49, :103-114
 *
 * Deliberate, documented divergences:
 *   D1  remove() filters the id out of every leaf (the evident intent) instead of only root leaves.
 *   D2  hyperplane sample pairs come from a seeded min-hash over the node's own members instead of
 *       rand::rng() over the whole partition (lsh.rs:197-201 is unseeded and nondeterministic).
 *   D3  ties in distance are broken by id ascending (reference sorts are unstable, lsh.rs:318,:561).
 *   D4  add() on an existing forest is a batch: append to leaves, then rebuild every leaf whose live
 *       length exceeds max_node_size (reference splits one vector at a time, racily, lsh.rs:445-462).
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#define ZBO_MAX_DEPTH 96
#define ZBO_MAX_ATTEMPTS 4

/* Metric codes.  0..2 run on the simsimd path (f64 bits); 3..11 are the scalar metrics of the `distances` crate
 * (distance.rs:51-190, f32 bits zero-extended, survey quirk Q6).  Minkowski / p-norm carry their power in bits 8..:
 * metric = code | (power << 8). */
enum {
    ZBO_COSINE = 0, ZBO_L2SQ = 1, ZBO_L2 = 2,
    ZBO_CHEBYSHEV = 3, ZBO_CANBERRA = 4, ZBO_BRAY_CURTIS = 5, ZBO_MANHATTAN = 6, ZBO_L3 = 7, ZBO_L4 = 8,
    ZBO_HAMMING = 9, ZBO_MINKOWSKI = 10, ZBO_PNORM = 11
};

/* ------------------------------------------------------------------------------------------------
 * Arithmetic: the "skylake-16" order.
 * simsimd's AVX-512 f32 kernels keep one 16-lane f32 accumulator: lane j receives elements
 * j, j+16, j+32, ... through one fused multiply-add each (a zero-filled masked load covers a tail,
 * which leaves the remaining lanes unchanged), then reduce horizontally in f32:
 *   x[i] = acc[i] + acc[i+8]   (i<8)      512 -> 256   (_mm512_shuffle_f32x4 (0,0,3,2))
 *   r[i] = x[i]   + x[i+4]     (i<4)      256 -> 128   (_mm512_shuffle_f32x4 (0,0,0,1))
 *   s    = (r0 + r1) + (r2 + r3)          two _mm_hadd_ps
 * and the f32 result is widened to f64 (simsimd_distance_t).
 * ---------------------------------------------------------------------------------------------- */
static inline float reduce16(const float acc[16]) {
    float x[8], r[4];
    for (int i = 0; i < 8; ++i) x[i] = acc[i] + acc[i + 8];
    for (int i = 0; i < 4; ++i) r[i] = x[i] + x[i + 4];
    float h0 = r[0] + r[1];
    float h1 = r[2] + r[3];
    return h0 + h1;
}

static float dot_scalar(const float* a, const float* b, int n) {
    float acc[16];
    for (int j = 0; j < 16; ++j) acc[j] = 0.0f;
    int c = 0;
    for (; c + 16 <= n; c += 16)
        for (int j = 0; j < 16; ++j) acc[j] = fmaf(a[c + j], b[c + j], acc[j]);
    for (int j = 0; c + j < n; ++j) acc[j] = fmaf(a[c + j], b[c + j], acc[j]);
    return reduce16(acc);
}

static float l2sq_scalar(const float* a, const float* b, int n) {
    float acc[16];
    for (int j = 0; j < 16; ++j) acc[j] = 0.0f;
    int c = 0;
    for (; c + 16 <= n; c += 16)
        for (int j = 0; j < 16; ++j) {
            float d = a[c + j] - b[c + j];
            acc[j] = fmaf(d, d, acc[j]);
        }
    for (int j = 0; c + j < n; ++j) {
        float d = a[c + j] - b[c + j];
        acc[j] = fmaf(d, d, acc[j]);
    }
    return reduce16(acc);
}

static void cos3_scalar(const float* a, const float* b, int n, float* ab, float* a2, float* b2) {
    float xab[16], xa2[16], xb2[16];
    for (int j = 0; j < 16; ++j) xab[j] = xa2[j] = xb2[j] = 0.0f;
    int c = 0;
    for (; c + 16 <= n; c += 16)
        for (int j = 0; j < 16; ++j) {
            xab[j] = fmaf(a[c + j], b[c + j], xab[j]);
            xa2[j] = fmaf(a[c + j], a[c + j], xa2[j]);
            xb2[j] = fmaf(b[c + j], b[c + j], xb2[j]);
        }
    for (int j = 0; c + j < n; ++j) {
        xab[j] = fmaf(a[c + j], b[c + j], xab[j]);
        xa2[j] = fmaf(a[c + j], a[c + j], xa2[j]);
        xb2[j] = fmaf(b[c + j], b[c + j], xb2[j]);
    }
    *ab = reduce16(xab);
    *a2 = reduce16(xa2);
    *b2 = reduce16(xb2);
}

#if defined(__x86_64__)
/* Same order with AVX-512 intrinsics (used for the CPU baseline; bit-identical to the scalar form,
 * asserted in tests/test_oracle_arith.py). */
__attribute__((target("avx512f,fma,sse3"))) static inline float reduce16_avx512(__m512 a) {
    __m512 x = _mm512_add_ps(a, _mm512_shuffle_f32x4(a, a, _MM_SHUFFLE(0, 0, 3, 2)));
    __m128 r = _mm512_castps512_ps128(_mm512_add_ps(x, _mm512_shuffle_f32x4(x, x, _MM_SHUFFLE(0, 0, 0, 1))));
    r = _mm_hadd_ps(r, r);
    return _mm_cvtss_f32(_mm_hadd_ps(r, r));
}
__attribute__((target("avx512f,fma,sse3,bmi2"))) static float dot_avx512(const float* a, const float* b, int n) {
    __m512 acc = _mm512_setzero_ps();
    int c = 0;
    for (; c + 16 <= n; c += 16) acc = _mm512_fmadd_ps(_mm512_loadu_ps(a + c), _mm512_loadu_ps(b + c), acc);
    if (c < n) {
        __mmask16 m = (__mmask16)((1u << (n - c)) - 1u);
        acc = _mm512_fmadd_ps(_mm512_maskz_loadu_ps(m, a + c), _mm512_maskz_loadu_ps(m, b + c), acc);
    }
    return reduce16_avx512(acc);
}
__attribute__((target("avx512f,fma,sse3,bmi2"))) static float l2sq_avx512(const float* a, const float* b, int n) {
    __m512 acc = _mm512_setzero_ps();
    int c = 0;
    for (; c + 16 <= n; c += 16) {
        __m512 d = _mm512_sub_ps(_mm512_loadu_ps(a + c), _mm512_loadu_ps(b + c));
        acc = _mm512_fmadd_ps(d, d, acc);
    }
    if (c < n) {
        __mmask16 m = (__mmask16)((1u << (n - c)) - 1u);
        __m512 d = _mm512_sub_ps(_mm512_maskz_loadu_ps(m, a + c), _mm512_maskz_loadu_ps(m, b + c));
        acc = _mm512_fmadd_ps(d, d, acc);
    }
    return reduce16_avx512(acc);
}
__attribute__((target("avx512f,fma,sse3,bmi2"))) static void cos3_avx512(const float* a, const float* b, int n,
                                                                        float* ab, float* a2, float* b2) {
    __m512 vab = _mm512_setzero_ps(), va2 = _mm512_setzero_ps(), vb2 = _mm512_setzero_ps();
    int c = 0;
    for (; c + 16 <= n; c += 16) {
        __m512 va = _mm512_loadu_ps(a + c), vb = _mm512_loadu_ps(b + c);
        vab = _mm512_fmadd_ps(va, vb, vab);
        va2 = _mm512_fmadd_ps(va, va, va2);
        vb2 = _mm512_fmadd_ps(vb, vb, vb2);
    }
    if (c < n) {
        __mmask16 m = (__mmask16)((1u << (n - c)) - 1u);
        __m512 va = _mm512_maskz_loadu_ps(m, a + c), vb = _mm512_maskz_loadu_ps(m, b + c);
        vab = _mm512_fmadd_ps(va, vb, vab);
        va2 = _mm512_fmadd_ps(va, va, va2);
        vb2 = _mm512_fmadd_ps(vb, vb, vb2);
    }
    *ab = reduce16_avx512(vab);
    *a2 = reduce16_avx512(va2);
    *b2 = reduce16_avx512(vb2);
}
#endif

static int g_use_avx512 = -1; /* -1 = probe, 0 = scalar, 1 = avx512 */
static int use_avx512(void) {
    if (g_use_avx512 < 0) {
#if defined(__x86_64__)
        __builtin_cpu_init();
        g_use_avx512 = __builtin_cpu_supports("avx512f") ? 1 : 0;
#else
        g_use_avx512 = 0;
#endif
    }
    return g_use_avx512;
}
void zbo_force_scalar(int on) { g_use_avx512 = on ? 0 : -1; }
int zbo_using_avx512(void) { return use_avx512(); }

/* simsimd_dot_f32 (lsh.rs:40, :224): f32 result widened to f64. */
double zbo_dot_f32(const float* a, const float* b, int n) {
#if defined(__x86_64__)
    if (use_avx512()) return (double)dot_avx512(a, b, n);
#endif
    return (double)dot_scalar(a, b, n);
}
/* simsimd_l2sq_f32 (distance.rs:41). */
double zbo_l2sq_f32(const float* a, const float* b, int n) {
#if defined(__x86_64__)
    if (use_avx512()) return (double)l2sq_avx512(a, b, n);
#endif
    return (double)l2sq_scalar(a, b, n);
}
/* simsimd_l2_f32 (distance.rs:106): sqrt of the squared distance, in f64. */
double zbo_l2_f32(const float* a, const float* b, int n) { return sqrt(zbo_l2sq_f32(a, b, n)); }
void zbo_cos3_f32(const float* a, const float* b, int n, float* ab, float* a2, float* b2) {
#if defined(__x86_64__)
    if (use_avx512()) {
        cos3_avx512(a, b, n, ab, a2, b2);
        return;
    }
#endif
    cos3_scalar(a, b, n, ab, a2, b2);
}
/* simsimd_cos_f32 (distance.rs:23): cosine DISTANCE = 1 - similarity, clipped at 0, normalised in f64.
 * simsimd uses rsqrt14 + one Newton-Raphson step (rel. error ~1e-8); the oracle uses the exactly rounded
 * 1/sqrt, which is what the 1e-5 relative tolerance of north_star absorbs. */
double zbo_cosdist_from3(float ab_, float a2_, float b2_) {
    double ab = (double)ab_, a2 = (double)a2_, b2 = (double)b2_;
    if (a2 == 0.0 && b2 == 0.0) return 0.0;
    if (ab == 0.0) return 1.0;
    double ra = 1.0 / sqrt(a2);
    double rb = 1.0 / sqrt(b2);
    double t = ab * ra;
    t = t * rb;
    double r = 1.0 - t;
    return r > 0.0 ? r : 0.0;
}
double zbo_cos_f32(const float* a, const float* b, int n) {
    float ab, a2, b2;
    zbo_cos3_f32(a, b, n, &ab, &a2, &b2);
    return zbo_cosdist_from3(ab, a2, b2);
}

static inline uint64_t f64_bits(double x) {
    uint64_t u;
    memcpy(&u, &x, 8);
    return u;
}

/* ------------------------------------------------------------------------------------------------
 * The ten scalar metrics (distance.rs:51-190).  They call the un-vendored crate `distances` (^1.8.0,
 * Cargo.toml:38; source NOT on this machine -- restated from its published algorithm, PARITY UNPINNED):
 * plain iterator folds over the zipped slices in input order, in f32 (T = U = EmbeddingPrecision), no
 * reassociation (rustc never contracts or reorders float arithmetic), result `.to_bits()` (u32) widened
 * to u64.  Spec the oracle fixes (oracle/README.md, "scalar metrics"):
 *   abs_diff(a,b) = |a - b|
 *   chebyshev    fold(0, |acc, v| if acc > v { acc } else { v })          distance.rs:57-60
 *   canberra     sum |a-b| / (|a| + |b|)   (0/0 = NaN, as IEEE gives)      distance.rs:69-72
 *   bray_curtis  (sum |a-b|) / (sum |a+b|)                                 distance.rs:81-84
 *   manhattan    sum |a-b|                                                 distance.rs:93-96
 *   l3_norm      cbrt(sum v*v*v),   l4_norm  sqrt(sqrt(sum (v*v)*(v*v)))   distance.rs:122-125, :134-137
 *   hamming      popcount over the LOW BYTE of every element's bit pattern distance.rs:145-155
 *   minkowski(p) powf(sum powi(v, p), 1/p);  minkowski_p(p) = the sum      distance.rs:168-172, :185-189
 * powi is compiler-rt's __powisf2 (square and multiply).  The two roots that are not IEEE operations
 * (cbrt, powf(., 1/p)) are replaced by ONE deterministic algorithm built from IEEE double operations only
 * (root_p below), so the CPU and the GPU agree bit for bit; against libm's cbrtf / powf it is within 1 ulp
 * of f32 (powf(s, fl32(1/p)) additionally differs from the true p-th root by about |ln s| * 1e-8 relative).
 * A NaN result is canonicalised to 0xFFC00000, the default NaN of x86 SSE arithmetic (what 0.0/0.0 yields on
 * the reference's host).
 * ---------------------------------------------------------------------------------------------- */
static inline uint64_t f32_key(float r) {
    uint32_t u;
    if (r != r) return 0xFFC00000ull;
    memcpy(&u, &r, 4);
    return (uint64_t)u;
}
static inline float powi_f32(float a, int b) { /* compiler-rt __powisf2, b >= 0 */
    float r = 1.0f;
    for (;;) {
        if (b & 1) r *= a;
        b /= 2;
        if (b == 0) break;
        a *= a;
    }
    return r;
}
static inline double powi_f64(double a, int b) {
    double r = 1.0;
    for (;;) {
        if (b & 1) r *= a;
        b /= 2;
        if (b == 0) break;
        a *= a;
    }
    return r;
}
/* p-th root of a non-negative f32 (p >= 1): Newton's iteration y <- ((p-1) y + x / y^(p-1)) / p in f64 from a
 * bit-pattern initial guess, stopped at the first step that does not decrease y, rounded once to f32. */
float zbo_root_p(float s, int p) {
    if (p == 1 || s != s || s == 0.0f || s == INFINITY) return s;
    double x = (double)s;
    uint64_t bx, by;
    memcpy(&bx, &x, 8);
    by = bx / (uint64_t)p + (0x3FF0000000000000ull / (uint64_t)p) * (uint64_t)(p - 1);
    double y;
    memcpy(&y, &by, 8);
    for (int it = 0; it < 1000; ++it) {
        double t = x / powi_f64(y, p - 1);
        double yn = ((double)(p - 1) * y + t) / (double)p;
        if (it > 0 && yn >= y) break;
        y = yn;
    }
    return (float)y;
}
static uint64_t scalar_metric_bits(int code, int power, const float* a, const float* b, int n) {
    float acc = 0.0f, den = 0.0f;
    uint32_t ham = 0;
    for (int i = 0; i < n; ++i) {
        float x = a[i], y = b[i];
        float v = fabsf(x - y);
        switch (code) {
            case ZBO_CHEBYSHEV: acc = acc > v ? acc : v; break;
            case ZBO_CANBERRA: acc = acc + v / (fabsf(x) + fabsf(y)); break;
            case ZBO_BRAY_CURTIS: acc = acc + v; den = den + fabsf(x + y); break;
            case ZBO_MANHATTAN: acc = acc + v; break;
            case ZBO_L3: acc = acc + (v * v) * v; break;
            case ZBO_L4: { float v2 = v * v; acc = acc + v2 * v2; break; }
            case ZBO_HAMMING: {
                uint32_t ux, uy;
                memcpy(&ux, &x, 4);
                memcpy(&uy, &y, 4);
                ham += (uint32_t)__builtin_popcount((ux ^ uy) & 0xFFu);
                break;
            }
            default: acc = acc + powi_f32(v, power); break; /* minkowski, p-norm */
        }
    }
    switch (code) {
        case ZBO_BRAY_CURTIS: return f32_key(acc / den);
        case ZBO_L3: return f32_key(zbo_root_p(acc, 3));
        case ZBO_L4: return f32_key(sqrtf(sqrtf(acc)));
        case ZBO_HAMMING: return (uint64_t)ham;
        case ZBO_MINKOWSKI:
            if (power == 0) return f32_key(acc == 1.0f ? 1.0f : INFINITY); /* powf(n, 1/0 = +inf), n >= 1 */
            return f32_key(zbo_root_p(acc, power));
        default: return f32_key(acc);
    }
}

/* Metric::distance -> DistanceUnit = u64 of the f64 bits (distance.rs:13, :19-32, :38-49, :103-114).
 * Argument order is (stored row, query) as at lsh.rs:314 and :559.  CosineDistance returns
 * (1.0 - simsimd cosine distance).to_bits() -- quirk Q4, reproduced literally. */
uint64_t zbo_distance_bits(int metric, const float* row, const float* query, int n) {
    switch (metric & 0xFF) {
        case ZBO_COSINE: return f64_bits(1.0 - zbo_cos_f32(row, query, n));
        case ZBO_L2SQ: return f64_bits(zbo_l2sq_f32(row, query, n));
        case ZBO_L2: return f64_bits(zbo_l2_f32(row, query, n));
        default: return scalar_metric_bits(metric & 0xFF, metric >> 8, row, query, n);
    }
}

/* Hyperplane::point_is_above, lsh.rs:39-43. */
int zbo_point_is_above(const float* coef, float constant, const float* x, int n) {
    return zbo_dot_f32(coef, x, n) + (double)constant >= 0.0;
}

/* Plane through the midpoint of a and b, normal b - a: lsh.rs:222-225 with subtract/average :174-190. */
void zbo_make_plane(const float* a, const float* b, int n, float* coef, float* constant) {
    float* mid = (float*)malloc(sizeof(float) * (size_t)n);
    for (int i = 0; i < n; ++i) {
        coef[i] = b[i] - a[i];
        mid[i] = (a[i] + b[i]) / 2.0f;
    }
    *constant = -(float)zbo_dot_f32(coef, mid, n);
    free(mid);
}

/* ------------------------------------------------------------------------------------------------
 * Deterministic sampling spec shared (by specification, not by code) with the CUDA build (D2).
 * ---------------------------------------------------------------------------------------------- */
static inline uint64_t mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
uint64_t zbo_mix64(uint64_t x) { return mix64(x); }
static inline uint64_t root_key(uint64_t seed, int tree) { return mix64(seed ^ mix64((uint64_t)tree)); }
static inline uint64_t child_key(uint64_t key, int side) {
    return mix64(key ^ (side ? 0xA5A5A5A5A5A5A5A5ull : 0x5A5A5A5A5A5A5A5Aull));
}
static inline uint64_t pick_hash(uint64_t key, int attempt, uint64_t ordinal) {
    return mix64(key ^ mix64(ordinal ^ ((uint64_t)attempt << 56)));
}

/* ------------------------------------------------------------------------------------------------
 * Forest: pointer-shaped like the reference's Node enum (lsh.rs:46-60).
 * ---------------------------------------------------------------------------------------------- */
typedef struct zbo_node {
    int is_leaf;
    int depth;
    uint64_t key;
    /* inner */
    float* coef;
    float constant;
    struct zbo_node* left;  /* below */
    struct zbo_node* right; /* above */
    /* leaf */
    uint64_t* ids;
    size_t len, cap;
    int export_id;
} zbo_node;

typedef struct zbo_index {
    int dim;
    int metric;
    size_t max_node_size;
    int num_trees;
    uint64_t seed;
    float* rows;
    size_t n_rows, cap_rows;
    uint8_t* tomb;
    size_t n_live;
    zbo_node** roots;
    int built;
} zbo_index;

static zbo_node* new_leaf(const uint64_t* ids, size_t len, uint64_t key, int depth) {
    zbo_node* nd = (zbo_node*)calloc(1, sizeof(zbo_node));
    nd->is_leaf = 1;
    nd->key = key;
    nd->depth = depth;
    nd->cap = len > 4 ? len : 4;
    nd->ids = (uint64_t*)malloc(sizeof(uint64_t) * nd->cap);
    if (len) memcpy(nd->ids, ids, sizeof(uint64_t) * len);
    nd->len = len;
    nd->export_id = -1;
    return nd;
}
static void free_node(zbo_node* nd) {
    if (!nd) return;
    if (nd->is_leaf) {
        free(nd->ids);
    } else {
        free(nd->coef);
        free_node(nd->left);
        free_node(nd->right);
    }
    free(nd);
}
static inline const float* row_of(const zbo_index* ix, uint64_t id) { return ix->rows + (size_t)id * (size_t)ix->dim; }

zbo_index* zbo_create(int dim, int metric, uint64_t max_node_size, int num_trees, uint64_t seed) {
    zbo_index* ix = (zbo_index*)calloc(1, sizeof(zbo_index));
    ix->dim = dim;
    ix->metric = metric;
    ix->max_node_size = (size_t)max_node_size;
    ix->num_trees = num_trees;
    ix->seed = seed;
    ix->roots = (zbo_node**)calloc((size_t)(num_trees > 0 ? num_trees : 1), sizeof(zbo_node*));
    return ix;
}
void zbo_destroy(zbo_index* ix) {
    if (!ix) return;
    for (int t = 0; t < ix->num_trees; ++t) free_node(ix->roots[t]);
    free(ix->roots);
    free(ix->rows);
    free(ix->tomb);
    free(ix);
}
uint64_t zbo_num_rows(const zbo_index* ix) { return ix->n_rows; }
uint64_t zbo_num_live(const zbo_index* ix) { return ix->n_live; }

static void append_rows(zbo_index* ix, size_t n, const float* rows) {
    if (ix->n_rows + n > ix->cap_rows) {
        size_t nc = ix->cap_rows ? ix->cap_rows * 2 : 1024;
        while (nc < ix->n_rows + n) nc *= 2;
        ix->rows = (float*)realloc(ix->rows, nc * (size_t)ix->dim * sizeof(float));
        ix->tomb = (uint8_t*)realloc(ix->tomb, nc);
        ix->cap_rows = nc;
    }
    memcpy(ix->rows + ix->n_rows * (size_t)ix->dim, rows, n * (size_t)ix->dim * sizeof(float));
    memset(ix->tomb + ix->n_rows, 0, n);
    ix->n_rows += n;
    ix->n_live += n;
}

/* build_a_tree, lsh.rs:250-267, with build_hyperplane lsh.rs:192-248 (sampling per D2).
 * `ids` are live members in ascending id order; the partition is stable so children stay ascending. */
static zbo_node* build_tree(zbo_index* ix, const uint64_t* ids, size_t len, uint64_t key, int depth) {
    if (len < ix->max_node_size || len < 2 || depth >= ZBO_MAX_DEPTH) return new_leaf(ids, len, key, depth);
    int n = ix->dim;
    float* coef = (float*)malloc(sizeof(float) * (size_t)n);
    float constant = 0.0f;
    uint64_t* below = (uint64_t*)malloc(sizeof(uint64_t) * len);
    uint64_t* above = (uint64_t*)malloc(sizeof(uint64_t) * len);
    size_t nb = 0, na = 0;
    int ok = 0;
    for (int attempt = 0; attempt < ZBO_MAX_ATTEMPTS && !ok; ++attempt) {
        /* a = member with the smallest (hash, id); b = the next smallest. */
        uint64_t ha = ~0ull, hb = ~0ull, ia = ~0ull, ib = ~0ull;
        for (size_t i = 0; i < len; ++i) {
            uint64_t h = pick_hash(key, attempt, ids[i]);
            if (h < ha || (h == ha && ids[i] < ia)) {
                hb = ha; ib = ia;
                ha = h; ia = ids[i];
            } else if (h < hb || (h == hb && ids[i] < ib)) {
                hb = h; ib = ids[i];
            }
        }
        zbo_make_plane(row_of(ix, ia), row_of(ix, ib), n, coef, &constant);
        nb = na = 0;
        for (size_t i = 0; i < len; ++i) {
            if (zbo_point_is_above(coef, constant, row_of(ix, ids[i]), n)) above[na++] = ids[i];
            else below[nb++] = ids[i];
        }
        ok = (na > 0 && nb > 0);
    }
    zbo_node* nd;
    if (!ok) {
        nd = new_leaf(ids, len, key, depth);
        free(coef);
    } else {
        nd = (zbo_node*)calloc(1, sizeof(zbo_node));
        nd->is_leaf = 0;
        nd->key = key;
        nd->depth = depth;
        nd->coef = coef;
        nd->constant = constant;
        nd->export_id = -1;
        nd->left = build_tree(ix, below, nb, child_key(key, 0), depth + 1);
        nd->right = build_tree(ix, above, na, child_key(key, 1), depth + 1);
    }
    free(below);
    free(above);
    return nd;
}

static size_t live_members(const zbo_index* ix, const zbo_node* leaf, uint64_t* out) {
    size_t m = 0;
    for (size_t i = 0; i < leaf->len; ++i)
        if (!ix->tomb[leaf->ids[i]]) {
            if (out) out[m] = leaf->ids[i];
            ++m;
        }
    return m;
}

static void leaf_push(zbo_node* leaf, uint64_t id) {
    if (leaf->len == leaf->cap) {
        leaf->cap *= 2;
        leaf->ids = (uint64_t*)realloc(leaf->ids, sizeof(uint64_t) * leaf->cap);
    }
    leaf->ids[leaf->len++] = id;
}

/* insert descent, lsh.rs:350-366. */
static zbo_node* descend(const zbo_index* ix, zbo_node* nd, const float* x) {
    while (!nd->is_leaf) nd = zbo_point_is_above(nd->coef, nd->constant, x, ix->dim) ? nd->right : nd->left;
    return nd;
}

/* D4: rebuild every leaf whose live length exceeds max_node_size (lsh.rs:367-378 applied per batch). */
static zbo_node* split_overfull(zbo_index* ix, zbo_node* nd) {
    if (!nd->is_leaf) {
        nd->left = split_overfull(ix, nd->left);
        nd->right = split_overfull(ix, nd->right);
        return nd;
    }
    size_t live = live_members(ix, nd, NULL);
    if (live <= ix->max_node_size) return nd;
    uint64_t* ids = (uint64_t*)malloc(sizeof(uint64_t) * live);
    live_members(ix, nd, ids);
    zbo_node* rebuilt = build_tree(ix, ids, live, nd->key, nd->depth);
    free(ids);
    free_node(nd);
    return rebuilt;
}

/* (0..num_trees).into_par_iter() of build_index, lsh.rs:420-427: one task per tree. */
typedef struct { zbo_index* ix; const uint64_t* ids; size_t live; volatile int* next; } build_job;
static void* build_worker(void* arg) {
    build_job* job = (build_job*)arg;
    for (;;) {
        int t = __atomic_fetch_add(job->next, 1, __ATOMIC_RELAXED);
        if (t >= job->ix->num_trees) break;
        job->ix->roots[t] = build_tree(job->ix, job->ids, job->live, root_key(job->ix->seed, t), 0);
    }
    return NULL;
}
static int g_build_threads = 0;
void zbo_set_build_threads(int n) { g_build_threads = n; }
static void build_forest_parallel(zbo_index* ix, const uint64_t* ids, size_t live) {
    volatile int next = 0;
    build_job job = {ix, ids, live, &next};
    int nth = g_build_threads > 0 ? g_build_threads : 1;
    if (nth > ix->num_trees) nth = ix->num_trees;
    if (nth <= 1) {
        build_worker(&job);
        return;
    }
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nth);
    for (int i = 0; i < nth; ++i) pthread_create(&th[i], NULL, build_worker, &job);
    for (int i = 0; i < nth; ++i) pthread_join(th[i], NULL);
    free(th);
}

/* LSHIndex::add, lsh.rs:440-466 (build_index :411-429 when there are no trees yet).
 * ids are the row ordinals n_rows_before .. n_rows_before+n-1, written to out_ids when non-NULL. */
int zbo_add(zbo_index* ix, uint64_t n, const float* rows, uint64_t* out_ids) {
    size_t first = ix->n_rows;
    append_rows(ix, (size_t)n, rows);
    if (out_ids)
        for (size_t i = 0; i < n; ++i) out_ids[i] = first + i;
    if (!ix->built) {
        if (n == 0) return 0;
        size_t live = 0;
        uint64_t* ids = (uint64_t*)malloc(sizeof(uint64_t) * ix->n_rows);
        for (size_t i = 0; i < ix->n_rows; ++i)
            if (!ix->tomb[i]) ids[live++] = i;
        build_forest_parallel(ix, ids, live);
        free(ids);
        ix->built = 1;
        return 0;
    }
    for (int t = 0; t < ix->num_trees; ++t) {
        for (size_t i = 0; i < n; ++i) leaf_push(descend(ix, ix->roots[t], row_of(ix, first + i)), first + i);
        ix->roots[t] = split_overfull(ix, ix->roots[t]);
    }
    return 0;
}

/* LSHIndex::remove, lsh.rs:473-503 under D1.  out_removed[i] = 1 when ids[i] was live and is now removed. */
int zbo_remove(zbo_index* ix, uint64_t n, const uint64_t* ids, uint8_t* out_removed) {
    for (size_t i = 0; i < n; ++i) {
        int ok = ids[i] < ix->n_rows && !ix->tomb[ids[i]];
        if (ok) {
            ix->tomb[ids[i]] = 1;
            ix->n_live--;
        }
        if (out_removed) out_removed[i] = (uint8_t)ok;
    }
    return 0;
}

/* LSHIndex::deduplicate, lsh.rs:270-288: walk the stored embeddings in key (id) order, keep the first row of every
 * distinct BIT pattern (f32::to_bits per element: +0.0 and -0.0 differ, NaNs compare by payload), remove the rest.
 * Returns the number of removed rows; their ids (ascending) go to out_ids[0..min(count, cap)). */
static const zbo_index* g_dd_ix;
static int cmp_rowbits(const void* pa, const void* pb) {
    uint64_t a = *(const uint64_t*)pa, b = *(const uint64_t*)pb;
    int c = memcmp(g_dd_ix->rows + (size_t)a * g_dd_ix->dim, g_dd_ix->rows + (size_t)b * g_dd_ix->dim, (size_t)g_dd_ix->dim * 4);
    if (c) return c;
    return a < b ? -1 : (a > b ? 1 : 0);
}
int64_t zbo_deduplicate(zbo_index* ix, uint64_t* out_ids, uint64_t cap) {
    uint64_t* live = (uint64_t*)malloc((ix->n_live ? ix->n_live : 1) * sizeof(uint64_t));
    size_t n = 0;
    for (size_t i = 0; i < ix->n_rows; ++i)
        if (!ix->tomb[i]) live[n++] = i;
    g_dd_ix = ix;
    qsort(live, n, sizeof(uint64_t), cmp_rowbits);   /* equal rows adjacent, lowest id first */
    uint8_t* dup = (uint8_t*)calloc(ix->n_rows ? ix->n_rows : 1, 1);
    for (size_t i = 1; i < n; ++i)
        if (!memcmp(ix->rows + (size_t)live[i] * ix->dim, ix->rows + (size_t)live[i - 1] * ix->dim, (size_t)ix->dim * 4)) dup[live[i]] = 1;
    int64_t count = 0;
    for (size_t i = 0; i < ix->n_rows; ++i)
        if (dup[i]) {
            if ((uint64_t)count < cap && out_ids) out_ids[count] = i;
            ++count;
            ix->tomb[i] = 1;
            ix->n_live--;
        }
    free(dup);
    free(live);
    return count;
}

/* ------------------------------------------------------------------------------------------------
 * Search.
 * ---------------------------------------------------------------------------------------------- */
typedef struct { uint64_t bits, id; } scored;
static int cmp_scored(const void* pa, const void* pb) {
    const scored* a = (const scored*)pa;
    const scored* b = (const scored*)pb;
    if (a->bits != b->bits) return a->bits < b->bits ? -1 : 1; /* u64 compare of raw bits, lsh.rs:318,:561 */
    if (a->id != b->id) return a->id < b->id ? -1 : 1;          /* D3 */
    return 0;
}
typedef struct { uint64_t* v; size_t len, cap; } idvec;
static void idvec_push(idvec* s, uint64_t id) {
    if (s->len == s->cap) {
        s->cap = s->cap ? s->cap * 2 : 256;
        s->v = (uint64_t*)realloc(s->v, sizeof(uint64_t) * s->cap);
    }
    s->v[s->len++] = id;
}
typedef struct { int32_t tree, leaf, nprime, live; } zbo_visit;
typedef struct { zbo_visit* v; size_t len, cap; } visitvec;

/* tree_result, lsh.rs:290-348. */
static int tree_result(const zbo_index* ix, const float* q, int n, const zbo_node* nd, idvec* cand, visitvec* trace,
                       int tree) {
    if (nd->is_leaf) {
        size_t live = live_members(ix, nd, NULL); /* D1: tombstoned ids are not members */
        if (trace) {
            if (trace->len == trace->cap) {
                trace->cap = trace->cap ? trace->cap * 2 : 64;
                trace->v = (zbo_visit*)realloc(trace->v, sizeof(zbo_visit) * trace->cap);
            }
            zbo_visit vv = {tree, nd->export_id, n, (int32_t)live};
            trace->v[trace->len++] = vv;
        }
        if (n < 0 || live < (size_t)n) { /* lsh.rs:301-308 */
            for (size_t i = 0; i < nd->len; ++i)
                if (!ix->tomb[nd->ids[i]]) idvec_push(cand, nd->ids[i]);
            return (int)live;
        }
        /* lsh.rs:309-330: score every member, keep the n nearest */
        scored* sc = (scored*)malloc(sizeof(scored) * (live ? live : 1));
        size_t m = 0;
        for (size_t i = 0; i < nd->len; ++i) {
            uint64_t id = nd->ids[i];
            if (ix->tomb[id]) continue;
            sc[m].id = id;
            sc[m].bits = zbo_distance_bits(ix->metric, row_of(ix, id), q, ix->dim);
            ++m;
        }
        qsort(sc, m, sizeof(scored), cmp_scored);
        for (int i = 0; i < n; ++i) idvec_push(cand, sc[i].id);
        free(sc);
        return n;
    }
    int above = zbo_point_is_above(nd->coef, nd->constant, q, ix->dim); /* lsh.rs:334 */
    const zbo_node* main_ = above ? nd->right : nd->left;
    const zbo_node* backup = above ? nd->left : nd->right;
    int k = tree_result(ix, q, n, main_, cand, trace, tree);
    if (k < n) return tree_result(ix, q, n - k, backup, cand, trace, tree); /* Q1: k is dropped */
    return k;
}

static int cmp_u64(const void* a, const void* b) {
    uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}

/* LSHIndex::search, lsh.rs:544-565.  Returns the number of results (<= top_k). */
int64_t zbo_search(const zbo_index* ix, const float* q, uint64_t top_k, uint64_t* out_ids, uint64_t* out_bits) {
    idvec cand = {0, 0, 0};
    if (ix->built)
        for (int t = 0; t < ix->num_trees; ++t) tree_result(ix, q, (int)top_k, ix->roots[t], &cand, NULL, t);
    /* DashSet: dedup */
    qsort(cand.v, cand.len, sizeof(uint64_t), cmp_u64);
    size_t m = 0;
    for (size_t i = 0; i < cand.len; ++i)
        if (i == 0 || cand.v[i] != cand.v[i - 1]) cand.v[m++] = cand.v[i];
    scored* sc = (scored*)malloc(sizeof(scored) * (m ? m : 1));
    for (size_t i = 0; i < m; ++i) { /* lsh.rs:557-560: rescore the union */
        sc[i].id = cand.v[i];
        sc[i].bits = zbo_distance_bits(ix->metric, row_of(ix, cand.v[i]), q, ix->dim);
    }
    qsort(sc, m, sizeof(scored), cmp_scored);
    size_t r = m < top_k ? m : (size_t)top_k;
    for (size_t i = 0; i < r; ++i) {
        out_ids[i] = sc[i].id;
        out_bits[i] = sc[i].bits;
    }
    free(sc);
    free(cand.v);
    return (int64_t)r;
}

/* Database::query_vectors' parallel loop, core.rs:299: one task per query. */
typedef struct {
    const zbo_index* ix;
    const float* queries;
    uint64_t nq, top_k;
    uint64_t* out_ids;
    uint64_t* out_bits;
    uint32_t* out_counts;
    volatile int64_t* next;
} batch_job;
static void* batch_worker(void* arg) {
    batch_job* job = (batch_job*)arg;
    for (;;) {
        int64_t i = __atomic_fetch_add(job->next, 1, __ATOMIC_RELAXED);
        if ((uint64_t)i >= job->nq) break;
        int64_t r = zbo_search(job->ix, job->queries + (size_t)i * (size_t)job->ix->dim, job->top_k,
                               job->out_ids + (size_t)i * job->top_k, job->out_bits + (size_t)i * job->top_k);
        job->out_counts[i] = (uint32_t)r;
    }
    return NULL;
}
int zbo_search_batch(const zbo_index* ix, uint64_t nq, const float* queries, uint64_t top_k, int nthreads,
                     uint64_t* out_ids, uint64_t* out_bits, uint32_t* out_counts) {
    volatile int64_t next = 0;
    batch_job job = {ix, queries, nq, top_k, out_ids, out_bits, out_counts, &next};
    if (nthreads <= 1) {
        batch_worker(&job);
        return 0;
    }
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthreads);
    for (int i = 0; i < nthreads; ++i) pthread_create(&th[i], NULL, batch_worker, &job);
    for (int i = 0; i < nthreads; ++i) pthread_join(th[i], NULL);
    free(th);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Export / import of the forest as flat arrays (the layout zb_index_load_forest takes).
 * nodes[i] = {plane, left, right, leaf}: inner nodes have leaf = -1, leaves have plane = -1.
 * Numbering is preorder (node, left subtree, right subtree), trees in order.
 * ---------------------------------------------------------------------------------------------- */
typedef struct { int64_t nodes, planes, leaves, members; } zbo_sizes;
static void count_nodes(const zbo_node* nd, zbo_sizes* s) {
    s->nodes++;
    if (nd->is_leaf) {
        s->leaves++;
        s->members += (int64_t)nd->len;
    } else {
        s->planes++;
        count_nodes(nd->left, s);
        count_nodes(nd->right, s);
    }
}
void zbo_forest_sizes(const zbo_index* ix, int64_t* out4) {
    zbo_sizes s = {0, 0, 0, 0};
    if (ix->built)
        for (int t = 0; t < ix->num_trees; ++t) count_nodes(ix->roots[t], &s);
    out4[0] = s.nodes; out4[1] = s.planes; out4[2] = s.leaves; out4[3] = s.members;
}
typedef struct {
    int32_t* nodes; float* coef; float* cst; int64_t* leaf_off; uint64_t* members;
    int64_t n_nodes, n_planes, n_leaves, n_members; int dim;
} exporter;
static int32_t export_node(zbo_node* nd, exporter* e) {
    int32_t me = (int32_t)e->n_nodes++;
    if (nd->is_leaf) {
        int32_t lf = (int32_t)e->n_leaves++;
        nd->export_id = lf;
        e->leaf_off[lf] = e->n_members;
        memcpy(e->members + e->n_members, nd->ids, sizeof(uint64_t) * nd->len);
        e->n_members += (int64_t)nd->len;
        e->leaf_off[lf + 1] = e->n_members;
        e->nodes[4 * me + 0] = -1; e->nodes[4 * me + 1] = -1; e->nodes[4 * me + 2] = -1; e->nodes[4 * me + 3] = lf;
    } else {
        int32_t pl = (int32_t)e->n_planes++;
        nd->export_id = -1;
        memcpy(e->coef + (size_t)pl * (size_t)e->dim, nd->coef, sizeof(float) * (size_t)e->dim);
        e->cst[pl] = nd->constant;
        int32_t l = export_node(nd->left, e);
        int32_t r = export_node(nd->right, e);
        e->nodes[4 * me + 0] = pl; e->nodes[4 * me + 1] = l; e->nodes[4 * me + 2] = r; e->nodes[4 * me + 3] = -1;
    }
    return me;
}
int zbo_export_forest(zbo_index* ix, int32_t* nodes, int32_t* roots, float* coef, float* cst, int64_t* leaf_off,
                      uint64_t* members) {
    exporter e = {nodes, coef, cst, leaf_off, members, 0, 0, 0, 0, ix->dim};
    if (!ix->built) return 0;
    leaf_off[0] = 0;
    for (int t = 0; t < ix->num_trees; ++t) roots[t] = export_node(ix->roots[t], &e);
    return 0;
}
static zbo_node* import_node(const zbo_index* ix, int32_t i, const int32_t* nodes, const float* coef, const float* cst,
                             const int64_t* leaf_off, const uint64_t* members, uint64_t key, int depth) {
    const int32_t* nd = nodes + 4 * (size_t)i;
    if (nd[0] < 0) {
        int32_t lf = nd[3];
        zbo_node* leaf = new_leaf(members + leaf_off[lf], (size_t)(leaf_off[lf + 1] - leaf_off[lf]), key, depth);
        leaf->export_id = lf;
        return leaf;
    }
    zbo_node* in = (zbo_node*)calloc(1, sizeof(zbo_node));
    in->key = key;
    in->depth = depth;
    in->export_id = -1;
    in->coef = (float*)malloc(sizeof(float) * (size_t)ix->dim);
    memcpy(in->coef, coef + (size_t)nd[0] * (size_t)ix->dim, sizeof(float) * (size_t)ix->dim);
    in->constant = cst[nd[0]];
    in->left = import_node(ix, nd[1], nodes, coef, cst, leaf_off, members, child_key(key, 0), depth + 1);
    in->right = import_node(ix, nd[2], nodes, coef, cst, leaf_off, members, child_key(key, 1), depth + 1);
    return in;
}
/* Inject a forest (hyperplanes as INPUT, survey quirk Q8) over n rows whose ids are 0..n-1. */
int zbo_load_forest(zbo_index* ix, uint64_t n, const float* rows, const int32_t* nodes, const int32_t* roots,
                    const float* coef, const float* cst, const int64_t* leaf_off, const uint64_t* members) {
    for (int t = 0; t < ix->num_trees; ++t) {
        free_node(ix->roots[t]);
        ix->roots[t] = NULL;
    }
    ix->n_rows = 0;
    ix->n_live = 0;
    append_rows(ix, (size_t)n, rows);
    for (int t = 0; t < ix->num_trees; ++t)
        ix->roots[t] = import_node(ix, roots[t], nodes, coef, cst, leaf_off, members, root_key(ix->seed, t), 0);
    ix->built = 1;
    return 0;
}

static void renumber(zbo_node* nd, int32_t* next_leaf) {
    if (nd->is_leaf) nd->export_id = (*next_leaf)++;
    else {
        renumber(nd->left, next_leaf);
        renumber(nd->right, next_leaf);
    }
}
void zbo_renumber_leaves(zbo_index* ix) {
    int32_t next = 0;
    if (ix->built)
        for (int t = 0; t < ix->num_trees; ++t) renumber(ix->roots[t], &next);
}

/* bucket key of a vector in every tree: the root-to-leaf sign path (MSB-first, 1 = above/right),
 * its length, and the leaf's preorder number (call zbo_renumber_leaves / zbo_export_forest first). */
int zbo_hash(const zbo_index* ix, uint64_t n, const float* rows, uint64_t* out_keys, uint32_t* out_depth,
             int32_t* out_leaf) {
    if (!ix->built) return -1;
    for (size_t i = 0; i < n; ++i) {
        const float* x = rows + i * (size_t)ix->dim;
        for (int t = 0; t < ix->num_trees; ++t) {
            const zbo_node* nd = ix->roots[t];
            uint64_t key = 0;
            uint32_t depth = 0;
            while (!nd->is_leaf) {
                int above = zbo_point_is_above(nd->coef, nd->constant, x, ix->dim);
                key = (key << 1) | (uint64_t)above;
                ++depth;
                nd = above ? nd->right : nd->left;
            }
            out_keys[i * (size_t)ix->num_trees + (size_t)t] = key;
            out_depth[i * (size_t)ix->num_trees + (size_t)t] = depth;
            out_leaf[i * (size_t)ix->num_trees + (size_t)t] = nd->export_id;
        }
    }
    return 0;
}

/* The visit plan of one query (which leaves tree_result touches, with which budget) -- test aid for Q1/Q3.
 * Returns the number of visits; writes up to cap records {tree, leaf, nprime, live}. */
int64_t zbo_trace(const zbo_index* ix, const float* q, uint64_t top_k, int32_t* out, uint64_t cap) {
    idvec cand = {0, 0, 0};
    visitvec tr = {0, 0, 0};
    if (ix->built)
        for (int t = 0; t < ix->num_trees; ++t) tree_result(ix, q, (int)top_k, ix->roots[t], &cand, &tr, t);
    size_t m = tr.len < cap ? tr.len : (size_t)cap;
    for (size_t i = 0; i < m; ++i) {
        out[4 * i + 0] = tr.v[i].tree; out[4 * i + 1] = tr.v[i].leaf;
        out[4 * i + 2] = tr.v[i].nprime; out[4 * i + 3] = tr.v[i].live;
    }
    int64_t total = (int64_t)tr.len;
    free(tr.v);
    free(cand.v);
    return total;
}

/* The candidate set C of one query (ids, ascending) -- test aid for the Q1/Q2 known-answer tests. */
int64_t zbo_candidates(const zbo_index* ix, const float* q, uint64_t top_k, uint64_t* out, uint64_t cap) {
    idvec cand = {0, 0, 0};
    if (ix->built)
        for (int t = 0; t < ix->num_trees; ++t) tree_result(ix, q, (int)top_k, ix->roots[t], &cand, NULL, t);
    qsort(cand.v, cand.len, sizeof(uint64_t), cmp_u64);
    size_t m = 0;
    for (size_t i = 0; i < cand.len; ++i)
        if (i == 0 || cand.v[i] != cand.v[i - 1]) cand.v[m++] = cand.v[i];
    for (size_t i = 0; i < m && i < cap; ++i) out[i] = cand.v[i];
    free(cand.v);
    return (int64_t)m;
}

/* Batched pair scoring for arithmetic parity tests: out[i] = distance_bits(metric, a_i, b_i). */
void zbo_distance_bits_batch(int metric, uint64_t n, const float* a, const float* b, int dim, uint64_t* out) {
    for (size_t i = 0; i < n; ++i) out[i] = zbo_distance_bits(metric, a + i * (size_t)dim, b + i * (size_t)dim, dim);
}
void zbo_above_batch(uint64_t n, const float* coef, const float* cst, const float* x, int dim, uint8_t* out) {
    for (size_t i = 0; i < n; ++i)
        out[i] = (uint8_t)zbo_point_is_above(coef + i * (size_t)dim, cst[i], x + i * (size_t)dim, dim);
}

/* ------------------------------------------------------------------------------------------------
 * Synthetic data of BASELINE.md (not part of the reference): Philox-4x32-10, counter = (row_lo, row_hi,
 * col/4, stream), key = seed; word -> f32 exactly: x = float((int32)w >> 8) * 2^-23 in [-1, 1).
 * kind 1 (clustered): row = centre[row % 4096] + 0.25 * noise with one fmaf.  Lets the CPU legs create the
 * very rows the GPU generates without any device involvement.
 * ---------------------------------------------------------------------------------------------- */
static inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}
static inline float word_to_unit(uint32_t w) { return (float)((int32_t)w >> 8) * 1.1920928955078125e-07f; }
typedef struct { float* out; uint64_t first, stride, n, seed; uint32_t dim, kind; volatile int64_t* next; } synth_job;
static void* synth_worker(void* arg) {
    synth_job* j = (synth_job*)arg;
    uint32_t quads = (j->dim + 3) / 4;
    for (;;) {
        int64_t b = __atomic_fetch_add(j->next, 1024, __ATOMIC_RELAXED);
        if ((uint64_t)b >= j->n) break;
        uint64_t e = (uint64_t)b + 1024 < j->n ? (uint64_t)b + 1024 : j->n;
        for (uint64_t ri = (uint64_t)b; ri < e; ++ri) {
            uint64_t row = j->first + ri * j->stride;
            for (uint32_t cq = 0; cq < quads; ++cq) {
                uint32_t c[4] = {(uint32_t)row, (uint32_t)(row >> 32), cq, 0u};
                philox4x32_10(c, (uint32_t)j->seed, (uint32_t)(j->seed >> 32));
                float v[4] = {word_to_unit(c[0]), word_to_unit(c[1]), word_to_unit(c[2]), word_to_unit(c[3])};
                if (j->kind == 1) {
                    uint64_t crow = row % 4096ull;
                    uint32_t cc[4] = {(uint32_t)crow, 0u, cq, 1u};
                    philox4x32_10(cc, (uint32_t)j->seed, (uint32_t)(j->seed >> 32));
                    for (int x = 0; x < 4; ++x) v[x] = fmaf(0.25f, v[x], word_to_unit(cc[x]));
                }
                for (uint32_t x = 0; x < 4; ++x) {
                    uint32_t col = cq * 4 + x;
                    if (col < j->dim) j->out[ri * j->dim + col] = v[x];
                }
            }
        }
    }
    return NULL;
}
void zbo_synth_fill(float* out, uint64_t first_row, uint64_t row_stride, uint64_t n, uint32_t dim, uint64_t seed,
                    uint32_t kind, int nthreads) {
    volatile int64_t next = 0;
    synth_job job = {out, first_row, row_stride, n, seed, dim, kind, &next};
    if (nthreads <= 1) {
        synth_worker(&job);
        return;
    }
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthreads);
    for (int i = 0; i < nthreads; ++i) pthread_create(&th[i], NULL, synth_worker, &job);
    for (int i = 0; i < nthreads; ++i) pthread_join(th[i], NULL);
    free(th);
}
