// zb_scan3_kernel.cuh -- body of the fused leaf-tile scan, third generation (tile_scan3_kernel, launched from zb_scan.cu).
//
// Replaces, for visits of large leaves, the leaf branch of tree_result
// (/root/reference/src/database/index/lsh.rs:299-331: fetch every member, metric.distance, sort, take n) and the
// rescoring of search (:557-563).  Work unit ("tile") = (leaf, <= 16 of the queries that visit it).
//
// What changed against the second generation (zb_scan.cu, tile_scan_kernel) and why:
//   * 16 queries per tile instead of 8: the average leaf of BASELINE config 2 is visited by 12.6 queries per batch, so
//     most leaves are now read from HBM exactly once per batch (algorithmic bytes 27.2 GB -> ~14 GB per batch).
//   * register tile 16 rows x 8 queries per HALF-WARP with ONE accumulator lane per thread (thread t of a half-warp owns
//     lane t of every pair's canonical 16-lane accumulator) instead of 8 x 4 per quad with four lanes per thread.  Shared
//     memory delivers 128 thread-bytes per clock whatever the access width or broadcast pattern (measured: 4.03
//     wavefronts per LDS.128 in profiles/r01_scan_g_raw.csv), so what counts is loaded floats per FMA per thread:
//     (8 + 4) x 4 per 128 before, 16 + 8 per 128 now -- half the shared-memory pipe time, which is what bound cosine.
//   * the query tile adapts to the tile's population: QH = 1 .. 8 queries per half-warp, so a 9-query tile pays the FP32
//     work of 10, not of 16 (measured with QH in {2, 4, 6, 8}: 17.6 % of the executed FFMA2 / FADD2 were padding,
//     profiles/r02c_scan3_l2_stalls.txt).
//   * the distances of a row block are transposed through shared memory so that ONE warp owns a query's top-n' list for the
//     whole tile (no per-slab lists, no end-of-tile merge, exact filter threshold per query).
//   * row blocks of 64 rows (12 KB stages of 3 chunks): smaller tails, more stages in flight per byte of ring.
//
// Canonical arithmetic (zb_device.cuh): lane j of the 16-lane accumulator receives elements j, j+16,
// ... by one fused multiply-add each; fold x[i] = acc[i] + acc[i+8], r[i] = x[i] + x[i+4], s = (r0 + r1) + (r2 + r3).
// Here the four folds are shuffles over the half-warp (xor 8, 4, 1, 2); at every step a thread keeps half of its rows and
// hands the other half to its partner, so the fold costs 15 shuffles + 15 adds per pair in total and leaves thread t with
// the finished sums of row fold_row(t) for its QH queries.
//
// This file holds only code that is also compiled for the CPU by tests/scan3_emu.cpp (one std::thread per CUDA thread,
// mbarriers / bulk copies / shuffles emulated): every hardware-specific operation goes through a t3_* / mbar_* wrapper.
#pragma once

namespace zb {

#define T3_TEAMS 2           // independent teams per CTA: each streams its own tiles through its own ring
#define T3_QT 16             // most queries per tile (two half-warp groups of up to 8)
#define T3_RB 64             // rows per row block = rows per ring stage (4 warps x 16 rows)
// Build variants measured on a B200 (profiles/r02g_*.json; kernel ms on BASELINE config 2 with L2 / cosine, and 384-dim L2
// squared): KC 3 + QH step 1: 4.18 / 3.09 / 2.51 (the default);  KC 3 + QH step 2: 4.28 / 3.18 / 2.45;  KC 2 with the
// operands software-pipelined across stages + QH step 1: 4.50 / 3.50 / 2.53;  KC 2 + QH step 2: 4.73 / 3.57 / 2.58.
#ifndef T3_KC
#define T3_KC 3              // 16-float chunks per stage (3: 12 KB stages, a stage's chunks straight-line; 2: 8 KB stages, software-pipelined)
#endif
#ifndef T3_EPW
#define T3_EPW 0             // 1: for n' <= 32 the per-query lists are kept by a dedicated epilogue warp per team (the math warps never leave
                             // the FP32 loop + fold); 0 (default): by the math warps themselves.  Measured (profiles/r02o_*, kernel ms L2 / cosine /
                             // 384-dim L2 squared): off 4.22 / 3.11 / 2.54; on with 232 / 40 registers 4.29 / 3.33 / 2.65; on with 224 / 56: 4.20 / 3.19 / 2.58
#endif
#ifndef T3_QH_STEP
#define T3_QH_STEP 1         // granularity of the per-half-warp query count (1: QH = 1..8; 2: QH in {2, 4, 6, 8}, half the code)
#endif
#define T3_SLICE_FLOATS (T3_KC * 16)
#define T3_STAGE_BYTES (T3_RB * T3_SLICE_FLOATS * 4)   // 12288 (KC 3) or 8192 (KC 2)
// Team shape (template parameter TW of the kernel body = math warps per team): 4 (one per SM sub-partition, 16 rows of a block
// each: 128 accumulator registers per thread) or 8 (two per sub-partition, 8 rows each: half the accumulators, so twice the
// warps fit -- four math warps per scheduler).  Measured on a B200 (profiles/r02t_*): TW = 8 executes 24 % more instructions
// (twice the per-stage control, more shared-memory loads per FMA, spills at 112 registers) and loses at 768 dims with short
// lists (3.10 -> 3.49 ms), but its sixteen warps keep the 128-entry lists of top-100 queries far better (6.06 -> 4.01 ms).
template <int TW>
struct T3Shape {
    static_assert(TW == 4 || TW == 8, "math warps per team");
    static constexpr int RW = T3_RB / TW;                  // rows of a row block per math warp: 16 or 8
    static constexpr int SLOTS = T3_QT / TW;               // tile slots (query lists) a math warp owns: 4 or 2
    static constexpr int TEAM_THREADS = TW * 32;
    static constexpr int CWARPS = T3_TEAMS * TW;           // math warps of a CTA
    static constexpr int THREADS = CWARPS * 32 + 128;      // + 1 producer warpgroup (one TMA-driving warp per team): 384 or 640
};
#define T3_MAX_STAGES 8
#define T3_KL 32             // lanes of a list: the register top-n' holds KR entries per lane (KR = 1: n' <= 32; KR = 4: n' <= 128)
#define T3_KR_MAX 4
#define T3_NOPOS 0xFFFFFFFFu
#define T3_NOTILE 0xFFFFFFFFu

struct T3TileInfo {
    u32 tile, leaf, first, nqt, L, qh;
    long long moff;
};

struct T3Params {
    const u32* tile_leaf;
    const u32* tile_first;
    const u32* tile_count;
    const u32* ntiles;      // device scalar
    u32* tile_counter;      // device scalar, zeroed before launch
    const u32* order;       // visits grouped by leaf
    const u32* v_np;
    const u32* v_q;
    const u32* v_ent_off;
    Entry* entries;
    const float* queries;
    const double* q_rinv;   // [nq] 1/sqrt(|q|^2) in f64 (cosine)
    const double* bm_rinv;  // [positions] 1/sqrt(|row|^2) in f64 (cosine)
    const u32* bm_tomb;     // bit per position
    u64* stats;             // [0] visits, [1] pairs, [2] moved bytes
    u64* gthr;              // [nq] per-query bound shared by all of the query's visits: min over full lists of their n'-th key
    u32 top_k;
    int nst;                // ring depth
    int qcap;               // query capacity of a tile (16, 8 or 4: what fits in shared memory next to a useful ring)
    int kr;                 // list entries per lane (the kernel's KR)
    // METRIC 3 (L2 / L2 squared through the dot-product filter, see below and DESIGN 3.1b)
    const float* bm_n2;        // [positions] canonical squared norm of every stored row
    const float* q_n2;         // [nq] canonical squared norm of every query
    const float* leaf_n2max;   // [leaves] largest usable bm_n2 of the leaf (rows above T3_N2_LIMIT or not finite are scanned exactly)
    float ecoef;               // (4 chunks + 32) * 2^-24 * 1.01: |A - exact| <= ecoef * (|row|^2 + |query|^2)
    u64* cand;                 // [visits][32]: ordered(A) | position << 32 of the visit's 32 best rows by A
    float* cand_cut;           // [visits] A-domain cutoff: every row of the exact top-n' has A <= cut
    u8* cand_flag;             // [visits] 1: the list cannot stand for the leaf (refine scans the leaf exactly)
    // projection mode (MODE == 1, zb_index_hash on flat tables): "leaves" are row ranges of the input, "queries" are planes
    // (tp.queries = plane coefficients, order / v_q are not read: tile slot q is plane tile_first + q)
    const float* pj_cst;    // [planes] constants
    u8* pj_sign;            // [rows][pj_hp] point_is_above of every (row, plane)
    int pj_hp;
};

// ---- shared memory layout of one team (byte offsets from the team's base) ----
struct T3Layout {
    u32 stage, queries, sums, lists, meta, info, bars, total;
};
__host__ __device__ __forceinline__ T3Layout t3_layout(int nst, int dimp, int qcap, int kr) {
    T3Layout l;
    u32 o = 0;
    l.stage = o; o += (u32)nst * T3_STAGE_BYTES;
    l.queries = o; o += ((u32)qcap * (u32)dimp + 16u) * 4u;     // two regions of qcap / 2 queries, the second 64 bytes further
    o = (o + 127u) & ~127u;
    l.sums = o; o += 2u * (u32)qcap * T3_RB * 4u;               // [2][qcap][64] f32: finished sums of a row block, double buffered
    l.lists = o; o += (u32)qcap * 3u * (u32)kr * T3_KL * 4u;    // [qcap][3][kr][32] u32: key lo, key hi, position of entry lane * kr + r
    l.meta = o; o += (u32)qcap * 5u * 4u;                       // [5][qcap] u32: visit, n', query of every tile slot; METRIC 3: |q|^2, Eq (f32 bits)
    o = (o + 15u) & ~15u;
    l.info = o; o += 2u * (u32)sizeof(T3TileInfo);
    l.bars = o; o += (2u * T3_MAX_STAGES + 8u) * 8u;            // full[8], empty[8], ifull[2], qfull, qempty, sfull[2], sempty[2]
    l.total = (o + 127u) & ~127u;
    return l;
}
// Queries per half-warp for a tile of nqt queries: the FP32 work of a tile is rows x 2 QH.
__host__ __device__ __forceinline__ u32 t3_qh(u32 nqt) {
    const u32 q = nqt <= 2 ? 1u : (nqt + 1u) / 2u;
    return (q + T3_QH_STEP - 1u) / T3_QH_STEP * T3_QH_STEP;
}
// Where tile slot q sits in the query block: region q / qh (the two regions are read by the two half-warps of every warp in
// the same instruction: 64 bytes apart modulo 128, so the two 64-byte segments fall into different bank halves).
__host__ __device__ __forceinline__ u32 t3_qslot_floats(u32 q, u32 qh, int dimp, int qcap) {
    const u32 region = q / qh, j = q - region * qh;
    return region * ((u32)(qcap / 2) * (u32)dimp + 16u) + j * (u32)dimp;
}
// Row (0..15 of the warp's 16) whose finished sums thread t of a half-warp holds after the fold.
// TW = 4 (16 rows per warp): every fold step halves a thread's rows.  TW = 8 (8 rows): the xor-8, xor-4 and xor-1 steps do, the
// xor-2 step leaves threads t and t ^ 2 with the same row.
template <int TW>
__host__ __device__ __forceinline__ int t3_fold_row(int t) {
    return TW == 4 ? (((t >> 1) & 1) | ((t & 1) << 1) | (t & 4) | (t & 8)) : ((t & 1) | ((t >> 1) & 2) | ((t >> 1) & 4));
}

__device__ __forceinline__ bool t3_kp_less(u64 ka, u32 pa, u64 kb, u32 pb) { return ka < kb || (ka == kb && pa < pb); }

// Cosine epilogue from precomputed reciprocal norms: the operation order of cos_bits (zb_device.cuh) with
// ra = 1/sqrt(a2), rb = 1/sqrt(b2) hoisted (ra is +inf exactly when a2 == 0).
__device__ __forceinline__ u64 t3_cos_bits_rinv(float ab_, double ra, double rb) {
    const double ab = (double)ab_;
    double c;
    if (t3_isinf_pos(ra) && t3_isinf_pos(rb)) c = 0.0;
    else if (ab == 0.0) c = 1.0;
    else {
        double t = t3_dmul(t3_dmul(ab, ra), rb);
        double r = t3_dsub(1.0, t);
        c = r > 0.0 ? r : 0.0;
    }
    return t3_dbits(t3_dsub(1.0, c));
}

// ---- METRIC 3: L2 / L2 squared through the dot-product filter ----------------------------------------------------------
// sum (a - b)^2 costs two FP32 operations per element, a dot product one: at 768 dimensions and ~12 queries per leaf the
// exact scan is FP32-issue bound (2.33 ms floor per config-2 batch against 1.88 ms of HBM).  The fused kernel therefore scores
// A = (|a|^2 + |q|^2) - 2 a.q from the canonical dot product and precomputed canonical norms.  A differs from the EXACT
// canonical f32 value D_c (what the reference's simsimd kernel returns, what METRIC 1 / 2 compute) by at most
// Eq = ecoef * (n2max(leaf) + |q|^2) for every row of the leaf (forward error of the two accumulations, derivation in DESIGN
// 3.1b; the same bound for all rows of a leaf, so ordering by A is ordering by A -+ Eq).  The kernel keeps the 32 best rows
// by A per visit; the n' rows with the smallest A have D_c <= A_(n') + Eq, so the n'-th smallest D_c is at most that and every
// row of the exact top-n' has A <= A_(n') + 2 Eq =: cut.  A second pass (t3_refine_warp) evaluates D_c -- the operation order
// of METRIC 1 / 2 -- for the listed rows with A <= cut and sorts them by (key, position): keys, ids and ties are
// bit-identical to the exact scan.  Where the list cannot stand for the leaf (more than 32 rows under the cut, rows whose
// norms are not finite or so large that a sum may overflow) the visit is flagged and the second pass scans its leaf exactly.
// The bound shared by a query's visits (gthr) holds, for this metric, ordered(G) with G >= the query's final k-th D_c
// (G = A_(n') + Eq of a full list): a row with A > G + Eq cannot reach the final top-k.
#define T3_N2_LIMIT 1e37f    // |row|^2 + |query|^2 above this: some sum may overflow, the row is scored exactly
// Float <-> unsigned with the same order (any non-NaN float).
__host__ __device__ __forceinline__ u32 t3_ford(float x) {
    u32 b = t3_fbits(x);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float t3_funord(u32 k) { return t3_bitsf((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k); }

// ------------------------------------------------------------------------------------------------------------------
// One row block (64 rows x the tile's queries) of one math warp: consumes the block's nsl ring stages -- rows
// 16 * tw .. 16 * tw + 15 of every stage against the QH queries of the thread's half-warp (half-warp h serves tile slots
// h * QH .. h * QH + QH - 1) -- and folds the 16 accumulator lanes over the half-warp.  Returns, in sum[j], the finished
// canonical sum of (row 16 * tw + t3_fold_row(t), tile slot h * QH + j).  METRIC 0: dot product, otherwise sum of (a - b)^2.
// ------------------------------------------------------------------------------------------------------------------
template <int METRIC, int QH, int TW>
__device__ __forceinline__ void t3_block_sums(unsigned char* tb, const T3Layout& lay, const int dimp, const int chunks, const int qcap,
                                              const u32 S, const int tw, const int lane, u32& buf, u32& ph, float (&sum)[QH]) {
    const int t = lane & 15, h = lane >> 4;
    const int nsl = (chunks + T3_KC - 1) / T3_KC;
    (void)nsl;
    const u32 bar_full = smem_u32(tb + lay.bars), bar_empty = bar_full + 8 * T3_MAX_STAGES;
    const float* qp = reinterpret_cast<const float*>(tb + lay.queries) + (u32)h * ((u32)(qcap / 2) * (u32)dimp + 16u) + t;
    const float* stage0 = reinterpret_cast<const float*>(tb + lay.stage) + (tw * T3Shape<TW>::RW) * T3_SLICE_FLOATS + t;
    constexpr int RP = T3Shape<TW>::RW / 2;   // row pairs of the warp
    u64 acc[RP][QH];  // acc[u][j] = packed lane partials of rows 2u, 2u+1 against query j
#pragma unroll
    for (int u = 0; u < RP; ++u)
#pragma unroll
        for (int j = 0; j < QH; ++j) acc[u][j] = 0ull;
#if T3_KC == 2
    static_assert(TW == 4, "the 2-chunk stage exists for 4 math warps per team");
    // A ring stage holds T3_KC == 2 chunks of the block's 64 rows.  The operands of a chunk (16 row floats + QH query floats
    // per thread) are loaded from shared memory HALF A CHUNK AHEAD of the FP32 work that consumes them, across stage
    // boundaries too: a row register is reloaded with the next chunk's value as soon as its last FFMA2 has been issued, the
    // queries alternate between two register sets, and the wait on the next stage's full barrier sits in the middle of the
    // current stage's last chunk -- neither the barrier round trip nor the shared-memory latency is in front of a warp's
    // FFMA2 stream (the second warp of the sub-partition is busy with its own tile and cannot be counted on to fill the gap).
    float r[16], qa[QH], qb[QH];
    auto loadq = [&](float (&q)[QH], const float* qq_, int c) {
#pragma unroll
        for (int j = 0; j < QH; ++j) q[j] = qq_[j * dimp + c * 16];
    };
    auto loadr = [&](const float* rp, int c, int half) {
#pragma unroll
        for (int i = 0; i < 8; ++i) r[8 * half + i] = rp[(8 * half + i) * T3_SLICE_FLOATS + c * 16];
    };
    auto math = [&](const float (&q)[QH], int half) {
#pragma unroll
        for (int u = 4 * half; u < 4 * half + 4; ++u) {
            const u64 rr = t3_pk2(r[2 * u], r[2 * u + 1]);
#pragma unroll
            for (int j = 0; j < QH; ++j) {
                const u64 qq = t3_pk2(q[j], q[j]);
                if (METRIC == 0 || METRIC == 3) acc[u][j] = t3_fma2(rr, qq, acc[u][j]);
                else {
                    const u64 d = t3_sub2(rr, qq);
                    acc[u][j] = t3_fma2(d, d, acc[u][j]);
                }
            }
        }
    };
    mbar_wait(bar_full + 8 * buf, ph);
    {
        const float* rp = stage0 + (size_t)buf * (T3_STAGE_BYTES / 4);
        loadq(qa, qp, 0);
        loadr(rp, 0, 0);
        loadr(rp, 0, 1);
    }
    const int nfull = chunks / T3_KC;          // stages that hold two chunks; an odd chunk count ends with a one-chunk stage
    const bool tail = (chunks & 1) != 0;
#pragma unroll 1
    for (int sl = 0; sl < nfull; ++sl) {
        const u32 cur = buf;
        const float* rp = stage0 + (size_t)cur * (T3_STAGE_BYTES / 4);
        const bool more = sl + 1 < nfull || tail;
        if (++buf == S) { buf = 0; ph ^= 1u; }
        const float* np_ = stage0 + (size_t)buf * (T3_STAGE_BYTES / 4);
        // ---- chunk 0 of the stage: rows in r, queries in qa; chunk 1's operands arrive behind it ----
        loadq(qb, qp, 1);
        math(qa, 0);
        loadr(rp, 1, 0);
        math(qa, 1);
        loadr(rp, 1, 1);
        qp += T3_SLICE_FLOATS;
        // ---- chunk 1: queries in qb; the next stage's chunk 0 arrives behind it ----
        math(qb, 0);
        if (more) {
            mbar_wait(bar_full + 8 * buf, ph);
            loadq(qa, qp, 0);
            loadr(np_, 0, 0);
        }
        math(qb, 1);
        if (more) loadr(np_, 0, 1);
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty + 8 * cur);
    }
    if (tail) {  // warp uniform
        math(qa, 0);
        math(qa, 1);
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty + 8 * buf);
        if (++buf == S) { buf = 0; ph ^= 1u; }
    }
#else
    for (int sl = 0; sl < nsl; ++sl) {
        const int kcs = min(T3_KC, chunks - sl * T3_KC);
        mbar_wait(bar_full + 8 * buf, ph);
        const float* rp = stage0 + (size_t)buf * (T3_STAGE_BYTES / 4);
        const float* qs = qp + sl * T3_SLICE_FLOATS;
        auto chunk = [&](int c) {
            u64 qq[QH];
#pragma unroll
            for (int j = 0; j < QH; ++j) {
                const float q = qs[j * dimp + c * 16];
                qq[j] = t3_pk2(q, q);
            }
#pragma unroll
            for (int u = 0; u < RP; ++u) {
                const u64 ra = t3_pk2(rp[(2 * u) * T3_SLICE_FLOATS + c * 16], rp[(2 * u + 1) * T3_SLICE_FLOATS + c * 16]);
#pragma unroll
                for (int j = 0; j < QH; ++j) {
                    if (METRIC == 0 || METRIC == 3) acc[u][j] = t3_fma2(ra, qq[j], acc[u][j]);
                    else {
                        const u64 d = t3_sub2(ra, qq[j]);
                        acc[u][j] = t3_fma2(d, d, acc[u][j]);
                    }
                }
            }
        };
        if (kcs == T3_KC) {  // the common case, straight-line: the next chunk's shared-memory loads can be hoisted over this chunk's math
#pragma unroll
            for (int c = 0; c < T3_KC; ++c) chunk(c);
        } else {
#pragma unroll 1
            for (int c = 0; c < kcs; ++c) chunk(c);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty + 8 * buf);
        if (++buf == S) { buf = 0; ph ^= 1u; }
    }
#endif
    // ---- fold: canonical tree over the 16 lanes of the half-warp; a thread keeps half of its rows per step ----
    if constexpr (TW == 4) {
        float v8[8][QH];
        {
            const bool t3b = (t & 8) != 0;
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int j = 0; j < QH; ++j) {
                    // rows a and a + 8 of the warp's 16: row r lives in acc[r / 2][j], half r % 2
                    float lo0, hi0, lo1, hi1;
                    t3_upk2(acc[a >> 1][j], lo0, hi0);
                    t3_upk2(acc[(a + 8) >> 1][j], lo1, hi1);
                    const float va = (a & 1) ? hi0 : lo0, vb = (a & 1) ? hi1 : lo1;
                    const float mine = t3b ? vb : va, send = t3b ? va : vb;
                    v8[a][j] = t3_fadd(mine, __shfl_xor_sync(0xffffffffu, send, 8));  // lane i + lane i + 8
                }
        }
        float v4[4][QH];
        {
            const bool t2b = (t & 4) != 0;
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int j = 0; j < QH; ++j) {
                    const float mine = t2b ? v8[a + 4][j] : v8[a][j], send = t2b ? v8[a][j] : v8[a + 4][j];
                    v4[a][j] = t3_fadd(mine, __shfl_xor_sync(0xffffffffu, send, 4));  // x[i] + x[i + 4]
                }
        }
        float v2[2][QH];
        {
            const bool t0b = (t & 1) != 0;
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int j = 0; j < QH; ++j) {
                    const float mine = t0b ? v4[a + 2][j] : v4[a][j], send = t0b ? v4[a][j] : v4[a + 2][j];
                    v2[a][j] = t3_fadd(mine, __shfl_xor_sync(0xffffffffu, send, 1));  // r0 + r1, r2 + r3
                }
        }
        {
            const bool t1b = (t & 2) != 0;
#pragma unroll
            for (int j = 0; j < QH; ++j) {
                const float mine = t1b ? v2[1][j] : v2[0][j], send = t1b ? v2[0][j] : v2[1][j];
                sum[j] = t3_fadd(mine, __shfl_xor_sync(0xffffffffu, send, 2));  // (r0 + r1) + (r2 + r3)
            }
        }
    } else {
        // 8 rows: xor 8 pairs rows a / a + 4, xor 4 rows a / a + 2, xor 1 rows 0 / 1; the last step (xor 2) adds the two halves
        // (r0 + r1) and (r2 + r3) of the one row left, on both threads
        float v4[4][QH];
        {
            const bool t3b = (t & 8) != 0;
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int j = 0; j < QH; ++j) {
                    float lo0, hi0, lo1, hi1;
                    t3_upk2(acc[a >> 1][j], lo0, hi0);
                    t3_upk2(acc[(a + 4) >> 1][j], lo1, hi1);
                    const float va = (a & 1) ? hi0 : lo0, vb = (a & 1) ? hi1 : lo1;
                    const float mine = t3b ? vb : va, send = t3b ? va : vb;
                    v4[a][j] = t3_fadd(mine, __shfl_xor_sync(0xffffffffu, send, 8));  // lane i + lane i + 8
                }
        }
        float v2[2][QH];
        {
            const bool t2b = (t & 4) != 0;
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int j = 0; j < QH; ++j) {
                    const float mine = t2b ? v4[a + 2][j] : v4[a][j], send = t2b ? v4[a][j] : v4[a + 2][j];
                    v2[a][j] = t3_fadd(mine, __shfl_xor_sync(0xffffffffu, send, 4));  // x[i] + x[i + 4]
                }
        }
        {
            const bool t0b = (t & 1) != 0;
#pragma unroll
            for (int j = 0; j < QH; ++j) {
                const float mine = t0b ? v2[1][j] : v2[0][j], send = t0b ? v2[0][j] : v2[1][j];
                const float h2 = t3_fadd(mine, __shfl_xor_sync(0xffffffffu, send, 1));  // r0 + r1, r2 + r3
                sum[j] = t3_fadd(h2, __shfl_xor_sync(0xffffffffu, h2, 2));              // (r0 + r1) + (r2 + r3)
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Scan mode: one math warp's share of one tile; per row block the sums are transposed through shared memory and the warp
// maintains the lists of the tile slots it owns (slot q is owned by warp q % T3_TWARPS).
// ------------------------------------------------------------------------------------------------------------------
// The finished sums of (row T3_RW * tw + t3_fold_row(t), tile slot h * QH + j) go to sums[slot][row] of the block's buffer.
template <int METRIC, int QH, int TW>
__device__ __forceinline__ void t3_block_sums_store(unsigned char* tb, const T3Layout& lay, const int dimp, const int chunks, const int qcap,
                                                    const u32 S, const int tw, const int lane, u32& buf, u32& ph, float* dst,
                                                    const u32 wait_bar = 0, const u32 wait_parity = 0) {
    float sum[QH];
    t3_block_sums<METRIC, QH, TW>(tb, lay, dimp, chunks, qcap, S, tw, lane, buf, ph, sum);
    if (wait_bar) mbar_wait(wait_bar, wait_parity);  // T3_EPW: the epilogue warp is done with this buffer's previous block
    const int h = lane >> 4;
#pragma unroll
    for (int j = 0; j < QH; ++j) dst[(size_t)(h * QH + j) * T3_RB] = sum[j];
}

// One copy of the block loop and of the epilogue for all query counts: only the FP32 loop + fold is specialised by QH (the
// instruction working set of a warp -- loop, fold, epilogue -- has to stay inside the SM's 32 KB instruction cache while the
// two teams of a CTA run different tiles).
template <int METRIC, int KR, int TW>
__device__ __forceinline__ void t3_scan_tile(unsigned char* tb, const T3Layout& lay, const ForestView& f, const T3Params& tp,
                                             const T3TileInfo& inf, const int tw, const int lane, u32& buf, u32& ph, u32& blk,
                                             const int team, u64 (&thr)[T3Shape<TW>::SLOTS]) {
    constexpr int T3_SLOTS = T3Shape<TW>::SLOTS, T3_RW = T3Shape<TW>::RW, T3_TWARPS = TW;
    const int t = lane & 15;
    const u32 S = (u32)tp.nst, L = inf.L, nqt = inf.nqt;
    const u32 nblocks = (L + T3_RB - 1) / T3_RB;
    float* s_sums = reinterpret_cast<float*>(tb + lay.sums);
    u32* s_lists = reinterpret_cast<u32*>(tb + lay.lists);
    const u32* s_meta = reinterpret_cast<const u32*>(tb + lay.meta);
    const int myrow = t3_fold_row<TW>(t);

    // The per-query bound shared by all of a query's visits (gthr) is read branch free right after a block's FP32 loop: four
    // independent loads in flight behind the transposition and the team barrier.  A stale bound only costs work.
    u32 gq4[T3_SLOTS];
#pragma unroll
    for (int e = 0; e < T3_SLOTS; ++e) {
        const u32 q = (u32)tw + (u32)T3_TWARPS * e;
        const u32 m = s_meta[2 * tp.qcap + (q < (u32)tp.qcap ? q : 0u)];
        gq4[e] = q < nqt ? m : 0u;
    }
    for (u32 b = 0; b < nblocks; ++b, ++blk) {
        const u32 nrows = min((u32)T3_RB, L - b * T3_RB);
        const u32 base = (u32)(inf.moff + (long long)b * T3_RB);  // position of row 0 of the block

        const u32 r_lo = (u32)lane, r_hi = (u32)lane + 32u;  // the two rows of the block this lane finishes in the epilogue
        // what the epilogue reads from global memory is pulled towards the SM while the block is being scored
        if (METRIC == 0) {
            t3_prefetch_l1(tp.bm_rinv + base + r_lo);
            t3_prefetch_l1(tp.bm_rinv + base + r_hi);
        }
        if (lane < 3) t3_prefetch_l1(tp.bm_tomb + (base >> 5) + lane);
        {
            float* dst = s_sums + (size_t)(blk & 1u) * tp.qcap * T3_RB + tw * T3_RW + myrow;
            switch (inf.qh) {
#if T3_QH_STEP == 1
                case 1: t3_block_sums_store<METRIC, 1, TW>(tb, lay, f.dimp, f.chunks, tp.qcap, S, tw, lane, buf, ph, dst); break;
                case 3: t3_block_sums_store<METRIC, 3, TW>(tb, lay, f.dimp, f.chunks, tp.qcap, S, tw, lane, buf, ph, dst); break;
                case 5: t3_block_sums_store<METRIC, 5, TW>(tb, lay, f.dimp, f.chunks, tp.qcap, S, tw, lane, buf, ph, dst); break;
                case 7: t3_block_sums_store<METRIC, 7, TW>(tb, lay, f.dimp, f.chunks, tp.qcap, S, tw, lane, buf, ph, dst); break;
#endif
                case 2: t3_block_sums_store<METRIC, 2, TW>(tb, lay, f.dimp, f.chunks, tp.qcap, S, tw, lane, buf, ph, dst); break;
                case 4: t3_block_sums_store<METRIC, 4, TW>(tb, lay, f.dimp, f.chunks, tp.qcap, S, tw, lane, buf, ph, dst); break;
                case 6: t3_block_sums_store<METRIC, 6, TW>(tb, lay, f.dimp, f.chunks, tp.qcap, S, tw, lane, buf, ph, dst); break;
                default: t3_block_sums_store<METRIC, 8, TW>(tb, lay, f.dimp, f.chunks, tp.qcap, S, tw, lane, buf, ph, dst); break;
            }
        }
        u64 gbound[T3_SLOTS];
#pragma unroll
        for (int e = 0; e < T3_SLOTS; ++e) gbound[e] = t3_ldcg_u64(tp.gthr + gq4[e]);
        // ---- what the epilogue needs from global memory (prefetched above; the team barrier below hides the rest) ----
        double rinv_lo = 0.0, rinv_hi = 0.0;
        if (METRIC == 0) {
            if (r_lo < nrows) rinv_lo = tp.bm_rinv[base + r_lo];
            if (r_hi < nrows) rinv_hi = tp.bm_rinv[base + r_hi];
        }
        float n2_lo = 0.f, n2_hi = 0.f;
        if (METRIC == 3) {
            if (r_lo < nrows) n2_lo = tp.bm_n2[base + r_lo];
            if (r_hi < nrows) n2_hi = tp.bm_n2[base + r_hi];
        }
        u32 tword = 0;  // tombstone words covering positions base .. base + 63 (at most 3 words), one per lane
        if (lane < 3) tword = tp.bm_tomb[(base >> 5) + lane];
        t3_team_sync(team, T3Shape<TW>::TEAM_THREADS);  // the block's sums are complete; the previous use of this buffer was consumed two blocks ago
        // ---- epilogue: this warp finishes the tile slots it owns (q = tw, tw + T3_TWARPS, ...): keys, filter, list insertion ----
        const float* sums = s_sums + (size_t)(blk & 1u) * tp.qcap * T3_RB;
#pragma unroll
        for (int e = 0; e < T3_SLOTS; ++e) {
            const u32 q = (u32)tw + (u32)T3_TWARPS * e;
            if (q >= nqt) break;  // warp uniform
            const int np = (int)s_meta[tp.qcap + q];
            const u32 gq = s_meta[2 * tp.qcap + q];
            const float s_lo = sums[(size_t)q * T3_RB + r_lo], s_hi = sums[(size_t)q * T3_RB + r_hi];
            // every lane loaded the bound itself; other tiles lower it concurrently and the epilogue's control flow depends on it,
            // so ONE lane's copy decides for the warp (the load has long arrived: no wait here)
            gbound[e] = t3_shfl64(gbound[e], 0);
            u64 k_lo = ZB_SENTINEL, k_hi = ZB_SENTINEL;
            float Eq = 0.f;   // METRIC 3: the error bound of this (leaf, query)
            if (METRIC == 3) {
                const float n2q = t3_bitsf(s_meta[3 * tp.qcap + q]);
                Eq = t3_bitsf(s_meta[4 * tp.qcap + q]);
                // the shared bound holds ordered(G), G >= the query's final k-th exact value: a row with A > G + Eq is out
                if (gbound[e] != ZB_SENTINEL) {
                    const u64 gf = (u64)t3_ford(t3_fadd_ru(t3_funord((u32)gbound[e]), Eq));
                    if (gf < thr[e]) thr[e] = gf;
                }
                const float w_lo = t3_fadd(n2_lo, n2q), w_hi = t3_fadd(n2_hi, n2q);
                const float a_lo = t3_fmaf(-2.0f, s_lo, w_lo), a_hi = t3_fmaf(-2.0f, s_hi, w_hi);
                // rows whose sums may have overflowed (or are not numbers): key 0, always listed first -> the visit is flagged
                if (r_lo < nrows) k_lo = (w_lo <= T3_N2_LIMIT && a_lo == a_lo) ? (u64)t3_ford(a_lo) : 0ull;
                if (r_hi < nrows) k_hi = (w_hi <= T3_N2_LIMIT && a_hi == a_hi) ? (u64)t3_ford(a_hi) : 0ull;
            } else if (gbound[e] < thr[e]) thr[e] = gbound[e];
            if (METRIC == 3) {
            } else if (METRIC == 0) {
                const double qr = tp.q_rinv[gq];
                if (r_lo < nrows) k_lo = t3_cos_bits_rinv(s_lo, rinv_lo, qr);
                if (r_hi < nrows) k_hi = t3_cos_bits_rinv(s_hi, rinv_hi, qr);
            } else {
                if (r_lo < nrows) k_lo = METRIC == 1 ? l2sq_bits(s_lo) : l2_bits(s_lo);
                if (r_hi < nrows) k_hi = METRIC == 1 ? l2sq_bits(s_hi) : l2_bits(s_hi);
            }
            const u64 thr0 = thr[e];
            // most blocks have no candidate under the filter: one ballot and out
            if (!__ballot_sync(0xffffffffu, (r_lo < nrows && k_lo <= thr0) || (r_hi < nrows && k_hi <= thr0))) continue;
            // the slot's list, sorted by (key, position): entry i sits in lane i / KR, register i % KR
            u32* lst = s_lists + (size_t)q * 3 * KR * T3_KL;
            u64 Lk[KR];
            u32 Lp[KR];
#pragma unroll
            for (int r = 0; r < KR; ++r) {
                Lk[r] = ((u64)lst[(KR + r) * T3_KL + lane] << 32) | lst[r * T3_KL + lane];
                Lp[r] = lst[(2 * KR + r) * T3_KL + lane];
            }
            const int tl = (np - 1) / KR, tr = (np - 1) % KR;  // where the n'-th best sits
            u64 th = thr0;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const u64 key = i ? k_hi : k_lo;
                const bool valid = (i ? r_hi : r_lo) < nrows;
                unsigned m = __ballot_sync(0xffffffffu, valid && key <= th);
                while (m) {
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    const u64 nk = t3_shfl64(key, src);
                    if (nk > th) continue;  // the filter tightened since the ballot
                    const u32 npos = base + (u32)src + 32u * i;
                    const u32 w = __shfl_sync(0xffffffffu, tword, (int)((npos >> 5) - (base >> 5)));
                    if ((w >> (npos & 31)) & 1u) continue;  // tombstoned (D1)
                    // insertion point: the entries that precede the candidate are a prefix of the list
                    int c = 0;
#pragma unroll
                    for (int r = 0; r < KR; ++r) c += t3_kp_less(Lk[r], Lp[r], nk, npos) ? 1 : 0;
                    const int il = __popc(__ballot_sync(0xffffffffu, c == KR));  // lanes whose entries all precede it
                    if (il >= T3_KL) continue;
                    const int ir = (int)__shfl_sync(0xffffffffu, (u32)c, il);
                    if (il * KR + ir >= (METRIC == 3 ? T3_KL : np)) continue;   // METRIC 3 keeps all 32 entries (the superset refine needs)
                    const u64 upk = t3_shfl_up64(Lk[KR - 1]);
                    const u32 upp = __shfl_up_sync(0xffffffffu, Lp[KR - 1], 1);
#pragma unroll
                    for (int r = KR - 1; r >= 1; --r)
                        if (lane > il || (lane == il && r > ir)) { Lk[r] = Lk[r - 1]; Lp[r] = Lp[r - 1]; }
                    if (lane > il) { Lk[0] = upk; Lp[0] = upp; }
                    if (lane == il) {
#pragma unroll
                        for (int r = 0; r < KR; ++r)
                            if (r == ir) { Lk[r] = nk; Lp[r] = npos; }
                    }
                    u64 tk = Lk[0];
#pragma unroll
                    for (int r = 1; r < KR; ++r)
                        if (r == tr) tk = Lk[r];
                    u64 lk = t3_shfl64(tk, tl);
                    // METRIC 3: the filter is the n'-th best A plus 2 Eq (everything that could still beat it exactly)
                    if (METRIC == 3 && lk != ZB_SENTINEL) lk = (u64)t3_ford(t3_fadd_ru(t3_funord((u32)lk), t3_fadd_ru(Eq, Eq)));
                    if (lk < th) th = lk;
                }
            }
#pragma unroll
            for (int r = 0; r < KR; ++r) {
                lst[r * T3_KL + lane] = (u32)Lk[r];
                lst[(KR + r) * T3_KL + lane] = (u32)(Lk[r] >> 32);
                lst[(2 * KR + r) * T3_KL + lane] = Lp[r];
            }
            thr[e] = th;
            // publish: a full list of n' == top_k distinct rows bounds the query's final k-th best
            if (METRIC == 3) {
                // the n' listed rows have exact values <= A_(n') + Eq (unless a row without a usable A, key 0, is among them)
                if (np == (int)tp.top_k) {
                    u64 tk = Lk[0];
#pragma unroll
                    for (int r = 1; r < KR; ++r)
                        if (r == tr) tk = Lk[r];
                    const u64 kn = t3_shfl64(tk, tl);
                    if (lane == 0 && kn != ZB_SENTINEL && Lk[0] != 0ull) {
                        const u64 g = (u64)t3_ford(t3_fadd_ru(t3_funord((u32)kn), Eq));
                        if (g < gbound[e]) t3_atomic_min_u64(tp.gthr + gq, g);
                    }
                }
            } else if (lane == 0 && np == (int)tp.top_k && th < gbound[e]) t3_atomic_min_u64(tp.gthr + gq, th);
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Epilogue warp (T3_EPW, n' <= 32): one warp per team keeps ALL the tile's lists.  The math warps hand over a row block's
// finished sums through the double-buffered sums area (mbarriers sfull / sempty) and go straight on to the next block's FP32
// loop; this warp turns the sums into keys, filters and inserts, one tile slot after the other.  Per-slot state (visit, n',
// query, filter) lives in lane q's registers and is broadcast by shuffles; the lists live in shared memory.
// ------------------------------------------------------------------------------------------------------------------
template <int METRIC>
__device__ __forceinline__ void t3_epilogue_warp(unsigned char* tb, const T3Layout& lay, const ForestView& f, const T3Params& tp,
                                                 const int lane, const u32 bar_ifull, const u32 bar_qempty, const u32 bar_sfull,
                                                 const u32 bar_sempty) {
    const T3TileInfo* s_info = reinterpret_cast<const T3TileInfo*>(tb + lay.info);
    const float* s_sums = reinterpret_cast<const float*>(tb + lay.sums);
    u32* s_lists = reinterpret_cast<u32*>(tb + lay.lists);
    u32 blk = 0;
    for (u32 it = 0;; ++it) {
        mbar_wait(bar_ifull + 8 * (it & 1), (it >> 1) & 1);
        const u32 tile = s_info[it & 1].tile;
        if (tile == T3_NOTILE) break;
        const u32 nqt = s_info[it & 1].nqt, L = s_info[it & 1].L, first = s_info[it & 1].first;
        const long long moff = s_info[it & 1].moff;
        // lane q: the state of tile slot q
        u32 visit_l = 0, np_l = 0, gq_l = 0;
        double qr_l = 0.0;
        if ((u32)lane < nqt) {
            visit_l = tp.order[first + lane];
            np_l = tp.v_np[visit_l];
            gq_l = tp.v_q[visit_l];
            if (METRIC == 0) qr_l = tp.q_rinv[gq_l];
        }
        u64 thr_l = ZB_SENTINEL;
        for (u32 q = 0; q < nqt; ++q) {
            u32* lst = s_lists + (size_t)q * 3 * T3_KL;
            lst[lane] = 0xFFFFFFFFu;
            lst[T3_KL + lane] = 0xFFFFFFFFu;
            lst[2 * T3_KL + lane] = T3_NOPOS;
        }
        __syncwarp();
        const u32 nblocks = (L + T3_RB - 1) / T3_RB;
        for (u32 b = 0; b < nblocks; ++b, ++blk) {
            const u32 nrows = min((u32)T3_RB, L - b * T3_RB);
            const u32 base = (u32)(moff + (long long)b * T3_RB);
            const u32 r_lo = (u32)lane, r_hi = (u32)lane + 32u;
            // what this block needs from global memory, requested before the wait for its sums
            double rinv_lo = 0.0, rinv_hi = 0.0;
            if (METRIC == 0) {
                if (r_lo < nrows) rinv_lo = tp.bm_rinv[base + r_lo];
                if (r_hi < nrows) rinv_hi = tp.bm_rinv[base + r_hi];
            }
            u32 tword = 0;
            if (lane < 3) tword = tp.bm_tomb[(base >> 5) + lane];
            const u64 gb_l = t3_ldcg_u64(tp.gthr + gq_l);   // lanes >= nqt read gthr[0]: unused
            if (gb_l < thr_l) thr_l = gb_l;
            mbar_wait(bar_sfull + 8 * (blk & 1u), (blk >> 1) & 1u);
            const float* sums = s_sums + (size_t)(blk & 1u) * tp.qcap * T3_RB;
#pragma unroll 1
            for (u32 q = 0; q < nqt; ++q) {
                const u64 thr0 = t3_shfl64(thr_l, (int)q);
                const float s_lo = sums[(size_t)q * T3_RB + r_lo], s_hi = sums[(size_t)q * T3_RB + r_hi];
                u64 k_lo = ZB_SENTINEL, k_hi = ZB_SENTINEL;
                if (METRIC == 0) {
                    const double qr = t3_shfl_f64(qr_l, (int)q);
                    if (r_lo < nrows) k_lo = t3_cos_bits_rinv(s_lo, rinv_lo, qr);
                    if (r_hi < nrows) k_hi = t3_cos_bits_rinv(s_hi, rinv_hi, qr);
                } else {
                    if (r_lo < nrows) k_lo = METRIC == 1 ? l2sq_bits(s_lo) : l2_bits(s_lo);
                    if (r_hi < nrows) k_hi = METRIC == 1 ? l2sq_bits(s_hi) : l2_bits(s_hi);
                }
                // most blocks have no candidate under the filter: one ballot and on to the next slot
                if (!__ballot_sync(0xffffffffu, (r_lo < nrows && k_lo <= thr0) || (r_hi < nrows && k_hi <= thr0))) continue;
                const int np = (int)__shfl_sync(0xffffffffu, np_l, (int)q);
                u32* lst = s_lists + (size_t)q * 3 * T3_KL;
                u64 Lk = ((u64)lst[T3_KL + lane] << 32) | lst[lane];
                u32 Lp = lst[2 * T3_KL + lane];
                u64 th = thr0;
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const u64 key = i ? k_hi : k_lo;
                    const bool valid = (i ? r_hi : r_lo) < nrows;
                    unsigned m = __ballot_sync(0xffffffffu, valid && key <= th);
                    while (m) {
                        const int src = __ffs(m) - 1;
                        m &= m - 1;
                        const u64 nk = t3_shfl64(key, src);
                        if (nk > th) continue;  // the filter tightened since the ballot
                        const u32 npos = base + (u32)src + 32u * i;
                        const u32 w = __shfl_sync(0xffffffffu, tword, (int)((npos >> 5) - (base >> 5)));
                        if ((w >> (npos & 31)) & 1u) continue;  // tombstoned (D1)
                        const unsigned mm = __ballot_sync(0xffffffffu, t3_kp_less(nk, npos, Lk, Lp));
                        const int ins = mm ? __ffs(mm) - 1 : 32;
                        if (ins >= np) continue;
                        const u64 upk = t3_shfl_up64(Lk);
                        const u32 upp = __shfl_up_sync(0xffffffffu, Lp, 1);
                        if (lane > ins) { Lk = upk; Lp = upp; }
                        else if (lane == ins) { Lk = nk; Lp = npos; }
                        const u64 lk = t3_shfl64(Lk, np - 1);
                        if (lk < th) th = lk;
                    }
                }
                lst[lane] = (u32)Lk;
                lst[T3_KL + lane] = (u32)(Lk >> 32);
                lst[2 * T3_KL + lane] = Lp;
                // publish: a full list of n' == top_k distinct rows bounds the query's final k-th best
                if ((u32)lane == q) {
                    if (np == (int)tp.top_k && th < thr_l) t3_atomic_min_u64(tp.gthr + gq_l, th);
                    thr_l = th;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_sempty + 8 * (blk & 1u));  // the block's sums buffer may be overwritten
        }
        // ---- end of tile: every visit's list ----
#pragma unroll 1
        for (u32 q = 0; q < nqt; ++q) {
            const u32 v = __shfl_sync(0xffffffffu, visit_l, (int)q);
            const int np = (int)__shfl_sync(0xffffffffu, np_l, (int)q);
            const u32* lst = s_lists + (size_t)q * 3 * T3_KL;
            const u64 k = ((u64)lst[T3_KL + lane] << 32) | lst[lane];
            const u32 p = lst[2 * T3_KL + lane];
            const u32 e0 = tp.v_ent_off[v], e1 = tp.v_ent_off[v + 1];
            if ((u32)lane < e1 - e0) {
                Entry en{ZB_SENTINEL, ZB_SENTINEL};
                if (lane < np && p != T3_NOPOS) en = Entry{k, f.ord[f.members[p]]};
                tp.entries[e0 + lane] = en;
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_qempty);  // this warp is done with the tile too: its info slot and lists may be reused
    }
}

// Math-warp side of T3_EPW: a tile is its row blocks' FP32 loops + folds, the sums handed to the epilogue warp.
template <int METRIC>
__device__ __forceinline__ void t3_scan_tile_math(unsigned char* tb, const T3Layout& lay, const ForestView& f, const T3Params& tp,
                                                  const T3TileInfo& inf, const int tw, const int lane, u32& buf, u32& ph, u32& blk,
                                                  const u32 bar_sfull, const u32 bar_sempty) {
    const u32 S = (u32)tp.nst;
    const u32 nblocks = (inf.L + T3_RB - 1) / T3_RB;
    float* s_sums = reinterpret_cast<float*>(tb + lay.sums);
    const int myrow = t3_fold_row<4>(lane & 15);
    for (u32 b = 0; b < nblocks; ++b, ++blk) {
        float* dst = s_sums + (size_t)(blk & 1u) * tp.qcap * T3_RB + tw * 16 + myrow;
        // the buffer's previous content (two blocks ago) has been consumed: checked right before the store inside, i.e. after the FP32 loop
        switch (inf.qh) {
#if T3_QH_STEP == 1
            case 1: t3_block_sums_store<METRIC, 1, 4>(tb, lay, f.dimp, f.chunks, tp.qcap, S, tw, lane, buf, ph, dst, bar_sempty + 8 * (blk & 1u), ((blk >> 1) & 1u) ^ 1u); break;
            case 3: t3_block_sums_store<METRIC, 3, 4>(tb, lay, f.dimp, f.chunks, tp.qcap, S, tw, lane, buf, ph, dst, bar_sempty + 8 * (blk & 1u), ((blk >> 1) & 1u) ^ 1u); break;
            case 5: t3_block_sums_store<METRIC, 5, 4>(tb, lay, f.dimp, f.chunks, tp.qcap, S, tw, lane, buf, ph, dst, bar_sempty + 8 * (blk & 1u), ((blk >> 1) & 1u) ^ 1u); break;
            case 7: t3_block_sums_store<METRIC, 7, 4>(tb, lay, f.dimp, f.chunks, tp.qcap, S, tw, lane, buf, ph, dst, bar_sempty + 8 * (blk & 1u), ((blk >> 1) & 1u) ^ 1u); break;
#endif
            case 2: t3_block_sums_store<METRIC, 2, 4>(tb, lay, f.dimp, f.chunks, tp.qcap, S, tw, lane, buf, ph, dst, bar_sempty + 8 * (blk & 1u), ((blk >> 1) & 1u) ^ 1u); break;
            case 4: t3_block_sums_store<METRIC, 4, 4>(tb, lay, f.dimp, f.chunks, tp.qcap, S, tw, lane, buf, ph, dst, bar_sempty + 8 * (blk & 1u), ((blk >> 1) & 1u) ^ 1u); break;
            case 6: t3_block_sums_store<METRIC, 6, 4>(tb, lay, f.dimp, f.chunks, tp.qcap, S, tw, lane, buf, ph, dst, bar_sempty + 8 * (blk & 1u), ((blk >> 1) & 1u) ^ 1u); break;
            default: t3_block_sums_store<METRIC, 8, 4>(tb, lay, f.dimp, f.chunks, tp.qcap, S, tw, lane, buf, ph, dst, bar_sempty + 8 * (blk & 1u), ((blk >> 1) & 1u) ^ 1u); break;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_sfull + 8 * (blk & 1u));
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Projection mode (flat-table hashing, Hyperplane::point_is_above of lsh.rs:39-43 for every (row, plane)): the tile's
// "queries" are planes tile_first .. tile_first + nqt - 1, its "leaf" is a range of input rows; the thread that holds a
// finished dot product tests its sign and stores it.  No lists, no transposition.
// ------------------------------------------------------------------------------------------------------------------
template <int QH, int TW>
__device__ __forceinline__ void t3_project_tile(unsigned char* tb, const T3Layout& lay, const ForestView& f, const T3Params& tp,
                                                const T3TileInfo& inf, const int tw, const int lane, u32& buf, u32& ph) {
    const int t = lane & 15, h = lane >> 4;
    const u32 S = (u32)tp.nst, L = inf.L, nqt = inf.nqt;
    const u32 nblocks = (L + T3_RB - 1) / T3_RB;
    const int myrow = tw * T3Shape<TW>::RW + t3_fold_row<TW>(t);
    float cst[QH];
#pragma unroll
    for (int j = 0; j < QH; ++j) {
        const u32 q = (u32)(h * QH + j);
        cst[j] = q < nqt ? tp.pj_cst[inf.first + q] : 0.0f;
    }
    for (u32 b = 0; b < nblocks; ++b) {
        const u32 nrows = min((u32)T3_RB, L - b * T3_RB);
        float sum[QH];
        t3_block_sums<0, QH, TW>(tb, lay, f.dimp, f.chunks, tp.qcap, S, tw, lane, buf, ph, sum);
        if ((u32)myrow < nrows) {
            u8* dst = tp.pj_sign + (size_t)(inf.moff + (long long)b * T3_RB + myrow) * tp.pj_hp + inf.first + h * QH;
#pragma unroll
            for (int j = 0; j < QH; ++j)
                if ((u32)(h * QH + j) < nqt) dst[j] = t3_above_from_dot(sum[j], cst[j]) ? 1 : 0;
        }
    }
}

// Two teams per CTA, each = 4 math warps (one per SM sub-partition) + 1 producer warp, each streaming its own tiles:
// the two math warps that share a sub-partition belong to different tiles, so one warp's fold / epilogue / tile change
// overlaps the other's FP32 loop, and a bandwidth-bound tile (few queries) shares the SM with a pipe-bound one.
// MODE 0: leaf scan (METRIC 0 cosine, 1 L2 squared, 2 L2).  MODE 1: flat-table projection (METRIC ignored).
template <int METRIC, int MODE, int KR, int TW>
__device__ __forceinline__ void t3_body(const T3Map& tmap, const ForestView& f, const T3Params& tp, unsigned char* smem) {
    constexpr int T3_TWARPS = TW, T3_CWARPS = T3Shape<TW>::CWARPS, T3_SLOTS = T3Shape<TW>::SLOTS;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int dimp = f.dimp, chunks = f.chunks;
    const int nsl = (chunks + T3_KC - 1) / T3_KC;
    const u32 S = (u32)tp.nst;
    const T3Layout lay = t3_layout(tp.nst, dimp, tp.qcap, tp.kr);
    const int team = warp < T3_CWARPS ? warp / T3_TWARPS : (warp - T3_CWARPS) % T3_TEAMS;
    unsigned char* tb = smem + (size_t)team * lay.total;
    T3TileInfo* s_info = reinterpret_cast<T3TileInfo*>(tb + lay.info);
    const u32 bar_full = smem_u32(tb + lay.bars), bar_empty = bar_full + 8 * T3_MAX_STAGES;
    const u32 bar_ifull = bar_full + 16 * T3_MAX_STAGES, bar_qfull = bar_ifull + 16, bar_qempty = bar_ifull + 24;
    const u32 bar_sfull = bar_ifull + 32, bar_sempty = bar_ifull + 48;
    constexpr bool EPW = T3_EPW && MODE == 0 && KR == 1 && METRIC != 3 && TW == 4;

    if (tid == 0) {
        for (int tm = 0; tm < T3_TEAMS; ++tm) {
            const u32 o = (u32)(tm * lay.total);  // this thread is in team 0: the other teams' barriers sit lay.total apart
            for (u32 i = 0; i < S; ++i) {
                mbar_init(bar_full + o + 8 * i, 1);
                mbar_init(bar_empty + o + 8 * i, T3_TWARPS);
            }
            mbar_init(bar_ifull + o, 1);
            mbar_init(bar_ifull + o + 8, 1);
            mbar_init(bar_qfull + o, 1);
            mbar_init(bar_qempty + o, T3_TWARPS + (EPW ? 1 : 0));
            for (u32 i = 0; i < 2; ++i) {
                mbar_init(bar_sfull + o + 8 * i, T3_TWARPS);
                mbar_init(bar_sempty + o + 8 * i, 1);
            }
        }
        t3_fence_barrier_init();
    }
    __syncthreads();

    if (warp >= T3_CWARPS) {
        // =========================== producer warpgroup: one thread per team drives TMA ===========================
        t3_setmaxnreg_dec<TW>();
        if (warp >= T3_CWARPS + T3_TEAMS) {  // the warpgroup's other two warps: one epilogue warp per team, or nothing
            if (EPW) t3_epilogue_warp<METRIC>(tb, lay, f, tp, lane, bar_ifull, bar_qempty, bar_sfull, bar_sempty);
            return;
        }
        if (lane != 0) return;
        t3_prefetch_map(tmap);
        float* s_q = reinterpret_cast<float*>(tb + lay.queries);
        const u32 ntiles = *tp.ntiles;
        u32 buf = 0, eph = 1;  // ring slot of the next stage to issue; parity to wait for on its empty barrier (first lap: free)
        u32 tile = atomicAdd(tp.tile_counter, 1u);
        for (u32 it = 0;; ++it) {
            T3TileInfo* inf = s_info + (it & 1);
            if (tile >= ntiles) {
                inf->tile = T3_NOTILE;
                mbar_arrive(bar_ifull + 8 * (it & 1));
                break;
            }
            const u32 leaf = tp.tile_leaf[tile], first = tp.tile_first[tile], nqt = tp.tile_count[tile];
            const u32 L = f.leaf_len[leaf];
            const long long moff = f.leaf_off[leaf];
            const u32 qh = t3_qh(nqt);
            inf->tile = tile; inf->leaf = leaf; inf->first = first; inf->nqt = nqt; inf->L = L; inf->qh = qh; inf->moff = moff;
            mbar_arrive(bar_ifull + 8 * (it & 1));
            const u32 nblocks = (L + T3_RB - 1) / T3_RB;
            const u32 total = nblocks * (u32)nsl;
            u32 b = 0, sl = 0;
            auto issue = [&](u32 count) {
                for (u32 j = 0; j < count; ++j) {
                    mbar_wait(bar_empty + 8 * buf, eph);
                    mbar_arrive_expect_tx(bar_full + 8 * buf, T3_STAGE_BYTES);
                    t3_tma_2d_g2s(smem_u32(tb + lay.stage + (size_t)buf * T3_STAGE_BYTES), tmap, (int)(sl * T3_SLICE_FLOATS),
                                  (long long)(moff + (long long)b * T3_RB), bar_full + 8 * buf);
                    if (++sl == (u32)nsl) { sl = 0; ++b; }
                    if (++buf == S) { buf = 0; eph ^= 1u; }
                }
            };
            // rows of this tile may run ahead into the ring while the math warps still finish the previous tile ...
            const u32 pre = total < S ? total : S;
            issue(pre);
            // ... the resident query block is single-buffered: wait until the previous tile is done with it
            if (it > 0) mbar_wait(bar_qempty, (it - 1) & 1);
            mbar_arrive_expect_tx(bar_qfull, nqt * (u32)dimp * 4u);
            {  // address loads batched (independent chains), then the copies
                u32 qi[T3_QT];
#pragma unroll
                for (int j = 0; j < T3_QT; ++j) qi[j] = (u32)j < nqt ? (MODE == 1 ? first + (u32)j : tp.order[first + j]) : 0u;
                if (MODE == 0) {
#pragma unroll
                    for (int j = 0; j < T3_QT; ++j) qi[j] = (u32)j < nqt ? tp.v_q[qi[j]] : 0u;
                }
#pragma unroll
                for (int j = 0; j < T3_QT; ++j)
                    if ((u32)j < nqt)
                        bulk_g2s(smem_u32(s_q + t3_qslot_floats((u32)j, qh, dimp, tp.qcap)), tp.queries + (size_t)qi[j] * dimp,
                                 (u32)dimp * 4u, bar_qfull);
            }
            issue(total - pre);
            // fetched only now (not while the tile is in flight), so consecutive tiles start back to back
            const u32 next_tile = atomicAdd(tp.tile_counter, 1u);
            atomicAdd(&tp.stats[0], (u64)nqt);
            atomicAdd(&tp.stats[1], (u64)nqt * L);
            atomicAdd(&tp.stats[2], ((u64)L + nqt) * (u64)dimp * 4ull);  // algorithmic bytes: leaf rows once + the tile's queries
            tile = next_tile;
        }
        return;
    }

    // =================================== consumer (math) warps ===================================
    t3_setmaxnreg_inc<TW>();
    const int tw = warp % T3_TWARPS;
    u32 rbuf = 0, rph = 0, blk = 0;  // ring position of the next stage to consume (slot, phase parity); running row-block count
    u32* s_lists = reinterpret_cast<u32*>(tb + lay.lists);
    u32* s_meta = reinterpret_cast<u32*>(tb + lay.meta);
    for (u32 it = 0;; ++it) {
        mbar_wait(bar_ifull + 8 * (it & 1), (it >> 1) & 1);
        const T3TileInfo inf = s_info[it & 1];
        if (inf.tile == T3_NOTILE) break;
        const u32 nqt = inf.nqt;
        if (MODE == 1) {
            mbar_wait(bar_qfull, it & 1);
            switch (inf.qh) {
#if T3_QH_STEP == 1
                case 1: t3_project_tile<1, TW>(tb, lay, f, tp, inf, tw, lane, rbuf, rph); break;
                case 3: t3_project_tile<3, TW>(tb, lay, f, tp, inf, tw, lane, rbuf, rph); break;
                case 5: t3_project_tile<5, TW>(tb, lay, f, tp, inf, tw, lane, rbuf, rph); break;
                case 7: t3_project_tile<7, TW>(tb, lay, f, tp, inf, tw, lane, rbuf, rph); break;
#endif
                case 2: t3_project_tile<2, TW>(tb, lay, f, tp, inf, tw, lane, rbuf, rph); break;
                case 4: t3_project_tile<4, TW>(tb, lay, f, tp, inf, tw, lane, rbuf, rph); break;
                case 6: t3_project_tile<6, TW>(tb, lay, f, tp, inf, tw, lane, rbuf, rph); break;
                default: t3_project_tile<8, TW>(tb, lay, f, tp, inf, tw, lane, rbuf, rph); break;
            }
            if (lane == 0) mbar_arrive(bar_qempty);
            continue;
        }
        if (EPW) {
            mbar_wait(bar_qfull, it & 1);
            t3_scan_tile_math<METRIC>(tb, lay, f, tp, inf, tw, lane, rbuf, rph, blk, bar_sfull, bar_sempty);
            if (lane == 0) mbar_arrive(bar_qempty);
            continue;
        }
        // the tile slots this warp owns: visit, n', query -> shared memory (read back as broadcasts in the epilogue), empty lists
        if (lane < T3_SLOTS) {
            const u32 q = (u32)tw + (u32)T3_TWARPS * lane;
            if (q < nqt) {
                const u32 visit = tp.order[inf.first + q];
                s_meta[q] = visit;
                s_meta[tp.qcap + q] = tp.v_np[visit];
                const u32 gq = tp.v_q[visit];
                s_meta[2 * tp.qcap + q] = gq;
                if (METRIC == 3) {
                    const float n2q = tp.q_n2[gq];
                    s_meta[3 * tp.qcap + q] = t3_fbits(n2q);
                    s_meta[4 * tp.qcap + q] = t3_fbits(t3_fadd_ru(t3_fmul_ru(tp.ecoef, t3_fadd_ru(tp.leaf_n2max[inf.leaf], n2q)), 1e-37f));
                }
            }
        }
#pragma unroll
        for (int e = 0; e < T3_SLOTS; ++e) {
            const u32 q = (u32)tw + (u32)T3_TWARPS * e;
            if (q < nqt) {
                u32* lst = s_lists + (size_t)q * 3 * KR * T3_KL;
#pragma unroll
                for (int x = 0; x < 3 * KR; ++x) lst[x * T3_KL + lane] = 0xFFFFFFFFu;  // empty: key all ones, position T3_NOPOS
            }
        }
        u64 thr[T3_SLOTS];  // filter of each owned slot: its list's n'-th key or the shared bound
#pragma unroll
        for (int e = 0; e < T3_SLOTS; ++e) thr[e] = ZB_SENTINEL;
        __syncwarp();
        mbar_wait(bar_qfull, it & 1);
        t3_scan_tile<METRIC, KR, TW>(tb, lay, f, tp, inf, tw, lane, rbuf, rph, blk, team, thr);
        // ---- end of tile: release the query block, write the owned visits' top lists ----
        if (lane == 0) mbar_arrive(bar_qempty);
#pragma unroll 1
        for (int e = 0; e < T3_SLOTS; ++e) {
            const u32 q = (u32)tw + (u32)T3_TWARPS * e;
            if (q >= nqt) break;
            const u32 v = s_meta[q];
            const int np = (int)s_meta[tp.qcap + q];
            const u32* lst = s_lists + (size_t)q * 3 * KR * T3_KL;
            if (METRIC == 3) {  // the visit's 32 best rows by A, the cutoff below which the exact top-n' lies, the flag
                const u64 k = ((u64)lst[KR * T3_KL + lane] << 32) | lst[lane];
                const u32 p = lst[2 * KR * T3_KL + lane];
                tp.cand[(size_t)v * T3_KL + lane] = (u64)(u32)k | ((u64)(k == ZB_SENTINEL ? T3_NOPOS : p) << 32);
                const u64 kn = t3_shfl64(k, np - 1), k31 = t3_shfl64(k, T3_KL - 1), k0 = t3_shfl64(k, 0);
                if (lane == 0) {
                    const float Eq = t3_bitsf(s_meta[4 * tp.qcap + q]);
                    float cut = t3_bitsf(0x7F800000u);  // +inf: fewer than n' rows listed, all of them are candidates
                    bool over = k0 == 0ull;             // a row without a usable A: the leaf is scanned exactly
                    if (!over && kn != ZB_SENTINEL) {
                        cut = t3_fadd_ru(t3_funord((u32)kn), t3_fadd_ru(Eq, Eq));
                        over = !(cut == cut) || (k31 != ZB_SENTINEL && (u32)k31 <= t3_ford(cut));
                    }
                    tp.cand_cut[v] = cut;
                    tp.cand_flag[v] = over ? 1 : 0;
                }
                continue;
            }
            const u32 e0 = tp.v_ent_off[v], e1 = tp.v_ent_off[v + 1];
#pragma unroll
            for (int r = 0; r < KR; ++r) {
                const u64 k = ((u64)lst[(KR + r) * T3_KL + lane] << 32) | lst[r * T3_KL + lane];
                const u32 p = lst[(2 * KR + r) * T3_KL + lane];
                const u32 idx = (u32)lane * KR + r;
                if (idx < e1 - e0) {
                    Entry en{ZB_SENTINEL, ZB_SENTINEL};
                    if ((int)idx < np && p != T3_NOPOS) en = Entry{k, f.ord[f.members[p]]};
                    tp.entries[e0 + idx] = en;
                }
            }
        }
        __syncwarp();  // the slots' shared-memory state may be re-initialised for the next tile
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Second pass of METRIC 3 (refine_visits_kernel): one warp per visit the fused kernel scored through the dot-product filter.
// The warp evaluates the EXACT canonical sum of (a - b)^2 -- one quad per row, thread `sub` of the quad keeps accumulator
// lanes 4 sub .. 4 sub + 3, folds by xor 2, xor 1, (r0 + r1) + (r2 + r3): the order of zb_device.cuh -- for the visit's
// candidates (listed rows with A <= cut and A <= G + Eq), or for every live row of the leaf when the visit is flagged,
// and keeps the n' smallest (key, position) in a list of one entry per lane: what METRIC 1 / 2 write for the visit.
// Visits are handed out through one counter (flagged visits take ~100 times longer than the others).
// ------------------------------------------------------------------------------------------------------------------
struct T3RefineParams {
    const float* bm_rows;      // [positions][dimp]
    const u32* bm_tomb;
    const float* queries;      // [nq][dimp]
    const u32* order;          // the fused kernel's visits, grouped by leaf
    const u32* nvisits;        // device scalar: how many
    const u32* v_leaf;
    const u32* v_np;
    const u32* v_q;
    const u32* v_ent_off;
    const u64* cand;
    const float* cand_cut;
    const u8* cand_flag;
    const u64* gthr;
    const float* q_n2;
    const float* leaf_n2max;
    float ecoef;
    Entry* entries;
    u32* work_counter;         // zeroed before launch
    u64* stats;                // [4] flagged visits, [5] rows scored exactly
};

template <int METRIC>   // 1: L2 squared, 2: L2 (the key of the exact sum)
__device__ __forceinline__ void t3_refine_warp(const ForestView& f, const T3RefineParams& rp, const int lane) {
    const int sub = lane & 3, quad = lane >> 2;
    const u32 nvis = *rp.nvisits;
    for (;;) {
        u32 i = 0;
        if (lane == 0) i = atomicAdd(rp.work_counter, 1u);
        i = __shfl_sync(0xffffffffu, i, 0);
        if (i >= nvis) break;
        const u32 v = rp.order[i];
        const u32 leaf = rp.v_leaf[v], gq = rp.v_q[v];
        const int np = (int)rp.v_np[v];
        const bool flagged = rp.cand_flag[v] != 0;
        const u64 G = t3_ldcg_u64(rp.gthr + gq);
        // exact-key filter: the n'-th key of the list so far, or the key of G (k distinct rows at or below it exist)
        u64 th = ZB_SENTINEL;
        if (G != ZB_SENTINEL) th = METRIC == 1 ? l2sq_bits(t3_funord((u32)G)) : l2_bits(t3_funord((u32)G));
        const long long moff = f.leaf_off[leaf];
        u32 nitems, cpos = T3_NOPOS;
        unsigned cmask = 0;
        if (flagged) nitems = f.leaf_len[leaf];
        else {
            const u64 c = rp.cand[(size_t)v * T3_KL + lane];
            const u32 oa = (u32)c;
            cpos = (u32)(c >> 32);
            bool is = cpos != T3_NOPOS && oa <= t3_ford(rp.cand_cut[v]);
            if (G != ZB_SENTINEL) {
                const float Eq = t3_fadd_ru(t3_fmul_ru(rp.ecoef, t3_fadd_ru(rp.leaf_n2max[leaf], rp.q_n2[gq])), 1e-37f);
                is = is && oa <= t3_ford(t3_fadd_ru(t3_funord((u32)G), Eq));
            }
            cmask = __ballot_sync(0xffffffffu, is);
            nitems = (u32)__popc(cmask);
        }
        if (lane == 0) {
            if (flagged) atomicAdd(&rp.stats[4], 1ull);
            atomicAdd(&rp.stats[5], (u64)nitems);
        }
        u64 Lk = ZB_SENTINEL;   // entry `lane` of the list, sorted by (key, position)
        u32 Lp = T3_NOPOS;
        const float* qrow = rp.queries + (size_t)gq * f.dimp + 4 * sub;
        unsigned rest = cmask;
        for (u32 base = 0; base < nitems; base += 8) {
            // the row of this quad in this round
            u32 p = T3_NOPOS;
            if (flagged) {
                const u32 item = base + (u32)quad;
                if (item < nitems) {
                    p = (u32)(moff + item);
                    if ((rp.bm_tomb[p >> 5] >> (p & 31)) & 1u) p = T3_NOPOS;   // tombstoned (D1)
                }
            } else {
#pragma unroll 1
                for (int j = 0; j < 8 && rest; ++j) {   // the next eight listed candidates, in list order
                    const int src = __ffs(rest) - 1;
                    rest &= rest - 1;
                    const u32 pp = __shfl_sync(0xffffffffu, cpos, src);
                    if (quad == j) p = pp;
                }
            }
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            if (p != T3_NOPOS) {
                const float* row = rp.bm_rows + (size_t)p * f.dimp + 4 * sub;
                // (measured, profiles/r02u_bench_l2.json: twelve ordered requests per thread at 116 registers and 16 warps per SM took
                // 0.22 ms per config-2 batch against 0.18 ms for this loop at 48 registers and 48 warps per SM)
#pragma unroll 4
                for (int c = 0; c < f.chunks; ++c) {
                    const T3F4 r = t3_ld_f4(row + c * 16), q = t3_ld_f4(qrow + c * 16);
                    const float d0 = t3_fsub(r.x, q.x), d1 = t3_fsub(r.y, q.y), d2 = t3_fsub(r.z, q.z), d3 = t3_fsub(r.w, q.w);
                    a0 = t3_fmaf(d0, d0, a0);
                    a1 = t3_fmaf(d1, d1, a1);
                    a2 = t3_fmaf(d2, d2, a2);
                    a3 = t3_fmaf(d3, d3, a3);
                }
            }
            a0 = t3_fadd(a0, __shfl_xor_sync(0xffffffffu, a0, 2));   // lane i + lane i + 8
            a1 = t3_fadd(a1, __shfl_xor_sync(0xffffffffu, a1, 2));
            a2 = t3_fadd(a2, __shfl_xor_sync(0xffffffffu, a2, 2));
            a3 = t3_fadd(a3, __shfl_xor_sync(0xffffffffu, a3, 2));
            a0 = t3_fadd(a0, __shfl_xor_sync(0xffffffffu, a0, 1));   // x[i] + x[i + 4]
            a1 = t3_fadd(a1, __shfl_xor_sync(0xffffffffu, a1, 1));
            a2 = t3_fadd(a2, __shfl_xor_sync(0xffffffffu, a2, 1));
            a3 = t3_fadd(a3, __shfl_xor_sync(0xffffffffu, a3, 1));
            const float ssum = t3_fadd(t3_fadd(a0, a1), t3_fadd(a2, a3));   // (r0 + r1) + (r2 + r3)
            const u64 key = METRIC == 1 ? l2sq_bits(ssum) : l2_bits(ssum);
            unsigned m = __ballot_sync(0xffffffffu, sub == 0 && p != T3_NOPOS && key <= th);
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const u64 nk = t3_shfl64(key, src);
                const u32 npos = __shfl_sync(0xffffffffu, p, src);
                if (nk > th) continue;  // the filter tightened since the ballot
                const unsigned mm = __ballot_sync(0xffffffffu, t3_kp_less(nk, npos, Lk, Lp));
                const int ins = mm ? __ffs(mm) - 1 : 32;
                if (ins >= np) continue;
                const u64 upk = t3_shfl_up64(Lk);
                const u32 upp = __shfl_up_sync(0xffffffffu, Lp, 1);
                if (lane > ins) { Lk = upk; Lp = upp; }
                else if (lane == ins) { Lk = nk; Lp = npos; }
                const u64 lk = t3_shfl64(Lk, np - 1);
                if (lk < th) th = lk;
            }
        }
        const u32 e0 = rp.v_ent_off[v], e1 = rp.v_ent_off[v + 1];
        if ((u32)lane < e1 - e0) {
            Entry en{ZB_SENTINEL, ZB_SENTINEL};
            if (lane < np && Lp != T3_NOPOS) en = Entry{Lk, f.ord[f.members[Lp]]};
            rp.entries[e0 + lane] = en;
        }
    }
}

}  // namespace zb
