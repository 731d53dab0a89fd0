// quadtile_emu.cpp -- TEST INFRASTRUCTURE.  Runs the SOURCE of quad_tile_kernel (zebra_b200/csrc/zb_quadtile_kernel.cuh) on the
// CPU: one std::thread per CUDA thread of a block, __syncthreads = a block barrier, __shfl_xor_sync = an exchange through
// a per-quad mailbox, shared memory = block-local storage, atomics = GCC atomics, the device helpers (quad_reduce16, cos_bits, ...)
// restated with plain IEEE operations (build with -ffp-contract=off, and with -fvisibility=hidden -Wl,-Bsymbolic: the host
// stub nvcc emits for the real kernel in libzebra_b200.so has the same mangled name as the emulated function).  Blocks run one after the other, which is a legal
// schedule of the persistent kernel (the first block drains the tile counter).  tests/test_quadtile.py feeds it leaves,
// tombstones and visits and compares every key with the oracle.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <barrier>
#include <memory>
#include <thread>
#include <vector>

// ---- CUDA language shims ----
struct float4 { float x, y, z, w; };
#define ZB_HOST_FLOAT4
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
#define __global__
#define __device__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n)
#define __shared__ static
struct Dim3 { unsigned x, y, z; };
static thread_local Dim3 emu_threadIdx;
#define threadIdx emu_threadIdx
static float* emu_dyn_smem = nullptr;
#define ZB_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(emu_dyn_smem)
static std::barrier<>* emu_block_barrier = nullptr;
#define __syncthreads() emu_block_barrier->arrive_and_wait()
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }

// quad mailbox for shuffles: 64 quads per block, one barrier of 4 threads each
struct QuadBox {
    float slot[4];
    std::barrier<> bar{4};
};
static std::vector<std::unique_ptr<QuadBox>> emu_quads;
static inline float __shfl_xor_sync(unsigned, float v, int lane_mask) {
    QuadBox& q = *emu_quads[threadIdx.x >> 2];
    const int sub = threadIdx.x & 3;
    q.slot[sub] = v;
    q.bar.arrive_and_wait();
    const float r = q.slot[sub ^ lane_mask];
    q.bar.arrive_and_wait();
    return r;
}

namespace zb {
typedef unsigned long long u64;
typedef unsigned int u32;
#define ZB_SENTINEL 0xFFFFFFFFFFFFFFFFull
struct ForestView {  // the fields quad_tile_kernel reads (names as in zb_kernels.cuh)
    const long long* leaf_off;
    const u32* leaf_len;
    const u32* members;
    const float* rows;
    const u32* tomb;
    int dimp, chunks, dim;
};
static inline unsigned quad_mask() { return 0xFu << ((threadIdx.x & 31u) & 28u); }
static inline float quad_reduce16(float4 acc, unsigned full) {  // zb_device.cuh, with the shuffles above
    acc.x = acc.x + __shfl_xor_sync(full, acc.x, 2);
    acc.y = acc.y + __shfl_xor_sync(full, acc.y, 2);
    acc.z = acc.z + __shfl_xor_sync(full, acc.z, 2);
    acc.w = acc.w + __shfl_xor_sync(full, acc.w, 2);
    acc.x = acc.x + __shfl_xor_sync(full, acc.x, 1);
    acc.y = acc.y + __shfl_xor_sync(full, acc.y, 1);
    acc.z = acc.z + __shfl_xor_sync(full, acc.z, 1);
    acc.w = acc.w + __shfl_xor_sync(full, acc.w, 1);
    return (acc.x + acc.y) + (acc.z + acc.w);
}
static inline bool tomb_test(const u32* tomb, u32 slot) { return (tomb[slot >> 5] >> (slot & 31)) & 1u; }
static inline u64 dbits(double x) { u64 u; memcpy(&u, &x, 8); return u; }
static inline u64 l2sq_bits(float s) { return dbits((double)s); }
static inline u64 l2_bits(float s) { return dbits(sqrt((double)s)); }
static inline u64 cos_bits(float ab_, float a2_, float b2_) {  // zb_device.cuh cos_bits, IEEE double operations
    const double ab = ab_, a2 = a2_, b2 = b2_;
    double c;
    if (a2 == 0.0 && b2 == 0.0) c = 0.0;
    else if (ab == 0.0) c = 1.0;
    else {
        const double ra = 1.0 / sqrt(a2), rb = 1.0 / sqrt(b2);
        const double t = (ab * ra) * rb;
        const double r = 1.0 - t;
        c = r > 0.0 ? r : 0.0;
    }
    return dbits(1.0 - c);
}
}  // namespace zb

#include "../zebra_b200/csrc/zb_quadtile_kernel.cuh"

// Tiles are given directly (what sq_count / sq_scatter / ts_* build on the device): tile t = leaf tile_leaf[t], visits
// order[tile_first[t] .. + tile_count[t]).  Runs `blocks` blocks of QT_THREADS threads.
extern "C" __attribute__((visibility("default"))) int emu_quad_tile(int metric, int blocks, int dim, const long long* leaf_off, const uint32_t* leaf_len,
                             const uint32_t* members, const float* rows_padded, const uint32_t* tomb, uint32_t ntiles,
                             const uint32_t* tile_leaf, const uint32_t* tile_first, const uint32_t* tile_count, const uint32_t* order,
                             const uint32_t* v_q, const uint64_t* v_pair_off, const float* queries_padded, uint64_t* pair_key,
                             uint64_t* stats3) {
    zb::ForestView f;
    f.leaf_off = leaf_off; f.leaf_len = leaf_len; f.members = members; f.rows = rows_padded; f.tomb = tomb;
    f.dim = dim; f.dimp = (dim + 15) / 16 * 16; f.chunks = f.dimp / 16;
    uint32_t counter = 0;
    zb::QuadTileParams tp;
    tp.tile_leaf = tile_leaf; tp.tile_first = tile_first; tp.tile_count = tile_count; tp.ntiles = &ntiles; tp.tile_counter = &counter;
    tp.order = order; tp.v_q = v_q; tp.v_pair_off = reinterpret_cast<const zb::u64*>(v_pair_off);
    tp.pair_key = reinterpret_cast<zb::u64*>(pair_key); tp.queries = queries_padded; tp.stats = reinterpret_cast<zb::u64*>(stats3);
    std::vector<float> smem((size_t)8 * f.dimp);
    emu_dyn_smem = smem.data();
    emu_quads.clear();
    for (int i = 0; i < QT_THREADS / 4; ++i) emu_quads.emplace_back(new QuadBox());
    for (int b = 0; b < blocks; ++b) {
        std::barrier<> bar(QT_THREADS);
        emu_block_barrier = &bar;
        std::vector<std::thread> th;
        for (unsigned t = 0; t < QT_THREADS; ++t)
            th.emplace_back([&, t] {
                emu_threadIdx = Dim3{t, 0, 0};
                if (metric == 0) zb::quad_tile_kernel<0>(f, tp);
                else if (metric == 1) zb::quad_tile_kernel<1>(f, tp);
                else zb::quad_tile_kernel<2>(f, tp);
            });
        for (auto& x : th) x.join();
    }
    return (int)counter;
}
