// zb_quadtile_kernel.cuh -- quad_tile_kernel: the keys-only leaf-tile scan for cosine / L2 visits outside the fused kernel
// (n' > 32; see zb_quadtile.cuh for the arithmetic and zb_scan.cu for the host side).  In a header of its own so that
// tests/quadtile_emu.cpp can compile THIS SOURCE for the CPU (CUDA built-ins shimmed, one std::thread per CUDA thread) and
// run whole tiles against the oracle without a GPU.  Needs from its includer: ForestView, u32 / u64 / Entry sentinels,
// quad_mask, quad_reduce16, tomb_test, cos_bits, l2sq_bits, l2_bits (zb_device.cuh on the device).
#pragma once
#include "zb_quadtile.cuh"

#ifndef ZB_DYN_SMEM
#define ZB_DYN_SMEM(type, name) extern __shared__ __align__(16) type name[]
#endif

namespace zb {

#define QT_THREADS 256

struct QuadTileParams {
    const u32* tile_leaf;
    const u32* tile_first;
    const u32* tile_count;
    const u32* ntiles;
    u32* tile_counter;
    const u32* order;
    const u32* v_q;
    const u64* v_pair_off;
    u64* pair_key;
    const float* queries;
    u64* stats;  // [0] visits, [1] pairs, [2] bytes asked of HBM by design
};

template <int METRIC>
__global__ void __launch_bounds__(QT_THREADS) quad_tile_kernel(ForestView f, QuadTileParams tp) {
    ZB_DYN_SMEM(float, s_q);  // [8][dimp]
    __shared__ u32 s_tile;
    __shared__ u32 s_v[8];
    __shared__ u64 s_pbase[8];
    const u32 ntiles = *tp.ntiles;
    const int q4 = f.dimp >> 2;
    const int quad = threadIdx.x >> 2, sub = threadIdx.x & 3;
    const unsigned mask = quad_mask();
    for (;;) {
        __syncthreads();  // the previous tile's queries are no longer read
        if (threadIdx.x == 0) s_tile = atomicAdd(tp.tile_counter, 1u);
        __syncthreads();
        const u32 tile = s_tile;
        if (tile >= ntiles) break;
        const u32 leaf = tp.tile_leaf[tile], first = tp.tile_first[tile], c = tp.tile_count[tile];
        if (threadIdx.x < 8) {
            const u32 v = tp.order[first + (threadIdx.x < c ? threadIdx.x : 0u)];
            s_v[threadIdx.x] = v;
            s_pbase[threadIdx.x] = tp.v_pair_off[v];
        }
        __syncthreads();
        const u32 nqp = c <= 4 ? 4u : 8u;  // query slots: one or two groups of ZB_QT_Q
        for (u32 idx = threadIdx.x; idx < nqp * (u32)q4; idx += QT_THREADS) {
            const u32 j = idx / (u32)q4, k = idx - j * (u32)q4;
            reinterpret_cast<float4*>(s_q)[idx] =
                j < c ? __ldg(reinterpret_cast<const float4*>(tp.queries + (size_t)tp.v_q[s_v[j]] * f.dimp) + k)
                      : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncthreads();
        const u32 len = f.leaf_len[leaf];
        const long long off = f.leaf_off[leaf];
        // 64 quads: with one query group every quad takes its own 4 rows (256 rows per pass), with two groups the quads of a
        // pair share 4 rows and split the queries (128 rows per pass)
        const u32 qg = nqp == 8 ? (u32)(quad & 1) : 0u;
        const u32 rg = nqp == 8 ? (u32)(quad >> 1) : (u32)quad;
        const u32 rows_per_pass = nqp == 8 ? 128u : 256u;
        const float4* qp = reinterpret_cast<const float4*>(s_q) + (size_t)qg * ZB_QT_Q * q4 + sub;
        for (u32 base = 0; base < len; base += rows_per_pass) {
            const u32 r0 = base + rg * ZB_QT_R;
            if (r0 >= len) continue;  // the whole quad skips together
            const float4* xr[ZB_QT_R];
            bool dead[ZB_QT_R];
#pragma unroll
            for (int i = 0; i < ZB_QT_R; ++i) {  // tail rows: clamp the loads, mask the stores
                const u32 r = r0 + i < len ? r0 + i : len - 1;
                const u32 slot = f.members[off + r];
                dead[i] = tomb_test(f.tomb, slot);
                xr[i] = reinterpret_cast<const float4*>(f.rows + (size_t)slot * f.dimp) + sub;
            }
            QtAcc acc;
            qt_init(acc);
#pragma unroll 2
            for (int ch = 0; ch < f.chunks; ++ch) {
                float4 x[ZB_QT_R], q[ZB_QT_Q];
#pragma unroll
                for (int i = 0; i < ZB_QT_R; ++i) x[i] = __ldg(xr[i] + ch * 4);
#pragma unroll
                for (int j = 0; j < ZB_QT_Q; ++j) q[j] = qp[(size_t)j * q4 + ch * 4];
                qt_chunk<METRIC>(acc, x, q);
            }
            float a2[ZB_QT_R], b2[ZB_QT_Q];
            if (METRIC == 0) {
#pragma unroll
                for (int i = 0; i < ZB_QT_R; ++i) a2[i] = quad_reduce16(acc.a2[i], mask);
#pragma unroll
                for (int j = 0; j < ZB_QT_Q; ++j) b2[j] = quad_reduce16(acc.b2[j], mask);
            }
#pragma unroll
            for (int i = 0; i < ZB_QT_R; ++i) {
#pragma unroll
                for (int j = 0; j < ZB_QT_Q; ++j) {
                    const float sum = quad_reduce16(acc.m[i][j], mask);  // every thread of the quad takes part
                    const u32 jq = qg * ZB_QT_Q + j;
                    if (sub == 0 && r0 + i < len && jq < c) {
                        u64 key;
                        if (dead[i]) key = ZB_SENTINEL;
                        else if (METRIC == 0) key = cos_bits(sum, a2[i], b2[j]);
                        else key = METRIC == 1 ? l2sq_bits(sum) : l2_bits(sum);
                        tp.pair_key[s_pbase[jq] + r0 + i] = key;
                    }
                }
            }
        }
        if (threadIdx.x == 0) {
            atomicAdd(&tp.stats[0], (u64)c);
            atomicAdd(&tp.stats[1], (u64)len * c);
            atomicAdd(&tp.stats[2], ((u64)len + c) * 4ull * (u64)f.dim);
        }
    }
}

}  // namespace zb
