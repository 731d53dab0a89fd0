// zb_quadtile.cuh -- register tile of the KEYS-ONLY leaf-tile scan for cosine / L2 (zb_scan.cu, quad_tile_kernel): the visits
// the fused kernel does not take (n' > 32, i.e. top_k > 32: BASELINE config 5 asks for top-100) scored with the rows of a leaf
// crossing HBM once per (leaf, <= 8 queries) tile instead of once per pair.
//
// Arithmetic = Metric::distance of /root/reference/src/distance.rs:19-32, :38-49, :103-114 in the canonical "skylake-16"
// order, exactly as score_pairs_kernel states it for one pair: a QUAD owns QT_R rows x QT_Q queries,
// thread `sub` keeps lanes 4 sub .. 4 sub + 3 of every pair's 16-lane accumulator(s).
//   cosine: ab += row * query, a2 += row * row (per row), b2 += query * query (per query)     -> cos_bits(ab, a2, b2)
//   L2 / L2 squared: d = row - query, sum += d * d                                              -> l2_bits / l2sq_bits
// __host__ __device__: tests/quadtile_twin.cpp compiles qt_chunk for the CPU and checks the sums against the oracle.
#pragma once
#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define ZB_QT_HD __host__ __device__ __forceinline__
#else
#include <math.h>
#ifndef ZB_HOST_FLOAT4
#define ZB_HOST_FLOAT4
struct float4 { float x, y, z, w; };
#endif
#define ZB_QT_HD static inline
#endif

namespace zb {

#define ZB_QT_R 4  // rows per quad tile
#define ZB_QT_Q 4  // queries per quad tile

struct QtAcc {
    float4 m[ZB_QT_R][ZB_QT_Q];  // ab (cosine) or the sum of squared differences (L2)
    float4 a2[ZB_QT_R];          // cosine only
    float4 b2[ZB_QT_Q];          // cosine only
};

ZB_QT_HD float qt_fma(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a, b, c);
#else
    return fmaf(a, b, c);
#endif
}
ZB_QT_HD float qt_sub(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fsub_rn(a, b);
#else
    return a - b;
#endif
}
ZB_QT_HD void qt_zero(float4& v) { v.x = v.y = v.z = v.w = 0.0f; }
ZB_QT_HD void qt_init(QtAcc& acc) {
    for (int i = 0; i < ZB_QT_R; ++i) {
        qt_zero(acc.a2[i]);
        for (int j = 0; j < ZB_QT_Q; ++j) qt_zero(acc.m[i][j]);
    }
    for (int j = 0; j < ZB_QT_Q; ++j) qt_zero(acc.b2[j]);
}
ZB_QT_HD void qt_fma4(float4& acc, const float4& a, const float4& b) {  // zb_device.cuh fma4: acc = fma(a, b, acc)
    acc.x = qt_fma(a.x, b.x, acc.x);
    acc.y = qt_fma(a.y, b.y, acc.y);
    acc.z = qt_fma(a.z, b.z, acc.z);
    acc.w = qt_fma(a.w, b.w, acc.w);
}
ZB_QT_HD void qt_l2acc4(float4& acc, const float4& a, const float4& b) {  // zb_device.cuh l2acc4: d = a - b; acc = fma(d, d, acc)
    const float dx = qt_sub(a.x, b.x), dy = qt_sub(a.y, b.y), dz = qt_sub(a.z, b.z), dw = qt_sub(a.w, b.w);
    acc.x = qt_fma(dx, dx, acc.x);
    acc.y = qt_fma(dy, dy, acc.y);
    acc.z = qt_fma(dz, dz, acc.z);
    acc.w = qt_fma(dw, dw, acc.w);
}
// One 16-float chunk: x[i] = floats [16 c + 4 sub, +4) of row i (a = stored row), q[j] = the same floats of query j (b = query).
template <int METRIC>
ZB_QT_HD void qt_chunk(QtAcc& acc, const float4 x[ZB_QT_R], const float4 q[ZB_QT_Q]) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < ZB_QT_R; ++i) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int j = 0; j < ZB_QT_Q; ++j) {
            if (METRIC == 0) qt_fma4(acc.m[i][j], x[i], q[j]);
            else qt_l2acc4(acc.m[i][j], x[i], q[j]);
        }
        if (METRIC == 0) qt_fma4(acc.a2[i], x[i], x[i]);
    }
    if (METRIC == 0) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int j = 0; j < ZB_QT_Q; ++j) qt_fma4(acc.b2[j], q[j], q[j]);
    }
}

}  // namespace zb
