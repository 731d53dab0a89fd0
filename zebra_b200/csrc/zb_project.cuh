// zb_project.cuh -- the register tile of the flat-table projection (north_star (a): "LSH hyperplane projection, where
// per-thread f32 FMA dot products reproduce the reference's accumulation order so sign bits ... are bit-exact").
//
// A FLAT table of K bits is the special case of the reference's tree (/root/reference/src/database/index/lsh.rs:46-60) in
// which every node at depth d of a tree shares hyperplane d: the root-to-leaf walk of lsh.rs:350-366
// asks the same K planes of every row, so hashing collapses into one dense [rows x N] . [N x K T] pass.  Each of the
// rows x planes dot products is still Hyperplane::point_is_above (lsh.rs:39-43) in the canonical "skylake-16" order
// (DESIGN.md section 4): a QUAD owns a tile of PJ_R rows x PJ_P planes, thread `sub` keeps lanes 4 sub .. 4 sub + 3 of every
// pair's 16-lane accumulator, so each 128-bit load of a row (plane) chunk feeds PJ_P (PJ_R) fused multiply-adds and
// the fold is the same quad_reduce16 as everywhere else.  Tensor cores are deliberately unused (bit-exact f32 FMA order).
//
// __host__ __device__: tests/project_twin.cpp compiles pj_chunk for the CPU and checks the accumulation against the oracle.
#pragma once
#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define ZB_PJ_HD __host__ __device__ __forceinline__
#else
#include <math.h>
#ifndef ZB_HOST_FLOAT4
#define ZB_HOST_FLOAT4
struct float4 { float x, y, z, w; };
#endif
#define ZB_PJ_HD static inline
#endif

namespace zb {

#define ZB_PJ_R 4  // rows per quad tile
#define ZB_PJ_P 4  // planes per quad tile

struct PjAcc {
    float4 a[ZB_PJ_R][ZB_PJ_P];
};

ZB_PJ_HD float pj_fma(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a, b, c);
#else
    return fmaf(a, b, c);
#endif
}
ZB_PJ_HD void pj_init(PjAcc& acc) {
    for (int i = 0; i < ZB_PJ_R; ++i)
        for (int j = 0; j < ZB_PJ_P; ++j) acc.a[i][j].x = acc.a[i][j].y = acc.a[i][j].z = acc.a[i][j].w = 0.0f;
}
// One 16-float chunk: x[i] = floats [16 c + 4 sub, +4) of row i, p[j] = the same floats of plane j.
// acc lane = fma(plane, row, acc), the operand order of the tree walk's quad_dot(plane, row).
ZB_PJ_HD void pj_chunk(PjAcc& acc, const float4 x[ZB_PJ_R], const float4 p[ZB_PJ_P]) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < ZB_PJ_R; ++i) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int j = 0; j < ZB_PJ_P; ++j) {
            float4& a = acc.a[i][j];
            a.x = pj_fma(p[j].x, x[i].x, a.x);
            a.y = pj_fma(p[j].y, x[i].y, a.y);
            a.z = pj_fma(p[j].z, x[i].z, a.z);
            a.w = pj_fma(p[j].w, x[i].w, a.w);
        }
    }
}

// Bucket key of one table from its K sign bits, MSB = plane 0 (the root decision of the equivalent tree, lsh.rs:358-363).
// b0 = ballot over planes 0..31 (bit l = plane l), b1 = ballot over planes 32..63.
ZB_PJ_HD unsigned long long pj_key_from_ballots(unsigned b0, unsigned b1, int K) {
    unsigned r0 = 0, r1 = 0;  // bit reversal: plane 0 -> bit 31
    for (int i = 0; i < 32; ++i) {
        r0 |= ((b0 >> i) & 1u) << (31 - i);
        r1 |= ((b1 >> i) & 1u) << (31 - i);
    }
    if (K <= 32) return (unsigned long long)(r0 >> (32 - K));
    return ((unsigned long long)r0 << (K - 32)) | (unsigned long long)(r1 >> (64 - K));
}

}  // namespace zb
