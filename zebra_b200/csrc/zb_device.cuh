// zb_device.cuh -- device-side building blocks shared by the kernels of libzebra_b200.
//
// Canonical arithmetic ("skylake-16"): the reference's dot / l2sq / cos run in simsimd's
// AVX-512 f32 kernels (call sites /root/reference/src/database/index/lsh.rs:40,:224 and
// /root/reference/src/distance.rs:23,:41,:106): lane j of a 16-lane f32 accumulator receives elements
// j, j+16, ... by one fused multiply-add each, then x[i]=acc[i]+acc[i+8], r[i]=x[i]+x[i+4],
// s=(r0+r1)+(r2+r3).  On the GPU one dot product is owned by a QUAD of 4 threads: thread `sub` keeps lanes
// 4*sub..4*sub+3 as a float4 (one 128-bit load per 16-element chunk), the two cross-thread folds are
// __shfl_xor 2 then 1, and the final two adds happen inside the thread.  No reassociation, explicit
// __fmaf_rn / __fadd_rn / __fsub_rn everywhere, so results are bit-identical to the CPU order.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace zb {

typedef unsigned long long u64;
typedef unsigned int u32;
typedef unsigned char u8;

struct Entry {
    u64 key;  // f64 bit pattern of the distance (DistanceUnit, distance.rs:13)
    u64 ord;  // global row ordinal (tie-break, D3)
};
#define ZB_SENTINEL 0xFFFFFFFFFFFFFFFFull

__device__ __forceinline__ bool entry_less(const Entry& a, const Entry& b) {
    return a.key < b.key || (a.key == b.key && a.ord < b.ord);
}

// ---- deterministic sampling spec: node keys and the min-hash that picks the two sample rows ----
__host__ __device__ __forceinline__ u64 mix64(u64 x) {
    x += 0x9E3779B97F4A7C15ull;
    u64 z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ u64 root_key(u64 seed, int tree) { return mix64(seed ^ mix64((u64)tree)); }
__host__ __device__ __forceinline__ u64 child_key(u64 key, int side) {
    return mix64(key ^ (side ? 0xA5A5A5A5A5A5A5A5ull : 0x5A5A5A5A5A5A5A5Aull));
}
__host__ __device__ __forceinline__ u64 pick_hash(u64 key, int attempt, u64 ordinal) {
    return mix64(key ^ mix64(ordinal ^ ((u64)attempt << 56)));
}

// ---- quad reduction: every thread of the quad ends with the canonical 16-lane sum ----
// `full` must name (at least) the calling quad; a quad mask (0xF << (lane & 28)) is safe in divergent code.
__device__ __forceinline__ unsigned quad_mask() { return 0xFu << ((threadIdx.x & 31u) & 28u); }
__device__ __forceinline__ float quad_reduce16(float4 acc, unsigned full) {
    acc.x = __fadd_rn(acc.x, __shfl_xor_sync(full, acc.x, 2));  // lane i + lane i+8
    acc.y = __fadd_rn(acc.y, __shfl_xor_sync(full, acc.y, 2));
    acc.z = __fadd_rn(acc.z, __shfl_xor_sync(full, acc.z, 2));
    acc.w = __fadd_rn(acc.w, __shfl_xor_sync(full, acc.w, 2));
    acc.x = __fadd_rn(acc.x, __shfl_xor_sync(full, acc.x, 1));  // x[i] + x[i+4]
    acc.y = __fadd_rn(acc.y, __shfl_xor_sync(full, acc.y, 1));
    acc.z = __fadd_rn(acc.z, __shfl_xor_sync(full, acc.z, 1));
    acc.w = __fadd_rn(acc.w, __shfl_xor_sync(full, acc.w, 1));
    return __fadd_rn(__fadd_rn(acc.x, acc.y), __fadd_rn(acc.z, acc.w));  // (r0+r1)+(r2+r3)
}

__device__ __forceinline__ void fma4(float4& acc, const float4& a, const float4& b) {
    acc.x = __fmaf_rn(a.x, b.x, acc.x);
    acc.y = __fmaf_rn(a.y, b.y, acc.y);
    acc.z = __fmaf_rn(a.z, b.z, acc.z);
    acc.w = __fmaf_rn(a.w, b.w, acc.w);
}
__device__ __forceinline__ void l2acc4(float4& acc, const float4& a, const float4& b) {
    float dx = __fsub_rn(a.x, b.x), dy = __fsub_rn(a.y, b.y), dz = __fsub_rn(a.z, b.z), dw = __fsub_rn(a.w, b.w);
    acc.x = __fmaf_rn(dx, dx, acc.x);
    acc.y = __fmaf_rn(dy, dy, acc.y);
    acc.z = __fmaf_rn(dz, dz, acc.z);
    acc.w = __fmaf_rn(dw, dw, acc.w);
}

// Quad dot product of two vectors of `chunks` 16-float chunks in global memory (16-byte aligned).
__device__ __forceinline__ float quad_dot(const float4* __restrict__ a, const float4* __restrict__ b, int chunks,
                                          int sub, unsigned mask) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int c = 0;
    for (; c + 4 <= chunks; c += 4) {
        float4 a0 = __ldg(a + (c + 0) * 4 + sub), a1 = __ldg(a + (c + 1) * 4 + sub);
        float4 a2 = __ldg(a + (c + 2) * 4 + sub), a3 = __ldg(a + (c + 3) * 4 + sub);
        float4 b0 = __ldg(b + (c + 0) * 4 + sub), b1 = __ldg(b + (c + 1) * 4 + sub);
        float4 b2 = __ldg(b + (c + 2) * 4 + sub), b3 = __ldg(b + (c + 3) * 4 + sub);
        fma4(acc, a0, b0);
        fma4(acc, a1, b1);
        fma4(acc, a2, b2);
        fma4(acc, a3, b3);
    }
    for (; c < chunks; ++c) fma4(acc, __ldg(a + c * 4 + sub), __ldg(b + c * 4 + sub));
    return quad_reduce16(acc, mask);
}

// Hyperplane::point_is_above, lsh.rs:39-43: (f64)dot + (f64)constant >= 0.0
__device__ __forceinline__ bool above_from_dot(float dot, float constant) {
    return __dadd_rn((double)dot, (double)constant) >= 0.0;
}

// Metric epilogues (distance.rs:19-32, :38-49, :103-114) -> DistanceUnit bits.
__device__ __forceinline__ u64 l2sq_bits(float s) { return (u64)__double_as_longlong((double)s); }
__device__ __forceinline__ u64 l2_bits(float s) { return (u64)__double_as_longlong(__dsqrt_rn((double)s)); }
__device__ __forceinline__ u64 cos_bits(float ab_, float a2_, float b2_) {
    double ab = (double)ab_, a2 = (double)a2_, b2 = (double)b2_;
    double c;
    if (a2 == 0.0 && b2 == 0.0) c = 0.0;
    else if (ab == 0.0) c = 1.0;
    else {
        double ra = __ddiv_rn(1.0, __dsqrt_rn(a2));
        double rb = __ddiv_rn(1.0, __dsqrt_rn(b2));
        double t = __dmul_rn(__dmul_rn(ab, ra), rb);
        double r = __dsub_rn(1.0, t);
        c = r > 0.0 ? r : 0.0;
    }
    return (u64)__double_as_longlong(__dsub_rn(1.0, c));  // Q4: 1.0 - cosine distance
}

// ---- PTX helpers: mbarrier and the 1-D bulk copy engine (TMA, SASS UBLKCP) ----
__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u32 bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(u32 bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(u32 bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u32 bar, u32 parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity), "r"(0x989680u)   // suspend-time hint: the warp sleeps in hardware instead of spinning
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(u32 dst, const void* src, u32 bytes, u32 bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

__device__ __forceinline__ bool tomb_test(const u32* __restrict__ tomb, u32 slot) {
    return (tomb[slot >> 5] >> (slot & 31)) & 1u;
}

}  // namespace zb
