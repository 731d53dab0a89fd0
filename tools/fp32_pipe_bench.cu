// fp32_pipe_bench.cu -- FFMA vs FFMA2 (fma.rn.f32x2) issue throughput on sm_100a, as a function of warps per SM sub-partition.
#include <cuda_runtime.h>
#include <cstdio>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
template <int MODE>  // 0: FFMA chains, 1: FFMA2 chains, 2: FADD->FFMA pairs (L2 pattern), 3: FADD2->FFMA2 pairs
__global__ void k(float* out, int iters, float seed) {
    float a[32]; u64 p[16];
#pragma unroll
    for (int i = 0; i < 32; ++i) a[i] = seed + i + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 16; ++i) p[i] = pack(seed + i, seed - i);
    float x = seed * 1.0001f, y = seed * 0.9999f;
    u64 xp = pack(x, y), yp = pack(y, x);
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 32; ++i) a[i] = __fmaf_rn(a[i], x, y);
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(xp), "l"(yp));
        } else if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 32; ++i) { float d = __fsub_rn(x, a[(i + 1) & 31] * 0.f + y + i); a[i] = __fmaf_rn(d, d, a[i]); }
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) { u64 d; asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(xp), "l"(p[(i + 1) & 15]));
                asm volatile("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(p[i]) : "l"(d)); }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += a[i];
#pragma unroll
    for (int i = 0; i < 16; ++i) { float lo, hi; asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[i])); s += lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(int threads, const char* name) {
    float* d; cudaMalloc(&d, 148 * 1024 * 4);
    const int iters = 20000;
    k<MODE><<<148, threads>>>(d, 100, 1.0f); cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<MODE><<<148, threads>>>(d, iters, 1.0f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    // scalar-lane operations per iteration per thread: mode 0: 32 fma; 1: 32 fma (16 x2); 2: 32 sub + 32 fma; 3: 32 sub + 32 fma
    double lane_ops = (double)148 * threads * iters * (MODE < 2 ? 32.0 : 64.0);
    printf("%-28s %4d thr/SM (%d warps/SMSP): %7.2f T lane-ops/s  (%.3f ms)\n", name, threads, threads / 128, lane_ops / ms / 1e9, ms);
    cudaFree(d);
}
int main() {
    for (int t : {128, 256, 512, 1024}) {
        run<0>(t, "FFMA  independent chains");
        run<1>(t, "FFMA2 independent chains");
        run<2>(t, "FADD->FFMA pairs");
        run<3>(t, "FADD2->FFMA2 pairs");
    }
    return 0;
}
