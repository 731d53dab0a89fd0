// tma_gather_bench.cu -- microbenchmark behind the stage-layout choice of zb_scan.cu:
// gathered row slices -> shared memory with cp.async.bulk (one copy per row slice), per-SM ring of stages,
// no compute.  Reports GB/s against copy size / rows per stage / issuing lanes.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
typedef unsigned int u32;
__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u32 bar, u32 count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_arrive_expect_tx(u32 bar, u32 bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(u32 bar, u32 parity) {
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(u32 dst, const void* src, u32 bytes, u32 bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// one warp per CTA: issues `rows` copies of `bytes` per stage, waits for the stage `depth` stages back.
__global__ void k(const float* base, const u32* ids, u32 nids, u32 row_floats, u32 bytes, u32 rows, u32 nst, u32 stages_total, u32 slices) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned long long* bars = (unsigned long long*)smem;
    unsigned char* buf = smem + 128;
    const int lane = threadIdx.x;
    if (lane == 0) { for (u32 i = 0; i < nst; ++i) mbar_init(smem_u32(bars + i), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncwarp();
    const u32 stage_bytes = rows * bytes;
    u32 idbase = blockIdx.x * 7919u;
    for (u32 st = 0; st < stages_total + nst; ++st) {
        if (st >= nst) mbar_wait(smem_u32(bars + (st % nst)), ((st / nst) - 1) & 1);   // stage st - nst landed
        if (st < stages_total) {
            const u32 b = st % nst;
            if (lane == 0) mbar_arrive_expect_tx(smem_u32(bars + b), stage_bytes);
            __syncwarp();
            const u32 blk = st / slices, sl = st % slices;
            for (u32 r = lane; r < rows; r += 32) {
                const u32 id = ids[(idbase + blk * rows + r) % nids];
                bulk_g2s(smem_u32(buf + (size_t)b * stage_bytes + (size_t)r * bytes), base + (size_t)id * row_floats + sl * (bytes / 4), bytes, smem_u32(bars + b));
            }
        }
    }
}
int main() {
    const size_t nrows = 1000000, row_floats = 768;
    float* d; cudaMalloc(&d, nrows * row_floats * 4); cudaMemset(d, 0, nrows * row_floats * 4);
    std::vector<u32> ids(1 << 20); for (auto& x : ids) x = (u32)(((unsigned long long)rand() * 48271ull) % nrows);
    u32* dids; cudaMalloc(&dids, ids.size() * 4); cudaMemcpy(dids, ids.data(), ids.size() * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    struct Cfg { u32 bytes, rows, nst; } cfgs[] = {{384, 128, 3}, {384, 128, 4}, {768, 64, 3}, {768, 64, 4}, {1536, 32, 4}, {3072, 16, 3}, {3072, 16, 4}, {3072, 16, 6}, {3072, 32, 2}, {3072, 8, 8}, {1536, 64, 2}, {768, 128, 2}};
    for (auto c : cfgs) {
        const u32 slices = 3072 / c.bytes;
        const u32 stages_total = 4000;
        size_t smem = 128 + (size_t)c.nst * c.rows * c.bytes;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        k<<<148, 32, smem>>>(d, dids, (u32)ids.size(), row_floats, c.bytes, c.rows, c.nst, 200, slices);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        k<<<148, 32, smem>>>(d, dids, (u32)ids.size(), row_floats, c.bytes, c.rows, c.nst, stages_total, slices);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double gb = 148.0 * stages_total * c.rows * c.bytes / 1e9;
        printf("copy %4u B x %3u rows/stage, %u stages (%3zu KB smem): %7.1f GB/s  (%.2f us/stage/SM, %.0f ns/copy)  err=%s\n", c.bytes, c.rows, c.nst, smem / 1024,
               gb / (ms / 1e3), ms * 1e3 / stages_total, ms * 1e6 / stages_total / c.rows, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
