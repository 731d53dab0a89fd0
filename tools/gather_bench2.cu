// gather_bench2.cu -- (1) does TMA small-copy throughput scale with the number of issuing warps?
// (2) cp.async (LDGSTS, 16 B per thread) staging of K-sliced gathered rows, 256 threads per CTA.
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
// TMA, W issuing warps per CTA
__global__ void k_tma(const float* base, const u32* ids, u32 nids, u32 row_floats, u32 bytes, u32 rows, u32 nst, u32 stages_total, u32 slices) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned long long* bars = (unsigned long long*)smem;
    unsigned char* buf = smem + 128;
    const int tid = threadIdx.x, nthr = blockDim.x;
    if (tid == 0) { for (u32 i = 0; i < nst; ++i) mbar_init(smem_u32(bars + i), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    const u32 stage_bytes = rows * bytes;
    u32 idbase = blockIdx.x * 7919u;
    for (u32 st = 0; st < stages_total + nst; ++st) {
        if (st >= nst) mbar_wait(smem_u32(bars + (st % nst)), ((st / nst) - 1) & 1);
        __syncthreads();
        if (st < stages_total) {
            const u32 b = st % nst;
            if (tid == 0) mbar_arrive_expect_tx(smem_u32(bars + b), stage_bytes);
            __syncthreads();
            const u32 blk = st / slices, sl = st % slices;
            for (u32 r = tid; r < rows; r += nthr) {
                const u32 id = ids[(idbase + blk * rows + r) % nids];
                bulk_g2s(smem_u32(buf + (size_t)b * stage_bytes + (size_t)r * bytes), base + (size_t)id * row_floats + sl * (bytes / 4), bytes, smem_u32(bars + b));
            }
        }
    }
}
// cp.async 16 B per thread, 256 threads, stage = rows x bytes; wait_group based ring
__global__ void k_ldgsts(const float* base, const u32* ids, u32 nids, u32 row_floats, u32 bytes, u32 rows, u32 nst, u32 stages_total, u32 slices) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* buf = smem;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const u32 stage_bytes = rows * bytes;
    const u32 per_row = bytes / 16, total16 = rows * per_row;
    u32 idbase = blockIdx.x * 7919u;
    for (u32 st = 0; st < stages_total + nst - 1; ++st) {
        if (st < stages_total) {
            const u32 b = st % nst;
            const u32 blk = st / slices, sl = st % slices;
            for (u32 i = tid; i < total16; i += nthr) {
                const u32 r = i / per_row, part = i - r * per_row;
                const u32 id = ids[(idbase + blk * rows + r) % nids];
                const float* src = base + (size_t)id * row_floats + sl * (bytes / 4) + part * 4;
                u32 dst = smem_u32(buf + (size_t)b * stage_bytes + (size_t)r * bytes + part * 16);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (st + 1 >= nst) { asm volatile("cp.async.wait_group %0;" ::"n"(2) : "memory"); }
        __syncthreads();
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}
int main() {
    const size_t nrows = 1000000, row_floats = 768;
    float* d; cudaMalloc(&d, nrows * row_floats * 4); cudaMemset(d, 0, nrows * row_floats * 4);
    std::vector<u32> ids(1 << 20); for (auto& x : ids) x = (u32)(((unsigned long long)rand() * 48271ull) % nrows);
    u32* dids; cudaMalloc(&dids, ids.size() * 4); cudaMemcpy(dids, ids.data(), ids.size() * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaFuncSetAttribute(k_ldgsts, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    struct Cfg { int kind; u32 bytes, rows, nst, threads; } cfgs[] = {
        {0, 384, 128, 3, 32}, {0, 384, 128, 3, 128}, {0, 384, 128, 3, 256}, {0, 768, 64, 3, 128}, {0, 3072, 16, 4, 32}, {0, 3072, 16, 4, 128}, {0, 3072, 32, 2, 64},
        {1, 384, 128, 3, 256}, {1, 768, 64, 3, 256}, {1, 1536, 32, 3, 256}, {1, 3072, 16, 3, 256}, {1, 384, 128, 3, 512}, {1, 192, 256, 3, 256}, {1, 128, 512, 3, 256}, {0, 192, 256, 3, 256}, {0, 128, 512, 3, 256}, {1, 384, 128, 2, 256}, {1, 384, 128, 4, 256}, {1, 384, 64, 6, 256}};
    for (auto c : cfgs) {
        const u32 slices = 3072 / c.bytes, stages_total = 4000;
        size_t smem = 128 + (size_t)c.nst * c.rows * c.bytes;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        auto run = [&](u32 n) { if (c.kind == 0) k_tma<<<148, c.threads, smem>>>(d, dids, (u32)ids.size(), row_floats, c.bytes, c.rows, c.nst, n, slices);
                                else k_ldgsts<<<148, c.threads, smem>>>(d, dids, (u32)ids.size(), row_floats, c.bytes, c.rows, c.nst, n, slices); };
        run(200); cudaDeviceSynchronize();
        cudaEventRecord(e0); run(stages_total); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double gb = 148.0 * stages_total * c.rows * c.bytes / 1e9;
        printf("%s copy %4u B x %3u rows/stage, %u stages, %3u threads: %7.1f GB/s (%.2f us/stage/SM) err=%s\n", c.kind ? "LDGSTS" : "TMA   ", c.bytes, c.rows, c.nst, c.threads,
               gb / (ms / 1e3), ms * 1e3 / stages_total, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
