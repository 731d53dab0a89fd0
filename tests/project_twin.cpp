// project_twin.cpp -- TEST INFRASTRUCTURE.  Compiles zebra_b200/csrc/zb_project.cuh (the register tile of the flat-table
// projection kernel) for the CPU and replays what a quad of the kernel does -- thread `sub` feeds floats [16 c + 4 sub, +4)
// of 4 rows and 4 planes to pj_chunk, then the fold of quad_reduce16 (lane i + lane i + 8, + 4, (r0 + r1) + (r2 + r3)) --
// so tests/test_flat_tables.py can check the accumulation against the oracle's dot product without a GPU.
// Built by the test with g++ -O2 -ffp-contract=off.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../zebra_b200/csrc/zb_project.cuh"

// dots[n][H] (f32): rows [n][dim], planes [H][dim]; tails handled like the kernel (indices clamped, results masked)
extern "C" void twin_project(const float* rows, uint64_t n, const float* planes, int H, int dim, float* dots) {
    const int dimp = (dim + 15) / 16 * 16, chunks = dimp / 16;
    std::vector<float> R((size_t)n * dimp, 0.0f), P((size_t)H * dimp, 0.0f);
    for (uint64_t i = 0; i < n; ++i) memcpy(&R[i * dimp], rows + i * dim, sizeof(float) * dim);
    for (int j = 0; j < H; ++j) memcpy(&P[(size_t)j * dimp], planes + (size_t)j * dim, sizeof(float) * dim);
    for (uint64_t row0 = 0; row0 < n; row0 += ZB_PJ_R)
        for (int pl0 = 0; pl0 < H; pl0 += ZB_PJ_P) {
            zb::PjAcc acc[4];
            for (int sub = 0; sub < 4; ++sub) {
                zb::pj_init(acc[sub]);
                for (int c = 0; c < chunks; ++c) {
                    float4 x[ZB_PJ_R], p[ZB_PJ_P];
                    for (int i = 0; i < ZB_PJ_R; ++i) {
                        const uint64_t r = row0 + i < n ? row0 + i : n - 1;
                        memcpy(&x[i], &R[r * dimp + 16 * c + 4 * sub], 16);
                    }
                    for (int j = 0; j < ZB_PJ_P; ++j) {
                        const int h = pl0 + j < H ? pl0 + j : H - 1;
                        memcpy(&p[j], &P[(size_t)h * dimp + 16 * c + 4 * sub], 16);
                    }
                    zb::pj_chunk(acc[sub], x, p);
                }
            }
            for (int i = 0; i < ZB_PJ_R; ++i)
                for (int j = 0; j < ZB_PJ_P; ++j) {
                    if (row0 + i >= n || pl0 + j >= H) continue;
                    float lane[16];
                    for (int sub = 0; sub < 4; ++sub) {
                        lane[4 * sub + 0] = acc[sub].a[i][j].x; lane[4 * sub + 1] = acc[sub].a[i][j].y;
                        lane[4 * sub + 2] = acc[sub].a[i][j].z; lane[4 * sub + 3] = acc[sub].a[i][j].w;
                    }
                    float x8[8], r4[4];
                    for (int k = 0; k < 8; ++k) x8[k] = lane[k] + lane[k + 8];   // __shfl_xor 2: thread sub with sub ^ 2
                    for (int k = 0; k < 4; ++k) r4[k] = x8[k] + x8[k + 4];       // __shfl_xor 1
                    dots[(row0 + i) * (uint64_t)H + pl0 + j] = (r4[0] + r4[1]) + (r4[2] + r4[3]);
                }
        }
}
extern "C" uint64_t twin_key_from_ballots(uint32_t b0, uint32_t b1, int K) { return zb::pj_key_from_ballots(b0, b1, K); }
