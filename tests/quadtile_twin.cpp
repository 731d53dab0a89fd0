// quadtile_twin.cpp -- TEST INFRASTRUCTURE.  Compiles zebra_b200/csrc/zb_quadtile.cuh (the register tile of the keys-only
// leaf-tile scan for cosine / L2) for the CPU and replays a quad of quad_tile_kernel: thread `sub` feeds floats
// [16 c + 4 sub, +4) of 4 rows and 4 queries to qt_chunk, then the quad_reduce16 fold -- so tests/test_quadtile.py can check
// the sums against the oracle without a GPU.  Built by the test with g++ -O2 -ffp-contract=off.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../zebra_b200/csrc/zb_quadtile.cuh"

static float fold(const float4 acc[4]) {  // acc[sub] = lanes 4 sub .. 4 sub + 3
    float lane[16], x8[8], r4[4];
    for (int sub = 0; sub < 4; ++sub) {
        lane[4 * sub + 0] = acc[sub].x; lane[4 * sub + 1] = acc[sub].y; lane[4 * sub + 2] = acc[sub].z; lane[4 * sub + 3] = acc[sub].w;
    }
    for (int k = 0; k < 8; ++k) x8[k] = lane[k] + lane[k + 8];
    for (int k = 0; k < 4; ++k) r4[k] = x8[k] + x8[k + 4];
    return (r4[0] + r4[1]) + (r4[2] + r4[3]);
}

template <int METRIC>
static void run(const float* rows, uint64_t n, const float* queries, int m, int dim, float* out_m, float* out_a2, float* out_b2) {
    const int dimp = (dim + 15) / 16 * 16, chunks = dimp / 16;
    std::vector<float> R((size_t)n * dimp, 0.0f), Q((size_t)m * dimp, 0.0f);
    for (uint64_t i = 0; i < n; ++i) memcpy(&R[i * dimp], rows + i * dim, sizeof(float) * dim);
    for (int j = 0; j < m; ++j) memcpy(&Q[(size_t)j * dimp], queries + (size_t)j * dim, sizeof(float) * dim);
    for (uint64_t r0 = 0; r0 < n; r0 += ZB_QT_R)
        for (int q0 = 0; q0 < m; q0 += ZB_QT_Q) {
            zb::QtAcc acc[4];
            for (int sub = 0; sub < 4; ++sub) {
                zb::qt_init(acc[sub]);
                for (int c = 0; c < chunks; ++c) {
                    float4 x[ZB_QT_R], q[ZB_QT_Q];
                    for (int i = 0; i < ZB_QT_R; ++i) {
                        const uint64_t r = r0 + i < n ? r0 + i : n - 1;   // the kernel clamps tail rows
                        memcpy(&x[i], &R[r * dimp + 16 * c + 4 * sub], 16);
                    }
                    for (int j = 0; j < ZB_QT_Q; ++j) {
                        float4 z = {0.f, 0.f, 0.f, 0.f};                  // ... and zero-fills query slots beyond the tile
                        if (q0 + j < m) memcpy(&z, &Q[(size_t)(q0 + j) * dimp + 16 * c + 4 * sub], 16);
                        q[j] = z;
                    }
                    zb::qt_chunk<METRIC>(acc[sub], x, q);
                }
            }
            for (int i = 0; i < ZB_QT_R; ++i)
                for (int j = 0; j < ZB_QT_Q; ++j) {
                    if (r0 + i >= n || q0 + j >= m) continue;
                    const float4 a[4] = {acc[0].m[i][j], acc[1].m[i][j], acc[2].m[i][j], acc[3].m[i][j]};
                    out_m[(r0 + i) * (uint64_t)m + q0 + j] = fold(a);
                    if (METRIC == 0) {
                        const float4 s[4] = {acc[0].a2[i], acc[1].a2[i], acc[2].a2[i], acc[3].a2[i]};
                        const float4 t[4] = {acc[0].b2[j], acc[1].b2[j], acc[2].b2[j], acc[3].b2[j]};
                        out_a2[r0 + i] = fold(s);
                        out_b2[q0 + j] = fold(t);
                    }
                }
        }
}

extern "C" void twin_quadtile(int metric, const float* rows, uint64_t n, const float* queries, int m, int dim, float* out_m,
                              float* out_a2, float* out_b2) {
    if (metric == 0) run<0>(rows, n, queries, m, dim, out_m, out_a2, out_b2);
    else run<1>(rows, n, queries, m, dim, out_m, out_a2, out_b2);
}
