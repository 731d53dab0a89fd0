// metric_twin.cpp -- TEST INFRASTRUCTURE.  Compiles zebra_b200/csrc/zb_metrics.cuh (the __host__ __device__ arithmetic
// the scalar-metric kernels execute) for the CPU, so tests/test_metric_twin.py can check it against the oracle without
// a GPU.  Built by the test with g++ -O2 -ffp-contract=off (no implicit FMA: the device side uses explicitly rounded
// intrinsics).  Never part of the product library.
#include <stdint.h>

#include "../zebra_b200/csrc/zb_metrics.cuh"

template <int C>
static void run(uint64_t n, const float* a, const float* b, int dim, int power, uint64_t* out) {
    for (uint64_t i = 0; i < n; ++i) out[i] = zb::seq_distance<C>(a + i * (uint64_t)dim, b + i * (uint64_t)dim, dim, power);
}

extern "C" int twin_distance_bits_batch(int code, int power, uint64_t n, const float* a, const float* b, int dim,
                                        uint64_t* out) {
    switch (code) {
        case zb::M_CHEBYSHEV: run<zb::M_CHEBYSHEV>(n, a, b, dim, power, out); break;
        case zb::M_CANBERRA: run<zb::M_CANBERRA>(n, a, b, dim, power, out); break;
        case zb::M_BRAY_CURTIS: run<zb::M_BRAY_CURTIS>(n, a, b, dim, power, out); break;
        case zb::M_MANHATTAN: run<zb::M_MANHATTAN>(n, a, b, dim, power, out); break;
        case zb::M_L3: run<zb::M_L3>(n, a, b, dim, power, out); break;
        case zb::M_L4: run<zb::M_L4>(n, a, b, dim, power, out); break;
        case zb::M_HAMMING: run<zb::M_HAMMING>(n, a, b, dim, power, out); break;
        case zb::M_MINKOWSKI: run<zb::M_MINKOWSKI>(n, a, b, dim, power, out); break;
        case zb::M_PNORM: run<zb::M_PNORM>(n, a, b, dim, power, out); break;
        default: return -1;
    }
    return 0;
}
extern "C" float twin_root_p(float s, int p) { return zb::root_p(s, p); }
