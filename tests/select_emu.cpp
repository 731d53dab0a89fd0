// select_emu.cpp -- TEST INFRASTRUCTURE.  Runs the SOURCE of warp_select_visits_kernel (zebra_b200/csrc/zb_select_kernel.cuh) on
// the CPU: one std::thread per lane of a warp, __shfl_sync / __shfl_up_sync / __ballot_sync through a per-warp mailbox with a
// barrier of 32.  One warp (= one visit) at a time.  tests/test_select.py compares every visit's list with a sort.
// Build with -fvisibility=hidden -Wl,-Bsymbolic (the real kernel's host stub in libzebra_b200.so has the same mangled name).
#include <stdint.h>
#include <string.h>

#include <barrier>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
struct Dim3 { unsigned x, y, z; };
static thread_local Dim3 emu_threadIdx, emu_blockIdx;
#define threadIdx emu_threadIdx
#define blockIdx emu_blockIdx

struct WarpBox {
    uint32_t slot[32];
    std::barrier<> bar{32};
};
static WarpBox emu_warp;
static inline uint32_t emu_exchange(uint32_t v, int src_lane) {
    const int lane = threadIdx.x & 31;
    emu_warp.slot[lane] = v;
    emu_warp.bar.arrive_and_wait();
    const uint32_t r = emu_warp.slot[src_lane & 31];
    emu_warp.bar.arrive_and_wait();
    return r;
}
static inline uint32_t __shfl_sync(unsigned, uint32_t v, int src) { return emu_exchange(v, src); }
static inline uint32_t __shfl_up_sync(unsigned, uint32_t v, unsigned delta) {
    const int lane = threadIdx.x & 31;
    return emu_exchange(v, lane >= (int)delta ? lane - (int)delta : lane);  // lanes below delta keep their own value
}
static inline unsigned __ballot_sync(unsigned, bool pred) {
    const int lane = threadIdx.x & 31;
    emu_warp.slot[lane] = pred ? 1u : 0u;
    emu_warp.bar.arrive_and_wait();
    unsigned m = 0;
    for (int l = 0; l < 32; ++l) m |= emu_warp.slot[l] << l;
    emu_warp.bar.arrive_and_wait();
    return m;
}
static inline int __ffs(unsigned x) { return __builtin_ffs((int)x); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }

namespace zb {
typedef unsigned long long u64;
typedef unsigned int u32;
typedef unsigned char u8;
#define ZB_SENTINEL 0xFFFFFFFFFFFFFFFFull
struct Entry { u64 key, ord; };
struct ForestView {  // the fields the kernel reads
    const long long* leaf_off;
    const u32* leaf_len;
    const u32* members;
    const u64* ord;
};
}  // namespace zb

#include "../zebra_b200/csrc/zb_select_kernel.cuh"

extern "C" __attribute__((visibility("default"))) void emu_warp_select(int KL, uint32_t nv, const long long* leaf_off,
                                                                       const uint32_t* leaf_len, const uint32_t* members,
                                                                       const uint64_t* ord, const uint32_t* vleaf, const uint32_t* vnp,
                                                                       const uint64_t* pair_off, const uint64_t* pair_key,
                                                                       const uint32_t* ent_off, uint64_t* entries2, const uint8_t* vdone) {
    zb::ForestView f{leaf_off, leaf_len, members, reinterpret_cast<const zb::u64*>(ord)};
    for (uint32_t v = 0; v < nv; ++v) {  // one warp per visit: block v / WS_WARPS, warp v % WS_WARPS
        std::vector<std::thread> th;
        for (unsigned l = 0; l < 32; ++l)
            th.emplace_back([&, l] {
                emu_blockIdx = Dim3{v / WS_WARPS, 0, 0};
                emu_threadIdx = Dim3{(v % WS_WARPS) * 32 + l, 0, 0};
                auto* po = reinterpret_cast<const zb::u64*>(pair_off);
                auto* pk = reinterpret_cast<const zb::u64*>(pair_key);
                auto* en = reinterpret_cast<zb::Entry*>(entries2);
                if (KL == 1) zb::warp_select_visits_kernel<1>(f, nv, vleaf, vnp, po, pk, ent_off, en, vdone);
                else if (KL == 2) zb::warp_select_visits_kernel<2>(f, nv, vleaf, vnp, po, pk, ent_off, en, vdone);
                else zb::warp_select_visits_kernel<4>(f, nv, vleaf, vnp, po, pk, ent_off, en, vdone);
            });
        for (auto& x : th) x.join();
    }
}
