// zb_select_kernel.cuh -- warp_select_visits_kernel: per-visit top-n' (/root/reference/src/database/index/lsh.rs:301-331:
// sort the members of a leaf by distance, keep the first n) for the gather path, one WARP per visit.
//
// The block-wide version (select_visits_kernel) sorts every 1024-key chunk of a visit with a shared-memory bitonic network:
// 55 barrier-separated stages for a list of which only n' <= 128 entries survive (4.9 ms for the 40 k visits of a config-2
// batch, half of a scalar-metric step).  Here the n' best (key, ordinal) pairs of a visit live in registers, entry i in
// lane i % 32, register i / 32, sorted; the warp streams the visit's keys 32 at a time (coalesced), filters them against
// the current n'-th best, and inserts the few survivors with ballots and shuffles -- the scheme of the fused scan's list.
// Order = (distance bits as u64, ordinal): lsh.rs:318 + ties by id (D3).  Tombstoned members carry the sentinel key.
//
// In a header of its own so that tests/select_emu.cpp can compile THIS SOURCE for the CPU (one std::thread per CUDA thread)
// and compare with a sort.  Needs from its includer: ForestView, Entry, u32 / u64 / u8, ZB_SENTINEL.
#pragma once

namespace zb {

#define WS_WARPS 4  // visits per block

__device__ __forceinline__ bool ws_less(u64 ka, u64 oa, u64 kb, u64 ob) { return ka < kb || (ka == kb && oa < ob); }

// value of `x` held by (lane src) -- 64-bit shuffle
__device__ __forceinline__ u64 ws_shfl64(u64 x, int src) {
    const u32 lo = __shfl_sync(0xffffffffu, (u32)x, src), hi = __shfl_sync(0xffffffffu, (u32)(x >> 32), src);
    return ((u64)hi << 32) | lo;
}
__device__ __forceinline__ u64 ws_shfl_up64(u64 x) {
    const u32 lo = __shfl_up_sync(0xffffffffu, (u32)x, 1), hi = __shfl_up_sync(0xffffffffu, (u32)(x >> 32), 1);
    return ((u64)hi << 32) | lo;
}

template <int KL>  // the list holds up to 32 * KL entries
__global__ void __launch_bounds__(32 * WS_WARPS) warp_select_visits_kernel(ForestView f, u32 nv, const u32* __restrict__ vleaf,
                                                                           const u32* __restrict__ vnp,
                                                                           const u64* __restrict__ pair_off,
                                                                           const u64* __restrict__ pair_key,
                                                                           const u32* __restrict__ ent_off,
                                                                           Entry* __restrict__ entries, const u8* __restrict__ vdone) {
    const int lane = threadIdx.x & 31;
    const u32 v = blockIdx.x * WS_WARPS + (threadIdx.x >> 5);
    if (v >= nv) return;  // whole warps leave together
    if (vdone && vdone[v]) return;
    const u32 leaf = vleaf[v];
    const u32 slots = ent_off[v + 1] - ent_off[v];
    const u32* members = f.members + f.leaf_off[leaf];
    const u64* keys = pair_key + pair_off[v];
    const u32 len = f.leaf_len[leaf];
    const int k = (int)vnp[v] < 32 * KL ? (int)vnp[v] : 32 * KL;
    u64 lk[KL], lo[KL];
#pragma unroll
    for (int r = 0; r < KL; ++r) lk[r] = lo[r] = ZB_SENTINEL;
    int cnt = 0;  // valid entries: list indices [0, cnt), cnt <= k
    for (u32 base = 0; base < len && k > 0; base += 32) {
        const u32 i = base + lane;
        const u64 key = i < len ? keys[i] : ZB_SENTINEL;
        // the current n'-th best (only meaningful when the list is full)
        u64 tk = ZB_SENTINEL, to = ZB_SENTINEL;
        if (cnt == k) {  // (every register row is shuffled and one is selected: indexing lk[] by a run-time row would put it in local memory)
#pragma unroll
            for (int r = 0; r < KL; ++r) {
                const u64 a = ws_shfl64(lk[r], (k - 1) & 31), b = ws_shfl64(lo[r], (k - 1) & 31);
                if (r == (k - 1) >> 5) { tk = a; to = b; }
            }
        }
        bool cand = key != ZB_SENTINEL && (cnt < k || key <= tk);   // cheap filter on the key alone
        u64 ord = ZB_SENTINEL;
        if (cand) {
            ord = f.ord[members[i]];
            cand = cnt < k || ws_less(key, ord, tk, to);
        }
        unsigned m = __ballot_sync(0xffffffffu, cand);
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const u64 ck = ws_shfl64(key, src), co = ws_shfl64(ord, src);
            if (cnt == k) {  // the list may have changed since the filter: compare with the n'-th best again
#pragma unroll
                for (int r = 0; r < KL; ++r) {
                    const u64 a = ws_shfl64(lk[r], (k - 1) & 31), b = ws_shfl64(lo[r], (k - 1) & 31);
                    if (r == (k - 1) >> 5) { tk = a; to = b; }
                }
                if (!ws_less(ck, co, tk, to)) continue;  // uniform over the warp
            }
            int pos = 0;  // entries that stay in front of the candidate
#pragma unroll
            for (int r = 0; r < KL; ++r)
                pos += __popc(__ballot_sync(0xffffffffu, r * 32 + lane < cnt && ws_less(lk[r], lo[r], ck, co)));
            // entries at indices >= pos move up by one (the one pushed past the end of the registers is dropped)
#pragma unroll
            for (int rr = 0; rr < KL; ++rr) {
                const int r = KL - 1 - rr;  // top register row first: it reads the row below before that one changes
                u64 uk = ws_shfl_up64(lk[r]), uo = ws_shfl_up64(lo[r]);
                if (r > 0) {  // lane 0 of this register row takes lane 31 of the row below (not yet modified)
                    const u64 pk = ws_shfl64(lk[r - 1], 31), po = ws_shfl64(lo[r - 1], 31);
                    if (lane == 0) { uk = pk; uo = po; }
                }
                const int idx = r * 32 + lane;
                if (idx > pos) { lk[r] = uk; lo[r] = uo; }
                else if (idx == pos) { lk[r] = ck; lo[r] = co; }
            }
            if (cnt < k) ++cnt;
        }
    }
    Entry* out = entries + ent_off[v];
#pragma unroll
    for (int r = 0; r < KL; ++r) {
        const u32 idx = (u32)(r * 32 + lane);
        if (idx < slots) out[idx] = (int)idx < cnt ? Entry{lk[r], lo[r]} : Entry{ZB_SENTINEL, ZB_SENTINEL};
    }
    for (u32 idx = (u32)(32 * KL) + lane; idx < slots; idx += 32) out[idx] = Entry{ZB_SENTINEL, ZB_SENTINEL};
}

}  // namespace zb
