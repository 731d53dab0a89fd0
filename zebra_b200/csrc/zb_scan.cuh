// zb_scan.cuh -- the fused leaf-tile scan (the bandwidth-critical kernel of the query path).
#pragma once
#include "zb_host.h"
#include "zb_kernels.cuh"

namespace zb {

struct ScanWorkspace {
    DBuf<u32> leaf_count, leaf_start, leaf_cursor, tile_per_leaf, tile_start, order, tile_leaf, tile_first, tile_cnt;
    DBuf<u32> counters;
    DBuf<u8> tmp;
};

// Handles every visit whose leaf holds at least `min_rows` rows: groups those visits by leaf, and for each
// (leaf, tile of <= tile_queries queries) streams the leaf's rows once through shared memory, scores them
// against all queries of the tile in the canonical order and keeps each visit's top-n' on chip.  Handled
// visits get v_done = 1, v_pair_len = 0 and their entries written; the rest is left to the generic path.
void tile_scan(ScanWorkspace& ws, const ForestView& f, u32 metric, const float* d_q, const float* d_qnorm, u32 nq, u32 nv, const u32* v_leaf,
               const u32* v_np, const u32* v_q, const u32* v_ent_off, u64* v_pair_len, u8* v_done, Entry* entries,
               u32 top_k, u32 min_rows, u32 tile_queries, u32 nleaves, cudaStream_t s, u64* tile_visits, u64* tile_pairs,
               u64* moved_bytes, u32* launches);

}  // namespace zb
