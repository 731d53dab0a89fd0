// zb_scan.cu -- fused leaf-tile scan.  (first slice: not yet enabled; every visit takes the generic path)
#include "zb_scan.cuh"

namespace zb {

void tile_scan(ScanWorkspace& ws, const ForestView& f, u32 metric, const float* d_q, u32 nq, u32 nv, const u32* v_leaf,
               const u32* v_np, const u32* v_q, const u32* v_ent_off, u64* v_pair_len, u8* v_done, Entry* entries,
               u32 top_k, u32 min_rows, u32 tile_queries, u32 nleaves, cudaStream_t s, u64* tile_visits, u64* tile_pairs,
               u64* moved_bytes, u32* launches) {
    *tile_visits = 0;
    *tile_pairs = 0;
    *moved_bytes = 0;
    *launches = 0;
}

}  // namespace zb
