// zb_scan.cu -- the fused leaf-tile scan: candidate gather + distance scoring + per-visit top-n' in one kernel.
//
// Replaces, for visits of large leaves, the leaf branch of tree_result
// (/root/reference/src/database/index/lsh.rs:299-331: fetch every member, metric.distance, sort, take n) and
// the rescoring of search (:557-563).  Work unit = (leaf, tile of <= 16 queries that visit it):
//   * a producer warp gathers the leaf's rows with 1-D TMA bulk copies (cp.async.bulk, one per row, row
//     addresses come from the leaf's member list) into a ring of shared-memory stages guarded by mbarriers;
//   * 8 consumer warps score stage rows x tile queries with 2x2 register tiles per quad, in the canonical
//     skylake-16 accumulation order (zb_device.cuh) -- every row is read from HBM once per tile;
//   * each finished distance is filtered against the visit's current n'-th best (shared-memory threshold) and
//     pushed to a per-visit candidate buffer; a warp merges buffer + current top list with a bitonic sort
//     when the buffer could overflow; tombstoned rows are masked at the push.
// The kernel is persistent (one CTA per SM, tiles handed out by an atomic counter).
#include <cub/device/device_scan.cuh>

#include "zb_scan.cuh"

namespace zb {

#define TS_QT 16          // queries per tile (4 query groups of 4)
#define TS_RB 128         // rows per row block (16 row groups of 8, interleaved: row = i * 16 + group)
#define TS_THREADS 256    // 8 warps; warp 0 also issues the TMA copies
#define TS_NSLOT 8        // row-block slot tables kept in flight

struct TileParams {
    const u32* tile_leaf;
    const u32* tile_first;
    const u32* tile_count;
    const u32* ntiles;   // device scalar
    u32* tile_counter;   // device scalar, zeroed before launch
    const u32* order;    // visits grouped by leaf
    const u32* v_np;
    const u32* v_q;
    const u32* v_ent_off;
    Entry* entries;
    const float* queries;
    const float* qnorm;  // [nq] squared norms of the queries (cosine)
    u64* stats;          // [0] visits, [1] pairs, [2] moved bytes
    int nst;             // ring depth
    int P;               // per-query top-k region (entries): kcap + cb, power of two
    int kcap;
    int kc;              // 16-float chunks per K slice
    u32 row_stride;      // bytes between row slices of a stage (slice bytes + bank padding)
    u32 stage_bytes;
};

// ---- PTX helpers: mbarrier + 1-D bulk copy (TMA, SASS UBLKCP) ----
__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u32 bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(u32 bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(u32 bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u32 bar, u32 parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(u32 dst, const void* src, u32 bytes, u32 bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(TS_THREADS) : "memory"); }

// Warp-level bitonic sort of P entries in shared memory (P power of two >= 64), ascending (key, ord).
__device__ __forceinline__ void warp_bitonic(Entry* s, int P, int lane) {
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = lane; i < P; i += 32) {
                int ixj = i ^ j;
                if (ixj > i) {
                    Entry a = s[i], b = s[ixj];
                    bool up = (i & k) == 0;
                    if (entry_less(b, a) == up) {
                        s[i] = b;
                        s[ixj] = a;
                    }
                }
            }
            __syncwarp();
        }
    }
}

// One quad owns 8 rows x 4 queries = 32 (row, query) pairs; thread `sub` keeps lanes 4*sub..4*sub+3 of every
// pair's 16-lane accumulator.  Per 16-float chunk a thread issues 4 + 8 LDS.128 and 256 FP32 instructions, which
// keeps the kernel FP32-issue bound rather than shared-memory bound (an LDS.128 occupies the shared-memory
// datapath for 4 cycles per warp whether or not its addresses broadcast).  128 accumulator registers per thread
// are why the CTA is exactly 8 warps (2 per SM sub-partition): warp 0 doubles as the TMA producer.
template <int METRIC>
__global__ void __launch_bounds__(TS_THREADS, 1) tile_scan_kernel(ForestView f, TileParams tp) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const u32 RS = tp.row_stride;
    const u32 slice_bytes_full = (u32)tp.kc * 64u;
    // ---- carve shared memory ----
    unsigned char* s_stage = smem;                                            // [nst]{[RB][RS] rows, [QT][slice] queries}
    Entry* s_top = reinterpret_cast<Entry*>(smem + (size_t)tp.nst * tp.stage_bytes);   // [QT][P]
    u64* s_thr = reinterpret_cast<u64*>(s_top + TS_QT * tp.P);                // [QT]
    u64* s_bar = s_thr + TS_QT;                                               // full[8], empty[8]
    u32* s_slot = reinterpret_cast<u32*>(s_bar + 16);                         // [NSLOT][RB]
    int* s_cnt = reinterpret_cast<int*>(s_slot + TS_NSLOT * TS_RB);           // [QT]
    int* s_topn = s_cnt + TS_QT;                                              // [QT]
    u32* s_np = reinterpret_cast<u32*>(s_topn + TS_QT);                       // [QT]
    u32* s_visit = s_np + TS_QT;                                              // [QT]
    float* s_qn = reinterpret_cast<float*>(s_visit + TS_QT);                  // [QT]
    int* s_any = reinterpret_cast<int*>(s_qn + TS_QT);                        // [2]
    __shared__ u32 s_tile;

    const u32 bar_full = smem_u32(s_bar), bar_empty = smem_u32(s_bar + 8);
    if (tid == 0) {
        for (int i = 0; i < tp.nst; ++i) {
            mbar_init(bar_full + 8 * i, TS_THREADS / 32);   // one arrive.expect_tx per warp
            mbar_init(bar_empty + 8 * i, TS_THREADS / 32);
        }
        s_any[0] = s_any[1] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    u32 cbuf = 0, cph = 0;  // consumer ring cursor: buffer and phase parity (stage = one K slice of one row block)
    u32 pbuf = 0, pph = 0;  // producer ring cursor (runs nst - 1 stages ahead)
    u32 bc0 = 0;  // row blocks processed before this tile (slot table ring)
    u32 rnd = 0;  // push rounds so far (parity of the retry flag)
    const int cb = tp.P - tp.kcap;
    const int chunks = f.chunks;
    const int nsl = (chunks + tp.kc - 1) / tp.kc;
    const int g = warp >> 1, qd = lane >> 2, sub = lane & 3;
    const int rg = (warp & 1) * 8 + qd;  // row group: rows i * 16 + rg, i = 0..7
    const int pq = g * 4 + sub;          // the query whose 8 candidates this thread owns

    for (;;) {
        __syncthreads();  // previous tile fully retired (top lists written out, shared state reusable)
        if (tid == 0) s_tile = atomicAdd(tp.tile_counter, 1u);
        __syncthreads();
        const u32 tile = s_tile;
        if (tile >= *tp.ntiles) break;
        const u32 leaf = tp.tile_leaf[tile], first = tp.tile_first[tile], nqt = tp.tile_count[tile];
        const u32 L = f.leaf_len[leaf];
        const long long moff = f.leaf_off[leaf];
        const u32 nblocks = (L + TS_RB - 1) / TS_RB;
        const u32 total_st = nblocks * (u32)nsl;

        if (tid < TS_QT) {
            const bool on = tid < (int)nqt;
            const u32 v = on ? tp.order[first + tid] : 0u;
            s_visit[tid] = v;
            s_np[tid] = on ? tp.v_np[v] : 0u;
            s_thr[tid] = ZB_SENTINEL;
            s_cnt[tid] = 0;
            s_topn[tid] = 0;
            s_qn[tid] = (METRIC == 0 && on) ? tp.qnorm[tp.v_q[v]] : 0.f;
        }
        // producer role, spread over all 8 warps (TMA small-copy throughput scales with issuing warps: DESIGN.md 5):
        // lanes 0..15 of warp w copy rows w*16 + lane of the row block, lanes 16..17 copy queries 2w, 2w+1.
        const int my_row = warp * 16 + lane;          // valid for lane < 16
        const int my_q = warp * 2 + (lane - 16);      // valid for lane 16, 17
        const float* qsrc = nullptr;
        if ((lane == 16 || lane == 17) && my_q < (int)nqt) qsrc = tp.queries + (size_t)tp.v_q[tp.order[first + my_q]] * f.dimp;
        if (tid == 0) {
            atomicAdd(&tp.stats[0], (u64)nqt);
            atomicAdd(&tp.stats[1], (u64)nqt * L);
            atomicAdd(&tp.stats[2], (u64)L * (u64)f.dimp * 4ull);
        }
        u32 pb = 0, psl = 0;  // next stage to issue: row block, K slice
        u32 pslot = 0xFFFFFFFFu;
        auto issue_stage = [&]() {
            const u32 nrows = min((u32)TS_RB, L - pb * TS_RB);
            const u32 sbytes = (u32)min(tp.kc, chunks - (int)psl * tp.kc) * 64u;
            mbar_wait(bar_empty + 8 * pbuf, pph ^ 1);
            if (psl == 0) {
                pslot = 0xFFFFFFFFu;
                if (lane < 16 && (u32)my_row < nrows) {
                    pslot = f.members[moff + pb * TS_RB + my_row];
                    s_slot[((bc0 + pb) % TS_NSLOT) * TS_RB + my_row] = pslot;
                }
            }
            const u32 mine = ((pslot != 0xFFFFFFFFu) ? 1u : 0u) + (qsrc ? 1u : 0u);
            const u32 ncopies = __reduce_add_sync(0xffffffffu, mine);
            if (lane == 0) mbar_arrive_expect_tx(bar_full + 8 * pbuf, ncopies * sbytes);
            __syncwarp();
            unsigned char* st = s_stage + (size_t)pbuf * tp.stage_bytes;
            const u32 fb = bar_full + 8 * pbuf;
            if (pslot != 0xFFFFFFFFu)
                bulk_g2s(smem_u32(st + (size_t)my_row * RS), f.rows + (size_t)pslot * f.dimp + psl * tp.kc * 16, sbytes, fb);
            if (qsrc)
                bulk_g2s(smem_u32(st + (size_t)TS_RB * RS + (size_t)my_q * slice_bytes_full), qsrc + psl * tp.kc * 16, sbytes, fb);
            if (++psl == (u32)nsl) { psl = 0; ++pb; }
            if (++pbuf == (u32)tp.nst) { pbuf = 0; pph ^= 1; }
        };
        __syncthreads();  // per-visit state initialised (also orders s_slot reuse)
        for (u32 j = 0; j + 1 < (u32)tp.nst && j < total_st; ++j) issue_stage();

        const bool warp_active = (u32)(g * 4) < nqt;  // whole query groups idle on small tiles
        u32 j = 0;
        for (u32 b = 0; b < nblocks; ++b) {
            const u32 nrows = min((u32)TS_RB, L - b * TS_RB);
            float4 acc[8][4];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int k = 0; k < 4; ++k) acc[i][k] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int sl = 0; sl < nsl; ++sl, ++j) {
                if (j + (u32)tp.nst - 1 < total_st) issue_stage();
                const u32 buf = cbuf;
                const int kcs = min(tp.kc, chunks - sl * tp.kc);
                mbar_wait(bar_full + 8 * buf, cph);
                if (warp_active) {
                    const unsigned char* st = s_stage + (size_t)buf * tp.stage_bytes;
                    const float4* rp = reinterpret_cast<const float4*>(st + (size_t)rg * RS) + sub;
                    const float4* qp = reinterpret_cast<const float4*>(st + (size_t)TS_RB * RS + (size_t)(g * 4) * slice_bytes_full) + sub;
                    const u32 rstep = RS;  // 16 rows apart, in float4 units: 16 * RS / 16
                    const u32 qstep = slice_bytes_full / 16u;
                    for (int c = 0; c < kcs; ++c) {
                        const float4 q0 = qp[c * 4], q1 = qp[c * 4 + qstep], q2 = qp[c * 4 + 2 * qstep], q3 = qp[c * 4 + 3 * qstep];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float4 r = rp[c * 4 + i * rstep];
                            if (METRIC == 0) {
                                fma4(acc[i][0], r, q0); fma4(acc[i][1], r, q1); fma4(acc[i][2], r, q2); fma4(acc[i][3], r, q3);
                            } else {
                                l2acc4(acc[i][0], r, q0); l2acc4(acc[i][1], r, q1); l2acc4(acc[i][2], r, q2); l2acc4(acc[i][3], r, q3);
                            }
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_empty + 8 * buf);
                if (++cbuf == (u32)tp.nst) { cbuf = 0; cph ^= 1; }
            }
            // ---- epilogue of the row block: 32 sums per quad; thread `sub` keeps query pq's 8 candidates ----
            u64 keys[8];
            u32 pend = 0;
            const u32* slots = s_slot + ((bc0 + b) % TS_NSLOT) * TS_RB;
            if (warp_active) {
                const unsigned m = quad_mask();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float s0 = quad_reduce16(acc[i][0], m), s1 = quad_reduce16(acc[i][1], m);
                    const float s2 = quad_reduce16(acc[i][2], m), s3 = quad_reduce16(acc[i][3], m);
                    const float sum = sub == 0 ? s0 : (sub == 1 ? s1 : (sub == 2 ? s2 : s3));
                    const u32 r = (u32)(i * 16 + rg);
                    keys[i] = ZB_SENTINEL;
                    if (r < nrows && pq < (int)nqt) {
                        const u32 slot = slots[r];
                        if (!tomb_test(f.tomb, slot)) {
                            if (METRIC == 0) keys[i] = cos_bits(sum, f.row_norm[slot], s_qn[pq]);
                            else keys[i] = METRIC == 1 ? l2sq_bits(sum) : l2_bits(sum);
                            pend |= 1u << i;
                        }
                    }
                }
            }
            // ---- push with retry: filter against the visit's n'-th best, append to its candidate buffer; a full
            //      buffer is merged (sorted together with the current top list) and the leftovers retried ----
            const bool last_block = b + 1 == nblocks;
            for (;;) {
                if (pend) {
                    const u64 thr = s_thr[pq];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        if (!((pend >> i) & 1u)) continue;
                        if (keys[i] > thr) { pend &= ~(1u << i); continue; }
                        const int pos = atomicAdd(&s_cnt[pq], 1);
                        if (pos < cb) {
                            s_top[pq * tp.P + tp.kcap + pos] = Entry{keys[i], f.ord[slots[i * 16 + rg]]};
                            pend &= ~(1u << i);
                        }
                    }
                    if (pend) s_any[rnd & 1] = 1;
                }
                consumer_sync();
                if (tid == 0) s_any[(rnd + 1) & 1] = 0;
                for (int jq = warp; jq < (int)nqt; jq += TS_THREADS / 32) {
                    const int raw = s_cnt[jq];
                    const int cnt = raw < cb ? raw : cb;
                    if (cnt == 0 || (!last_block && raw <= cb / 2)) continue;
                    Entry* reg = s_top + jq * tp.P;
                    const int topn = s_topn[jq];
                    for (int i = lane; i < tp.P; i += 32) {
                        const bool keep = i < topn || (i >= tp.kcap && i < tp.kcap + cnt);
                        if (!keep) reg[i] = Entry{ZB_SENTINEL, ZB_SENTINEL};
                    }
                    __syncwarp();
                    warp_bitonic(reg, tp.P, lane);
                    if (lane == 0) {
                        const int np = (int)s_np[jq];
                        const int tot = topn + cnt;
                        const int nt = tot < np ? tot : np;
                        s_topn[jq] = nt;
                        s_cnt[jq] = 0;
                        s_thr[jq] = nt == np ? reg[np - 1].key : ZB_SENTINEL;
                    }
                    __syncwarp();
                }
                consumer_sync();
                const int again = s_any[rnd & 1];
                ++rnd;
                if (!again) break;
            }
        }
        // ---- write the visits' top lists (min(n', local live) entries, padded to the visit's slot count) ----
        for (int jq = warp; jq < (int)nqt; jq += TS_THREADS / 32) {
            const u32 v = s_visit[jq];
            const u32 e0 = tp.v_ent_off[v], e1 = tp.v_ent_off[v + 1];
            const int topn = s_topn[jq];
            const Entry* reg = s_top + jq * tp.P;
            for (u32 i = lane; i < e1 - e0; i += 32)
                tp.entries[e0 + i] = (int)i < topn ? reg[i] : Entry{ZB_SENTINEL, ZB_SENTINEL};
        }
        bc0 += nblocks;
    }
}

// =====================================================================================================
// grouping of visits by leaf and tile construction (all on device; no host round trip)
// =====================================================================================================
__global__ void ts_count_kernel(ForestView f, u32 nv, const u32* __restrict__ v_leaf, const u32* __restrict__ v_np,
                                u32 min_rows, u32 kmax, u32* __restrict__ leaf_count) {
    u32 v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    u32 leaf = v_leaf[v];
    if (f.leaf_len[leaf] >= min_rows && v_np[v] <= kmax) atomicAdd(&leaf_count[leaf], 1u);
}
__global__ void ts_tilecount_kernel(u32 nleaves, const u32* __restrict__ leaf_count, u32 tq, u32* __restrict__ tile_cnt) {
    u32 l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l <= nleaves) tile_cnt[l] = l < nleaves ? (leaf_count[l] + tq - 1) / tq : 0u;
}
__global__ void ts_scatter_kernel(ForestView f, u32 nv, const u32* __restrict__ v_leaf, const u32* __restrict__ v_np,
                                  u32 min_rows, u32 kmax, const u32* __restrict__ leaf_start, u32* __restrict__ leaf_cursor,
                                  u32* __restrict__ order, u64* __restrict__ v_pair_len, u8* __restrict__ v_done) {
    u32 v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    u32 leaf = v_leaf[v];
    if (f.leaf_len[leaf] >= min_rows && v_np[v] <= kmax) {
        u32 pos = leaf_start[leaf] + atomicAdd(&leaf_cursor[leaf], 1u);
        order[pos] = v;
        v_pair_len[v] = 0;
        v_done[v] = 1;
    }
}
__global__ void ts_filltiles_kernel(u32 nleaves, const u32* __restrict__ leaf_count, const u32* __restrict__ leaf_start,
                                    const u32* __restrict__ tile_start, u32 tq, u32* __restrict__ tile_leaf,
                                    u32* __restrict__ tile_first, u32* __restrict__ tile_count) {
    u32 l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nleaves) return;
    const u32 c = leaf_count[l], ts = tile_start[l];
    if (!c) return;
    const u32 nt = (c + tq - 1) / tq, base = c / nt, rem = c % nt;   // balanced: tiles differ by at most one query
    for (u32 j = 0, done = 0; j < nt; ++j) {
        const u32 n = base + (j < rem ? 1u : 0u);
        tile_leaf[ts + j] = l;
        tile_first[ts + j] = leaf_start[l] + done;
        tile_count[ts + j] = n;
        done += n;
    }
}

static size_t ts_smem_bytes(int nst, u32 stage_bytes, int P) {
    return (size_t)nst * stage_bytes + (size_t)TS_QT * P * sizeof(Entry) + TS_QT * 8 + 16 * 8 +
           (size_t)TS_NSLOT * TS_RB * 4 + TS_QT * 4 * 5 + 8 + 256;
}

void tile_scan(ScanWorkspace& ws, const ForestView& f, u32 metric, const float* d_q, const float* d_qnorm, u32 nq, u32 nv, const u32* v_leaf,
               const u32* v_np, const u32* v_q, const u32* v_ent_off, u64* v_pair_len, u8* v_done, Entry* entries,
               u32 top_k, u32 min_rows, u32 tile_queries, u32 nleaves, cudaStream_t s, u64* tile_visits, u64* tile_pairs,
               u64* moved_bytes, u32* launches) {
    *tile_visits = *tile_pairs = *moved_bytes = 0;
    *launches = 0;
    if (!nv || !nleaves || top_k > 128) return;
    const u32 tq = tile_queries >= 1 && tile_queries <= TS_QT ? tile_queries : TS_QT;
    // per-query top-k region: kcap + candidate buffer, power of two
    int kcap = 16, P = 64;
    if (top_k > 16) { kcap = 32; P = 128; }
    if (top_k > 32) { kcap = 128; P = 256; }
    // K slice: up to 6 chunks (384 B per row slice); rows of a stage are padded so that adjacent rows fall into
    // different halves of the 32 banks (two quads share one LDS.128 phase).
    const int kc = f.chunks < 6 ? f.chunks : 6;
    const u32 slice = (u32)kc * 64u;
    const u32 row_stride = slice + ((slice % 128u) == 0 ? 64u : 0u);
    const u32 stage_bytes = TS_RB * row_stride + TS_QT * slice;
    int dev = 0, max_smem = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int nst = 0;
    for (int cand = 6; cand >= 2; --cand)
        if (ts_smem_bytes(cand, stage_bytes, P) <= (size_t)max_smem) { nst = cand; break; }
    if (!nst) return;
    const size_t smem = ts_smem_bytes(nst, stage_bytes, P);

    ws.leaf_count.ensure(nleaves + 1);
    ws.leaf_start.ensure(nleaves + 1);
    ws.leaf_cursor.ensure(nleaves + 1);
    ws.order.ensure(nv);
    ws.tile_leaf.ensure(nv);
    ws.tile_first.ensure(nv);
    ws.counters.ensure(16);
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, (const u32*)nullptr, (u32*)nullptr, (long long)(nleaves + 1));
    ws.tmp.ensure(tmp_bytes + 256);

    ZB_CUDA(cudaMemsetAsync(ws.leaf_count.p, 0, (size_t)(nleaves + 1) * 4, s));
    ZB_CUDA(cudaMemsetAsync(ws.leaf_cursor.p, 0, (size_t)(nleaves + 1) * 4, s));
    ZB_CUDA(cudaMemsetAsync(ws.counters.p, 0, 16 * 4, s));
    const u32 kmax = top_k;
    ts_count_kernel<<<(nv + 255) / 256, 256, 0, s>>>(f, nv, v_leaf, v_np, min_rows, kmax, ws.leaf_count.p);
    cub::DeviceScan::ExclusiveSum(ws.tmp.p, tmp_bytes, ws.leaf_count.p, ws.leaf_start.p, (long long)(nleaves + 1), s);
    ts_scatter_kernel<<<(nv + 255) / 256, 256, 0, s>>>(f, nv, v_leaf, v_np, min_rows, kmax, ws.leaf_start.p,
                                                       ws.leaf_cursor.p, ws.order.p, v_pair_len, v_done);
    // tiles: per-leaf counts -> exclusive scan (tile_start, total at [nleaves]) -> fill
    ws.tile_per_leaf.ensure(nleaves + 1);
    ws.tile_start.ensure(nleaves + 1);
    ws.tile_cnt.ensure(nv);
    ts_tilecount_kernel<<<(nleaves + 256) / 256, 256, 0, s>>>(nleaves, ws.leaf_count.p, tq, ws.tile_per_leaf.p);
    cub::DeviceScan::ExclusiveSum(ws.tmp.p, tmp_bytes, ws.tile_per_leaf.p, ws.tile_start.p, (long long)(nleaves + 1), s);
    ts_filltiles_kernel<<<(nleaves + 255) / 256, 256, 0, s>>>(nleaves, ws.leaf_count.p, ws.leaf_start.p, ws.tile_start.p, tq,
                                                              ws.tile_leaf.p, ws.tile_first.p, ws.tile_cnt.p);

    TileParams tp;
    tp.tile_leaf = ws.tile_leaf.p;
    tp.tile_first = ws.tile_first.p;
    tp.tile_count = ws.tile_cnt.p;
    tp.ntiles = ws.tile_start.p + nleaves;
    tp.tile_counter = ws.counters.p;
    tp.order = ws.order.p;
    tp.v_np = v_np;
    tp.v_q = v_q;
    tp.v_ent_off = v_ent_off;
    tp.entries = entries;
    tp.queries = d_q;
    tp.qnorm = d_qnorm;
    tp.stats = reinterpret_cast<u64*>(ws.counters.p + 4);
    tp.nst = nst;
    tp.P = P;
    tp.kcap = kcap;
    tp.kc = kc;
    tp.row_stride = row_stride;
    tp.stage_bytes = stage_bytes;
    const int grid = sms;
    if (metric == 0) {
        ZB_CUDA(cudaFuncSetAttribute(tile_scan_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tile_scan_kernel<0><<<grid, TS_THREADS, smem, s>>>(f, tp);
    } else if (metric == 1) {
        ZB_CUDA(cudaFuncSetAttribute(tile_scan_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tile_scan_kernel<1><<<grid, TS_THREADS, smem, s>>>(f, tp);
    } else {
        ZB_CUDA(cudaFuncSetAttribute(tile_scan_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tile_scan_kernel<2><<<grid, TS_THREADS, smem, s>>>(f, tp);
    }
    ZB_CUDA(cudaGetLastError());
    u64 h[3] = {0, 0, 0};
    ZB_CUDA(cudaMemcpyAsync(h, tp.stats, 24, cudaMemcpyDeviceToHost, s));
    ZB_CUDA(cudaStreamSynchronize(s));
    *tile_visits = h[0];
    *tile_pairs = h[1];
    *moved_bytes = h[2];
    *launches = 7;
}

}  // namespace zb
