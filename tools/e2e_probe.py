"""Where the end-to-end call spends its time beyond the device-resident step (1 GPU, BASELINE config 2 shape)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zebra_b200 as z

dev = torch.device("cuda", 0); torch.cuda.set_device(0)
rows, dim, nq, k, nb = 1_000_000, 768, 10_000, 10, 12
ix = z.LSHIndex(dim, z.LSHIndexOptions(2048, 4), z.L2Distance(), device=0, seed=0)
d_rows = torch.empty((rows, dim), dtype=torch.float32, device=dev)
z.synth_fill_device(0, d_rows.data_ptr(), 0, 1, rows, dim, 0, 1)
ix.add_device(d_rows.data_ptr(), rows); del d_rows
d_q = torch.empty((nb, nq, dim), dtype=torch.float32, device=dev)
for b in range(nb):
    z.synth_fill_device(0, d_q[b].data_ptr(), b * nq, 1, nq, dim, 1, 1)
h_q = d_q.cpu().pin_memory()
d_ord = torch.empty((nq, k), dtype=torch.int64, device=dev); d_bits = torch.empty_like(d_ord); d_cnt = torch.empty((nq,), dtype=torch.int32, device=dev)
h_ord = torch.empty((nq, k), dtype=torch.int64).pin_memory(); h_bits = torch.empty_like(h_ord).pin_memory(); h_cnt = torch.empty((nq,), dtype=torch.int32).pin_memory()
def wall(f, n=8):
    torch.cuda.synchronize(); t = time.perf_counter()
    for i in range(n): f(2 + i)
    torch.cuda.synchronize(); return (time.perf_counter() - t) * 1e3 / n
for b in range(2):
    ix.search_batch_device(nq, d_q[b].data_ptr(), k, d_ord.data_ptr(), d_bits.data_ptr(), d_cnt.data_ptr())
    ix.search_slice_ptr(nq, h_q[b].data_ptr(), k, h_ord.data_ptr(), h_bits.data_ptr(), h_cnt.data_ptr())
stage = torch.empty((nq, dim), dtype=torch.float32, device=dev)
print("H2D 30.7 MB pinned, torch copy_      %.3f ms" % wall(lambda b: stage.copy_(h_q[b], non_blocking=True)))
print("D2H ord+bits+cnt                     %.3f ms" % wall(lambda b: (h_ord.copy_(d_ord, non_blocking=True), h_bits.copy_(d_bits, non_blocking=True), h_cnt.copy_(d_cnt, non_blocking=True))))
print("device-resident call                 %.3f ms" % wall(lambda b: ix.search_batch_device(nq, d_q[b].data_ptr(), k, d_ord.data_ptr(), d_bits.data_ptr(), d_cnt.data_ptr())))
print("host-buffer call                     %.3f ms" % wall(lambda b: ix.search_slice_ptr(nq, h_q[b].data_ptr(), k, h_ord.data_ptr(), h_bits.data_ptr(), h_cnt.data_ptr())))
def pf(b):
    ix.search_prefetch_ptr(nq, h_q[b + 1].data_ptr())
    ix.search_slice_ptr(nq, h_q[b].data_ptr(), k, h_ord.data_ptr(), h_bits.data_ptr(), h_cnt.data_ptr())
print("host-buffer call + prefetch of next  %.3f ms" % wall(pf))
def manual(b):
    stage.copy_(h_q[b], non_blocking=True); torch.cuda.synchronize()
    ix.search_batch_device(nq, stage.data_ptr(), k, d_ord.data_ptr(), d_bits.data_ptr(), d_cnt.data_ptr())
    h_ord.copy_(d_ord, non_blocking=True); h_bits.copy_(d_bits, non_blocking=True); h_cnt.copy_(d_cnt, non_blocking=True); torch.cuda.synchronize()
print("manual H2D + device call + D2H       %.3f ms" % wall(manual))
st = ix.stats(); print({k_: st[k_] for k_ in ("last_ms_plan", "last_ms_scan", "last_ms_select", "last_ms_merge", "last_ms_total")})
# host time of the announcement alone, and of the search that follows it
tp, ts = [], []
for b in range(2, 10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ix.search_prefetch_ptr(nq, h_q[b + 1].data_ptr()); t1 = time.perf_counter()
    ix.search_slice_ptr(nq, h_q[b].data_ptr(), k, h_ord.data_ptr(), h_bits.data_ptr(), h_cnt.data_ptr()); t2 = time.perf_counter()
    tp.append((t1 - t0) * 1e3); ts.append((t2 - t1) * 1e3)
print("prefetch call host ms", [round(x, 3) for x in tp]); print("search call host ms  ", [round(x, 3) for x in ts])
# the same overlap done by hand with torch streams: copy of the next batch on a side stream during the device call
side = torch.cuda.Stream(); stage2 = [torch.empty((nq, dim), dtype=torch.float32, device=dev) for _ in range(2)]
def by_hand(b):
    with torch.cuda.stream(side):
        stage2[(b + 1) & 1].copy_(h_q[b + 1], non_blocking=True)
    ix.search_batch_device(nq, stage2[b & 1].data_ptr(), k, d_ord.data_ptr(), d_bits.data_ptr(), d_cnt.data_ptr())
    side.synchronize()
stage2[0].copy_(h_q[2]); print("device call with a torch side-stream H2D of the next batch  %.3f ms" % wall(by_hand))

def pf_sync(b):
    pf(b); torch.cuda.synchronize()
print("tight loop, prefetch                 %.3f ms" % wall(pf))
print("tight loop, prefetch + device sync   %.3f ms" % wall(pf_sync))
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
cs = torch.cuda.ExternalStream(ix.stream_ptr(), device=dev)
for b in range(2, 6):   # when does the announced copy finish relative to the search?  (events on torch's current stream = legacy default: orders after everything)
    t0 = time.perf_counter(); ix.search_prefetch_ptr(nq, h_q[b + 1].data_ptr())
    ix.search_slice_ptr(nq, h_q[b].data_ptr(), k, h_ord.data_ptr(), h_bits.data_ptr(), h_cnt.data_ptr()); t1 = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print("  search %.3f ms, then device sync %.3f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
