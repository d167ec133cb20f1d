"""Where a single-query search call spends its time (N=20000 x 384, k=10)."""
import os, sys, time, ctypes
from ctypes import c_void_p
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import latent_rag_b200 as lrb
from latent_rag_b200 import _native as nat
rng = np.random.default_rng(0)
n, d, k = 20000, 384, 10
emb = torch.from_numpy(rng.standard_normal((n, d)).astype(np.float32))
r = lrb.BruteForceRetriever(emb, [""] * n, None)
ix = r.index
lib = nat.load()
q = emb[:1].clone()
qp = q.pin_memory()
qd = q.cuda()
D = torch.empty((1, k), dtype=torch.float32).pin_memory(); I = torch.empty((1, k), dtype=torch.int64).pin_memory()
Dd = torch.empty((1, k), dtype=torch.float32, device="cuda"); Id = torch.empty((1, k), dtype=torch.int64, device="cuda")
st = c_void_p(int(torch.cuda.current_stream(0).cuda_stream))
def raw(qt, mem, outd, outi, omem):
    nat.check(lib.lk_index_search(ix._h, c_void_p(qt.data_ptr()), nat.LK_F32, mem, 1, k, c_void_p(outd.data_ptr()), c_void_p(outi.data_ptr()), omem, 0, 0, st), "s")
def bench(name, fn, iters=2000):
    for _ in range(50): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(iters): fn()
    torch.cuda.synchronize(); print(f"{name:55s} {(time.perf_counter() - t0) / iters * 1e6:8.1f} us/call", flush=True)
bench("C ABI, pageable host query -> pinned host out", lambda: raw(q, nat.LK_HOST, D, I, nat.LK_HOST))
bench("C ABI, pinned host query -> pinned host out", lambda: raw(qp, nat.LK_HOST, D, I, nat.LK_HOST))
bench("C ABI, device query -> device out (no sync per call)", lambda: raw(qd, nat.LK_DEVICE, Dd, Id, nat.LK_DEVICE))
bench("C ABI, device query -> pinned host out", lambda: raw(qd, nat.LK_DEVICE, D, I, nat.LK_HOST))
bench("ExactIndex.search(host tensor)", lambda: ix.search(q, k))
bench("BruteForceRetriever.search(host tensor)", lambda: r.search(q, k))
bench("BruteForceRetriever.retrieve(host 1-D tensor)", lambda: r.retrieve(q[0], top_k=k))
ix.set_timing(True); ix.search(qd, k, device_out=True); print("device: kernel %.1f us, whole device side %.1f us" % tuple(1e3 * t for t in ix.last_timing()))
for nn, dd in [(315, 64), (2000, 384)]:
    e2 = torch.from_numpy(rng.standard_normal((nn, dd)).astype(np.float32))
    r2 = lrb.BruteForceRetriever(e2, [""] * nn, None)
    r2.index.set_timing(True); r2.index.search(e2[:1].cuda(), k, device_out=True); r2.index.search(e2[:1].cuda(), k, device_out=True)
    print("N=%d D=%d device: kernel %.1f us, whole device side %.1f us" % ((nn, dd) + tuple(1e3 * t for t in r2.index.last_timing())))
    r2.index.set_timing(False)
    bench(f"N={nn} D={dd} BruteForceRetriever.retrieve", lambda: r2.retrieve(e2[0], top_k=k))
