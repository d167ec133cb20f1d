"""Host-side cost of one search call (no device sync inside the loop):
    python tools/prof_host.py [--storage bf16|fp32] [--rows 20000] [--batch 10000]"""
import argparse, cProfile, pstats, sys, time, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import latent_rag_b200 as lrb
ap = argparse.ArgumentParser()
ap.add_argument("--storage", default="bf16")
ap.add_argument("--rows", type=int, default=20000)
ap.add_argument("--batch", type=int, default=10000)
ap.add_argument("--profile", type=int, default=0)
a = ap.parse_args()
g = torch.Generator(device="cuda").manual_seed(1)
ix = lrb.ExactIndex(384, a.rows, metric="cosine", storage=a.storage); ix.add(torch.randn((a.rows, 384), generator=g, device="cuda"))
q = torch.randn((a.batch, 384), generator=g, device="cuda")
for _ in range(5): ix.search(q, 10, device_out=True)
torch.cuda.synchronize()
for n in (1, 10, 50, 200, 400, 50):
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): out = ix.search(q, 10, device_out=True)
    e1.record()
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"{n} calls: host enqueue {1e6 * (t1 - t0) / n:.1f} us/call, until device idle {1e6 * (t2 - t0) / n:.1f} us/call, "
          f"events {1e3 * e0.elapsed_time(e1) / n:.1f} us/call")
if a.profile:
    pr = cProfile.Profile(); pr.enable()
    for _ in range(50): ix.search(q, 10, device_out=True)
    pr.disable(); torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
