"""Host-side cost of one search call (no device sync inside the loop): python tools/prof_host.py"""
import cProfile, pstats, sys, time, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import latent_rag_b200 as lrb
g = torch.Generator(device="cuda").manual_seed(1)
ix = lrb.ExactIndex(384, 20000, metric="cosine"); ix.add(torch.randn((20000, 384), generator=g, device="cuda"))
q = torch.randn((10000, 384), generator=g, device="cuda")
for _ in range(5): ix.search(q, 10, device_out=True)
torch.cuda.synchronize()
for n in (1, 50):
    t0 = time.perf_counter()
    for _ in range(n): ix.search(q, 10, device_out=True)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"{n} calls: host enqueue {1e6 * (t1 - t0) / n:.1f} us/call, until device idle {1e6 * (t2 - t0) / n:.1f} us/call")
pr = cProfile.Profile(); pr.enable()
for _ in range(50): ix.search(q, 10, device_out=True)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
