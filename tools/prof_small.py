"""Per-call latency on a tiny corpus (the shape of the reference's own logged run: 315 x 64, one query per
call, logs/benchmarks/experiments.csv): python tools/prof_small.py"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import latent_rag_b200 as lrb
rng = np.random.default_rng(0)
for n, d in [(315, 64), (20000, 384)]:
    emb = torch.from_numpy(rng.standard_normal((n, d)).astype(np.float32))
    emb /= emb.norm(dim=1, keepdim=True)
    q = emb[:2000 % n + 200].clone()
    for cls in ("brute", "faiss"):
        if cls == "brute":
            r = lrb.BruteForceRetriever(emb, [""] * n, None)
        else:
            r = lrb.FAISSEmbeddingRetriever(d, index_type="flatip"); r.build(emb, [""] * n)
        for qq in q[:20]:
            r.retrieve(qq, top_k=10)
        r.get_stats(reset=True)
        t0 = time.perf_counter()
        for qq in q:
            r.retrieve(qq, top_k=10)
        wall = (time.perf_counter() - t0) / len(q) * 1e3
        st = r.get_stats()
        p50 = float(np.percentile(st["per_query_ms"], 50))
        print(f"{cls:5s} N={n} D={d}: retrieve() {wall:.4f} ms/query wall, StatsTracker p50 {p50:.4f} ms -> qps=1000/p50 {1000 / p50:,.0f}", flush=True)
        t0 = time.perf_counter()
        ids, _ = r.retrieve_batch(q, top_k=10)
        print(f"      retrieve_batch({len(q)}) {(time.perf_counter() - t0) * 1e3:.3f} ms total", flush=True)
