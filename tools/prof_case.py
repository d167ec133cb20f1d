"""One search configuration for ncu: build an index on the device, run a few searches.
    python tools/prof_case.py --rows 10000000 --batch 4096 --iters 3 [--dim 384 --k 10 --metric cosine]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import latent_rag_b200 as lrb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=10_000_000)
ap.add_argument("--dim", type=int, default=384)
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--k", type=int, default=10)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--metric", default="cosine")
ap.add_argument("--kernel", default="auto")
ap.add_argument("--storage", default="bf16")
a = ap.parse_args()

ix = lrb.ExactIndex(a.dim, a.rows, metric=a.metric, storage=a.storage)
dt = torch.bfloat16 if a.storage == "bf16" else torch.float32
g = torch.Generator(device="cuda").manual_seed(1)
for lo in range(0, a.rows, 1_000_000):
    n = min(1_000_000, a.rows - lo)
    ix.add(torch.randn((n, a.dim), generator=g, device="cuda").to(dt))
q = torch.randn((a.batch, a.dim), generator=g, device="cuda").to(dt)
ix.set_timing(True)
for it in range(a.iters):
    ix.search(q, a.k, device_out=True, kernel=a.kernel)
    print(f"iter {it}: search kernel {ix.last_timing()[0]:.3f} ms, device total {ix.last_timing()[1]:.3f} ms", flush=True)
torch.cuda.synchronize()
