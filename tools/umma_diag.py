"""Bring-up aid for the tcgen05 kernel: dump unit 0's raw 128x128 accumulator
(LK_UMMA_DUMP) and compare it with Q @ E^T computed on the CPU from the same bf16 values.
Run on a B200:  python tools/umma_diag.py [--lbo N --sbo N]"""
import argparse
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("--lbo", type=int)
ap.add_argument("--sbo", type=int)
ap.add_argument("--dims", default="64,384,768")
args = ap.parse_args()
if args.lbo is not None:
    os.environ["LK_UMMA_LBO"] = str(args.lbo)
if args.sbo is not None:
    os.environ["LK_UMMA_SBO"] = str(args.sbo)
os.environ["LK_FORCE_KERNEL"] = "umma"

import latent_rag_b200 as lrb  # noqa: E402

rng = np.random.default_rng(0)
status = 0
for dim in [int(x) for x in args.dims.split(",")]:
    for n, b in [(128, 128), (300, 40)]:
        emb = torch.from_numpy(rng.standard_normal((n, dim)).astype(np.float32)).bfloat16().float()
        q = torch.from_numpy(rng.standard_normal((b, dim)).astype(np.float32)).bfloat16().float()
        dump = os.path.join(tempfile.gettempdir(), f"lk_tile_{dim}_{n}.bin")
        os.environ["LK_UMMA_DUMP"] = dump
        ix = lrb.ExactIndex(dim, n, metric="euclidean")
        ix.add(emb)
        try:
            d, i = ix.search(q, 10)
        except Exception as e:  # noqa: BLE001
            print(f"dim={dim} n={n} b={b}: search FAILED: {e}", flush=True)
            status = 1
            continue
        tile = np.fromfile(dump, dtype=np.float32).reshape(128, 128)
        ref = (q.double() @ emb[:128].double().T).numpy()
        rows, cols = min(b, 128), min(n, 128)
        got = tile[:rows, :cols]
        err = np.abs(got - ref[:rows, :cols])
        scale = np.abs(ref).max()
        print(f"dim={dim} n={n} b={b}: tile max|err|={np.nanmax(err):.3e} (scale {scale:.2f}) "
              f"nan={int(np.isnan(got).sum())}", flush=True)
        if not np.nanmax(err) < 1e-3 * scale or np.isnan(got).any():
            status = 1
            # hypotheses: transposed, or a permutation of rows / columns
            if rows == cols and np.abs(got.T - ref[:rows, :cols]).max() < 1e-3 * scale:
                print("  -> tile is TRANSPOSED")
            flat_ref = {round(float(v), 3): (r, c) for (r, c), v in np.ndenumerate(ref[:rows, :cols])}
            hits = [(rc, flat_ref.get(round(float(v), 3))) for rc, v in np.ndenumerate(got[:4, :16])]
            print("  first got[r,c] -> ref position:", hits[:24])
            print("  got[0,:8]", got[0, :8], "\n  ref[0,:8]", ref[0, :8])
        e2 = (emb * emb).sum(1)
        q2 = (q * q).sum(1, keepdim=True)
        sc = -(q2 + e2[None, :] - 2 * (q @ emb.T))
        vals, idx = torch.topk(sc, 10, dim=1)
        same = (idx.numpy() == i).mean()
        print(f"    top-10 index agreement with CPU: {same:.4f}; max score err "
              f"{np.abs(vals.numpy() - d).max():.3e}", flush=True)
        if same < 0.999:
            status = 1
        ix.close()
print("DIAG", "OK" if status == 0 else "MISMATCH")
sys.exit(status)
