"""Measure every BASELINE.json config on one GPU (kernel time by CUDA events, QPS, roofline
fractions).  One JSON line per case -> gpurun_out/configs.jsonl.
    python tools/bench_configs.py [--only c1,c2,c3,c4,c5] [--scale 1.0]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import latent_rag_b200 as lrb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--only", default="c1,c2,c3,c4,c5")
ap.add_argument("--scale", type=float, default=1.0, help="scale corpus sizes (smoke runs)")
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "configs.jsonl"))
args = ap.parse_args()
only = set(args.only.split(","))
PK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}
dev = torch.device("cuda:0")
os.makedirs(os.path.dirname(args.out), exist_ok=True)
out_f = open(args.out, "a")


def emit(rec):
    line = json.dumps(rec)
    print(line, flush=True)
    out_f.write(line + "\n")
    out_f.flush()


def gen(n, d, seed, unit=True, chunk=1_000_000, aniso=None):
    g = torch.Generator(device=dev).manual_seed(seed)
    for lo in range(0, n, chunk):
        x = torch.randn((min(chunk, n - lo), d), generator=g, device=dev)
        if aniso is not None:
            x = x @ aniso
        if unit:
            x = x / x.norm(dim=1, keepdim=True)
        yield x


def time_search(ix, q, k, iters, kernel="auto"):
    ix.set_timing(True)
    ks, ts = [], []
    for _ in range(2):
        ix.search(q, k, device_out=True, kernel=kernel)
    torch.cuda.synchronize()
    for _ in range(iters):
        t0 = time.perf_counter()
        ix.search(q, k, device_out=True, kernel=kernel)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
        ks.append(ix.last_timing()[0])
    ix.set_timing(False)
    return float(np.median(ks)), float(np.median(ts)) * 1e3


def search_case(name, ix, n, d, q, k, iters, elem=2, note=""):
    b = q.size(0)
    kms, wall_ms = time_search(ix, q, k, iters)
    flops = 2.0 * b * n * d
    byts = n * d * elem + n * 4
    emit({"case": name, "rows": n, "dim": d, "batch": b, "k": k, "kernel_ms": kms, "call_ms": wall_ms,
          "qps": b / (wall_ms * 1e-3), "tflops": flops / (kms * 1e-3) / 1e12,
          "tensor_frac_sustained": flops / (kms * 1e-3) / 1e12 / PK["bf16_tflops_sustained"],
          "gbs_one_pass": byts / (kms * 1e-3) / 1e9, "hbm_frac_one_pass": byts / (kms * 1e-3) / 1e9 / PK["hbm_gbs"],
          "note": note})


S = args.scale

if "c1" in only:  # SQuAD-shaped: 10k queries x 20k docs x 384, cosine top-10
    n, b, d = 20_000, 10_000, 384
    ix = lrb.ExactIndex(d, n, metric="cosine")
    for x in gen(n, d, 1234):
        ix.add(x)
    q = next(gen(b, d, 4321))
    search_case("c1 cosine top-10 (bf16, tcgen05)", ix, n, d, q, 10, 20)
    ix32 = lrb.ExactIndex(d, n, metric="cosine", storage="fp32")
    for x in gen(n, d, 1234):
        ix32.add(x)
    search_case("c1 cosine top-10 (fp32 exact, SIMT)", ix32, n, d, q, 10, 5, elem=4)
    ix.close(); ix32.close()

if "c2" in only:  # CAE/VAE latent corpus encode + cosine top-10: 1M docs x 10k queries
    n, b = int(1_000_000 * S), 10_000
    gold = os.path.join(ROOT, "tests", "golden")
    for kind in ("cae", "vae"):
        ae = lrb.load_autoencoder(kind, os.path.join(gold, f"ae_weights_{kind}.npz"), device=0)
        x = next(gen(n, 384, 1234, chunk=n))
        xq = next(gen(b, 384, 4321))
        enc = lambda t: (ae.encode(t)[0] if kind == "vae" else ae.encode(t))
        m = n + b
        for kern in ("simt", "umma"):
            ae.set_kernel(kern)
            enc(x[:1024]); torch.cuda.synchronize()
            best = 1e30
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); z = enc(x); zq = enc(xq); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            ms = best
            emit({"case": f"c2 {kind} encode 384->512->64 ({kern})", "vectors": m, "ms": ms,
                  "vec_per_s": m / (ms * 1e-3), "tflops": m * 458752 / (ms * 1e-3) / 1e12,
                  "gbs": m * 1792 / (ms * 1e-3) / 1e9, "hbm_frac": m * 1792 / (ms * 1e-3) / 1e9 / PK["hbm_gbs"]})
        ix = lrb.ExactIndex(64, n, metric="cosine")
        ix.add(z)
        search_case(f"c2 {kind} latent cosine top-10", ix, n, 64, zq, 10, 10)
        ix.close()
        del x, z

if "c3" in only:  # Mahalanobis top-10, 384-d, full covariance, 10M docs, batch 1/64/4096
    n, d = int(10_000_000 * S), 384
    g = torch.Generator(device=dev).manual_seed(7)
    rot = torch.linalg.qr(torch.randn((d, d), generator=g, device=dev))[0]
    A = torch.diag(torch.linspace(0.2, 2.0, d, device=dev)) @ rot
    cov = (A.T @ A).double().cpu().numpy()  # population covariance of x = g A
    prec = np.linalg.inv(cov)
    from latent_rag_b200.retrieval.common import whitener_from_precision
    t0 = time.perf_counter()
    ix = lrb.ExactIndex(d, n, metric="mahalanobis", whiten=whitener_from_precision(prec))
    for x in gen(n, d, 1234, unit=False, aniso=A):
        ix.add(x)
    torch.cuda.synchronize()
    emit({"case": "c3 mahalanobis index build (fp64 whitening + tiling)", "rows": n, "s": time.perf_counter() - t0})
    for b in (1, 64, 4096):
        q = next(gen(b, d, 4321, unit=False, aniso=A))
        search_case(f"c3 mahalanobis top-10 B={b}", ix, n, d, q, 10, 10 if b < 4096 else 5)
    ix.close()

if "c4" in only:  # Euclidean top-100 over 10M x 768 (one GPU holds all 15.4 GB here)
    n, d = int(10_000_000 * S), 768
    ix = lrb.ExactIndex(d, n, metric="euclidean")
    for x in gen(n, d, 1234, unit=False, chunk=500_000):
        ix.add(x)
    for b in (1, 64, 4096):
        q = next(gen(b, d, 4321, unit=False))
        for k in (10, 100):
            if b == 4096 and k == 100 and S >= 1.0:
                iters = 2
            else:
                iters = 5
            search_case(f"c4 euclidean top-{k} B={b}", ix, n, d, q, k, iters)
    ix.close()

if "c5" in only:  # one rank's shard of config 5: 12.5M x 384, batch sweep in powers of 4
    n, d = int(12_500_000 * S), 384
    ix = lrb.ExactIndex(d, n, metric="cosine")
    for x in gen(n, d, 1234):
        ix.add(x.to(torch.bfloat16))
    for b in (1, 4, 16, 64, 256, 1024, 4096):
        q = next(gen(b, d, 4321))
        search_case(f"c5 shard cosine top-10 B={b}", ix, n, d, q, 10, 10 if b < 1024 else 5)
    ix.close()
