"""Throughput of the sentence-encoder forward (all-MiniLM-L6-v2 architecture, seeded weights) on one
B200, beside the CPU oracle (the reference's arithmetic: transformers BertModel in fp32) on a bounded sample.

    python tools/bench_sbert.py [--sentences 16384] [--seq 128] [--precision fp32|bf16] [--cpu-sentences 64]

One JSON line per configuration: sentences/s and tokens/s (CUDA events around encode_tokens with
device-resident token ids, median of --iters), algorithmic TFLOP/s (linear layers 2*K*N per token and
layer + attention 4*S*hidden per token and layer) and its share of the measured sustained bf16 peak
(MEASURED_PEAKS.json; with fp32-level operands every product costs three MMAs, so a third of the peak
is the ceiling), and the CPU oracle's sentences/s on --cpu-sentences sentences of the same shape.
"""
import argparse
import json
import os
import statistics
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import latent_rag_b200 as lrb  # noqa: E402
import oracle  # noqa: E402
from tests.golden import inputs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--sentences", type=int, default=16384)
ap.add_argument("--seq", type=int, default=128)
ap.add_argument("--precision", default="fp32")
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--cpu-sentences", type=int, default=64)
a = ap.parse_args()

cfg = oracle.MINILM_L6
w = inputs.sbert_weights(cfg)
enc = lrb.SentenceEncoder(w, heads=cfg["heads"], precision=a.precision)
g = torch.Generator().manual_seed(3)
ids = torch.randint(0, cfg["vocab"], (a.sentences, a.seq), generator=g)
mask = torch.ones_like(ids)  # full-length sentences: every token is work
ids_d, mask_d = ids.cuda().to(torch.int32), mask.cuda().to(torch.int32)
for _ in range(2):
    enc.encode_tokens(ids_d, mask_d)
enc.check()
times = []
for _ in range(a.iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = enc.encode_tokens(ids_d, mask_d)
    e1.record()
    torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1) * 1e-3)
enc.check()
t = statistics.median(times)
h, f, L = cfg["hidden"], cfg["ffn"], cfg["layers"]
flop_tok = L * (2 * (h * 3 * h + h * h + 2 * h * f) + 4 * a.seq * h)
tokens = a.sentences * a.seq
peak = 1404.1
try:
    with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
        peak = json.load(fh).get("bf16_tflops_sustained", peak)
except OSError:
    pass
line = {"workload": f"all-MiniLM-L6-v2 forward, {a.sentences} x {a.seq} tokens, {a.precision} operands",
        "sentences_per_s": a.sentences / t, "tokens_per_s": tokens / t, "ms": t * 1e3,
        "tflops_algorithmic": flop_tok * tokens / t / 1e12, "peak_tflops_bf16_sustained": peak,
        "frac_of_peak": flop_tok * tokens / t / 1e12 / peak,
        "mma_per_product": 3 if a.precision == "fp32" else 1}
if a.cpu_sentences > 0:
    torch.set_num_threads(os.cpu_count() or 1)
    n = min(a.cpu_sentences, a.sentences)
    oracle.sbert_encode(w, cfg, ids[:8], mask[:8])
    t0 = time.perf_counter()
    ref = oracle.sbert_encode(w, cfg, ids[:n], mask[:n])
    tc = time.perf_counter() - t0
    line["cpu_oracle"] = {"sentences_per_s": n / tc, "cores": torch.get_num_threads(), "sample": f"{n} x {a.seq} tokens"}
    line["max_abs_err_vs_cpu"] = float((out[:n].cpu() - ref).abs().max())
print(json.dumps(line), flush=True)
