"""AE encoder timing: python tools/prof_ae.py [--rows 1000000] [--kernel umma]"""
import argparse, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import latent_rag_b200 as lrb
ap = argparse.ArgumentParser(); ap.add_argument("--rows", type=int, default=1_000_000); ap.add_argument("--kernel", default="umma"); ap.add_argument("--iters", type=int, default=5); ap.add_argument("--precision", default="fp32")
a = ap.parse_args()
gold = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
ae = lrb.load_autoencoder("cae", os.path.join(gold, "ae_weights_cae.npz"), device=0).set_kernel(a.kernel).set_precision(a.precision)
x = torch.randn((a.rows, 384), device="cuda"); x /= x.norm(dim=1, keepdim=True)
ae.encode(x[:4096]); torch.cuda.synchronize()
for it in range(a.iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); z = ae.encode(x); e1.record(); torch.cuda.synchronize()
    print(f"iter {it}: {e0.elapsed_time(e1):.3f} ms  ({a.rows / e0.elapsed_time(e1) / 1e3:.1f} M vec/s)", flush=True)
