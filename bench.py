#!/usr/bin/env python
"""Headline benchmark: exact top-10 QPS over a 100M x 384 bf16 corpus (BASELINE.json config 5),
row-sharded over N B200s (strong scaling: the corpus is fixed, each rank holds 1/N of it).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step is one `search` of a 4096-query batch (the tensor-core regime); the same run also
measures batch-1 search (the HBM-streaming regime) and reports both rooflines.  Synthetic
data: seeded unit-norm Gaussian rows generated on the device, rounded to bf16; 1/8 of the
queries are perturbed corpus rows whose true neighbour is known, which is asserted after
the timed region.  Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

CHUNK = 1_000_000
SEED_CORPUS, SEED_QUERIES = 1234, 4321
METRIC_NAME = "exact top-10 QPS, 100M x384 bf16 corpus, cosine"
UNIT = "queries/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=int(os.environ.get("LK_BENCH_ROWS", 100_000_000)))
    ap.add_argument("--dim", type=int, default=384)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--batch1-steps", type=int, default=20)
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="multi-GPU candidate exchange: fused peer-to-peer kernel, or NCCL all-gather + merge")
    ap.add_argument("--cpu-sample-rows", type=int, default=1_000_000)
    ap.add_argument("--cpu-sample-queries", type=int, default=1024)
    return ap.parse_args()


def traffic_ratios():
    """measured DRAM bytes / algorithmic bytes of the search kernel (ncu, profiles/r01_traffic.json)"""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            t = json.load(f)
        return float(t["batch1"]["ratio"]), float(t["batch4096"]["ratio"])
    except Exception:
        return None, None


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm_gbs": float(p["hbm_gbs"]), "bf16_tflops": float(p["bf16_tflops"]),
                "bf16_tflops_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "src": "measured"}
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "src": "fallback"}


# ---------------------------------------------------------------------------------------
# synthetic data (identical whatever the world size)
# ---------------------------------------------------------------------------------------
def corpus_chunk(c: int, rows: int, dim: int, device) -> torch.Tensor:
    """Global rows [c*CHUNK, c*CHUNK+rows): unit-norm Gaussian, rounded to bf16."""
    g = torch.Generator(device=device).manual_seed(SEED_CORPUS + c)
    x = torch.randn((rows, dim), generator=g, device=device, dtype=torch.float32)
    x = x / x.norm(dim=1, keepdim=True)
    return x.to(torch.bfloat16)


def make_queries(batch: int, dim: int, rows_total: int, device):
    """[batch, dim] fp32 on the host (bf16-representable); every 8th query is a perturbed
    corpus row from chunk 0.  Returns (queries_cpu, planted query positions, planted row ids)."""
    g = torch.Generator().manual_seed(SEED_QUERIES)
    q = torch.randn((batch, dim), generator=g, dtype=torch.float32)
    q = q / q.norm(dim=1, keepdim=True)
    c0_rows = min(CHUNK, rows_total)
    c0 = corpus_chunk(0, c0_rows, dim, device)
    qpos = torch.arange(0, batch, 8)
    rows = (qpos * 7919) % c0_rows
    q[qpos] = c0[rows.to(device)].float().cpu() + 0.1 * q[qpos]
    q = q.to(torch.bfloat16).to(torch.float32)
    return q, qpos.numpy(), rows.numpy()


# ---------------------------------------------------------------------------------------
# the reference's CPU path (oracle port), bounded sample, extrapolated linearly in N
# ---------------------------------------------------------------------------------------
def cpu_baseline(args, steps: int, warmup: int):
    import oracle

    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    n_s = min(args.cpu_sample_rows, args.rows)
    b_s = min(args.cpu_sample_queries, args.batch)
    g = torch.Generator().manual_seed(SEED_CORPUS)
    emb = torch.randn((n_s, args.dim), generator=g)
    emb = oracle.bf16_round(emb / emb.norm(dim=1, keepdim=True))
    q = torch.randn((b_s, args.dim), generator=torch.Generator().manual_seed(SEED_QUERIES))
    q = oracle.bf16_round(q / q.norm(dim=1, keepdim=True))
    index = oracle.bruteforce_build(emb, "cosine")  # retrieval/bruteforce.py:49-50
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        oracle.bruteforce_search(index, q, args.k, "cosine")  # retrieval/bruteforce.py:58-83
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    t = statistics.median(times)
    qps_sample = b_s / t
    qps_full = qps_sample * n_s / args.rows
    sample = (f"{b_s} queries x {n_s} rows x {args.dim} fp32 (torch CPU mm + topk, the reference's "
              f"retrieval/bruteforce.py arithmetic); QPS scaled linearly in N to {args.rows} rows (extrapolated)")
    return {"value": qps_full, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
            "sample_qps": qps_sample, "sample_s_per_step": t}


def run_reference(args, rank: int):
    if rank != 0:
        return
    base = cpu_baseline(args, max(1, args.steps), max(1, min(args.warmup, 2)))
    line = {
        "impl": "reference", "metric": METRIC_NAME, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": base["sample_s_per_step"] * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.gpus), "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {
        "workload": f"config 5: cosine top-{args.k} over {args.rows} x {args.dim} bf16 rows, "
                    f"{args.batch}-query batches (+ batch-1 sweep point), row-sharded over {world} GPU(s)",
        "rows": args.rows, "dim": args.dim, "batch": args.batch, "k": args.k, "rows_per_gpu": -(-args.rows // world),
        "parallelism": (f"row-shard x{world} + candidate exchange fused with the k-merge "
                        f"({'peer-to-peer stores over NVLink' if args.exchange == 'p2p' else 'NCCL all-gather'})"
                        if world > 1 else "one GPU holds the whole corpus"),
        "l2": "inputs larger than L2 (each rank streams its whole corpus shard every step)",
    }


# ---------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu: int):
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.strip().splitlines():
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); pw.append(float(parts[2]))
            except ValueError:
                continue
            for name, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus and world > 1:
        args.gpus = world

    import latent_rag_b200 as lrb

    lrb._native.require_device()
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    # ---- build this rank's shard on the device -------------------------------------
    lo, hi = lrb.shard_bounds(args.rows, world)[rank]
    t_build = time.perf_counter()
    index = lrb.ExactIndex(args.dim, max(1, hi - lo), metric="cosine", storage="bf16", device=local_rank)
    for c in range(lo // CHUNK, -(-hi // CHUNK)):
        c_lo, c_hi = c * CHUNK, min((c + 1) * CHUNK, args.rows)
        chunk = corpus_chunk(c, c_hi - c_lo, args.dim, dev)
        index.add(chunk[max(lo, c_lo) - c_lo : min(hi, c_hi) - c_lo])
        del chunk
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t_build
    assert index.size == hi - lo

    q_host, qpos, planted = make_queries(args.batch, args.dim, args.rows, dev)
    q_pin = q_host.pin_memory()
    q_dev = q_host.to(dev)
    k = args.k

    xchg = None
    if world > 1 and args.exchange == "p2p":
        xchg = lrb.PeerExchange(local_rank, rank, world, max_b=max(args.batch, 1)).connect()

    def gather_merge(d, i):
        """The one exchange step: the [B,k] candidates of every rank -> the global top-k on every
        rank.  p2p: one kernel per rank (stores into every peer's buffer over NVLink, flags, wait,
        merge); nccl: all-gather + merge kernel."""
        if world == 1:
            return d, i
        if xchg is not None:
            return xchg.exchange_merge(d, i, k)
        b = d.size(0)
        gd = torch.empty((world * b, k), dtype=torch.float32, device=dev)
        gi = torch.empty((world * b, k), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(gd, d)
        dist.all_gather_into_tensor(gi, i)
        return lrb.merge_topk(gd.view(world, b, k).permute(1, 0, 2).contiguous(),
                              gi.view(world, b, k).permute(1, 0, 2).contiguous(), k)

    def step_device(q):
        d, i = index.search(q, k, idx_base=lo, device_out=True)
        return gather_merge(d, i)

    def step_e2e(q_pinned):
        """What a user calls: host queries in, host results out.  One GPU: the C-ABI search with
        HOST buffers (H2D of the queries and D2H of the results inside the call)."""
        if world == 1:
            d, i = index.search(q_pinned, k)
            return torch.from_numpy(d), torch.from_numpy(i)
        d, i = index.search(q_pinned.to(dev, non_blocking=True), k, idx_base=lo, device_out=True)
        d, i = gather_merge(d, i)
        return d.cpu(), i.cpu()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, arg, steps, warmup):
        for _ in range(warmup):
            fn(arg)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        out = None
        for _ in range(steps):
            out = fn(arg)
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = torch.tensor([e0.elapsed_time(e1), wall * 1e3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms[0]) / steps, float(ms[1]) / steps, out

    def kernel_ms(q, steps):
        """average device time of the fused search kernel alone (CUDA events on its stream)."""
        index.set_timing(True)
        ts = []
        for _ in range(steps):
            index.search(q, k, idx_base=lo, device_out=True)
            ts.append(index.last_timing()[0])
        index.set_timing(False)
        t = torch.tensor([sum(ts) / len(ts)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    K, W = args.steps, max(3, args.warmup)
    launches0 = lrb._native.launch_count()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms_dev, _, out = timed(step_device, q_dev, K, W)
    launches = lrb._native.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    _, ms_e2e, out_e2e = timed(step_e2e, q_pin, K, 1)
    kms = kernel_ms(q_dev, K)

    # batch-1 (HBM-streaming regime)
    q1_dev, q1_pin = q_dev[:1].contiguous(), q_pin[:1].clone().pin_memory()
    ms1_dev, _, _ = timed(step_device, q1_dev, args.batch1_steps, W)
    _, ms1_e2e, _ = timed(step_e2e, q1_pin, args.batch1_steps, 1)
    kms1 = kernel_ms(q1_dev, args.batch1_steps)

    # ---- full-size correctness properties (outside the timed regions) -----------------
    index.check()  # no search kernel hit a pipeline timeout
    if xchg is not None:
        xchg.check()  # no rank timed out waiting for a peer's candidates
    d_fin, i_fin = out[0].cpu().numpy(), out[1].cpu().numpy()
    planted_ok = bool((i_fin[qpos, 0] == planted).all())
    sorted_ok = bool((np.diff(d_fin, axis=1) <= 0).all())
    same_e2e = bool((out_e2e[1].numpy() == i_fin).all())

    if rank == 0:
        pk = peaks()
        rows_per_gpu = -(-args.rows // world)
        flops = 2.0 * args.batch * rows_per_gpu * args.dim
        bytes_per_launch = rows_per_gpu * args.dim * 2 + rows_per_gpu * 4  # bf16 rows + fp32 side values
        ach_tf = flops / (kms * 1e-3) / 1e12
        ach_gbs1 = bytes_per_launch / (kms1 * 1e-3) / 1e9
        base = cpu_baseline(args, 3, 1) if world == 1 else None
        r1, r4096 = traffic_ratios()
        line = {
            "metric": METRIC_NAME, "value": args.batch / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16 inputs, f32 accumulate", "data": "synthetic",
            "config": workload_config(args, world),
            "e2e": {"value": args.batch / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(q_pin.numel() * 4), "d2h_bytes_per_step": int(args.batch * k * 12)},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "achieved": ach_tf, "peak": pk["bf16_tflops_sustained"],
                         "unit": "TFLOP/s", "frac": ach_tf / pk["bf16_tflops_sustained"],
                         "traffic": None if r4096 is None else r4096 * bytes_per_launch,
                         "kernel": "umma_search_kernel", "kernel_ms": kms, "flops_per_launch": flops,
                         "peak_src": pk["src"] + " sustained bf16 (kernel runs for hundreds of ms per launch)",
                         "frac_of_burst_peak": ach_tf / pk["bf16_tflops"]},
            "batch1": {"value": 1.0 / (ms1_dev * 1e-3), "unit": UNIT, "ms_per_query": ms1_dev,
                       "e2e": {"value": 1.0 / (ms1_e2e * 1e-3), "unit": UNIT, "ms_per_query": ms1_e2e,
                               "h2d_bytes_per_step": args.dim * 4, "d2h_bytes_per_step": k * 12},
                       "roofline": {"bound": "hbm", "achieved": ach_gbs1, "peak": pk["hbm_gbs"], "unit": "GB/s",
                                    "frac": ach_gbs1 / pk["hbm_gbs"],
                                    "traffic": None if r1 is None else r1 * bytes_per_launch, "kernel_ms": kms1,
                                    "bytes_per_launch": bytes_per_launch, "peak_src": pk["src"]}},
            "clocks": clocks,
            "checks": {"planted_neighbours_found": planted_ok, "scores_sorted": sorted_ok,
                       "e2e_equals_device_path": same_e2e},
            "build_s": build_s,
        }
        if base is not None:
            line["cpu_baseline"] = base
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if not (planted_ok and sorted_ok and same_e2e):
        sys.exit(3)


if __name__ == "__main__":
    main()
