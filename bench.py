#!/usr/bin/env python
"""Headline benchmark: exact top-10 QPS over a 100M x 384 bf16 corpus (BASELINE.json config 5),
row-sharded over N B200s (strong scaling: the corpus is fixed, each rank holds 1/N of it).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--configs all|none|c1,c2,...]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step is one `search` of a 4096-query batch (the tensor-core regime); the same run also
measures batch-1 search (the HBM-streaming regime) and reports both rooflines.  Synthetic
data: seeded unit-norm Gaussian rows generated on the device, rounded to bf16; 1/8 of the
queries are perturbed corpus rows whose true neighbour is known, which is asserted after
the timed region.  Prints ONE JSON line (rank 0).

The same line carries, under "configs", every other BASELINE.json configuration with its own kernel
time, roofline fractions, end-to-end (host buffers) figure and -- on one GPU -- the reference's CPU
arithmetic timed beside it on a bounded sample:
  c1         10k queries x 20k x 384, cosine top-10 (bf16 tcgen05 path and the exact-fp32 path) + the
             one-query-per-call `.retrieve` loop of the reference's caller (main.py:270-271)
  c2_encode  CAE encoder 384 -> 512 -> 64 over 1.01M unit-norm vectors (fp32-level and bf16 operands)
  c2_search  cosine top-10, 10k latent queries x 1M x 64
  c3_b*      Mahalanobis top-10 over 10M x 384 anisotropic rows, batch 1 / 64 / 4096
  c4_b*      Euclidean top-100 over 10M x 768, batch 1 / 64 / 4096 -- on N > 1 GPUs row-sharded through
             ShardedRetriever (the only side config run at N > 1)
under "batch_sweep" the points of config 5's batch sweep between its two ends (4, 16, 64, 256, 1024 queries),
and under "checks" an ORACLE comparison of the sharded data path run by this very process group:
a 1M-row side index sharded over the N ranks against oracle.bruteforce_search on rank 0.
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
AE_FLOPS_PER_VEC = 2 * (384 * 512 + 512 * 64)  # 458,752
AE_BYTES_PER_VEC = 384 * 4 + 64 * 4            # fp32 in, fp32 latents out (the reference's layout): 1,792


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
    ap.add_argument("--configs", default=os.environ.get("LK_BENCH_CONFIGS", "all"),
                    help="side configurations to measure: all | none | comma list of c1,c2,c3,c4")
    ap.add_argument("--config-scale", type=float, default=float(os.environ.get("LK_BENCH_CONFIG_SCALE", 1.0)),
                    help="scale the corpora of the side configurations (smoke runs)")
    ap.add_argument("--cpu-rows", type=int, default=10_000_000,
                    help="rows of the CPU-baseline sample (BASELINE.md section 3: N' = 10M when 100M x 384 fp32 "
                         "does not fit host memory); shrunk when the box has less memory")
    ap.add_argument("--cpu-queries", type=int, default=64, help="queries per step of the CPU baseline")
    return ap.parse_args()


def traffic_ratios():
    """measured DRAM bytes / algorithmic bytes of the search kernel (ncu, profiles/r02_traffic.json)"""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                t = json.load(f)
            return float(t["batch1"]["ratio"]), float(t["batch4096"]["ratio"])
        except Exception:
            continue
    return None, None


TRAFFIC_SRC = ("estimated: this shape's algorithmic bytes x the DRAM/algorithmic ratio ncu measured for the same "
               "kernel on a 10M-row (batch 1) / 4M-row (batch 4096) run, profiles/r02_traffic.json; not captured "
               "in this run")


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
def corpus_chunk(c: int, rows: int, dim: int, device, unit: bool = True, aniso=None, seed: int = SEED_CORPUS):
    """Global rows [c*CHUNK, c*CHUNK+rows): Gaussian (unit-norm unless `unit` is False; x A when `aniso`
    is given), rounded to bf16."""
    g = torch.Generator(device=device).manual_seed(seed + c)
    x = torch.randn((rows, dim), generator=g, device=device, dtype=torch.float32)
    if aniso is not None:
        x = x @ aniso
    if unit:
        x = x / x.norm(dim=1, keepdim=True)
    return x.to(torch.bfloat16)


def make_queries(batch: int, dim: int, rows_total: int, device, unit: bool = True, aniso=None, seed: int = SEED_CORPUS):
    """[batch, dim] fp32 on the host (bf16-representable); every 8th query is a perturbed
    corpus row from chunk 0.  Returns (queries_cpu, planted query positions, planted row ids)."""
    g = torch.Generator().manual_seed(SEED_QUERIES)
    q = torch.randn((batch, dim), generator=g, dtype=torch.float32)
    if aniso is not None:
        q = q @ aniso.cpu()
    if unit:
        q = q / q.norm(dim=1, keepdim=True)
    c0_rows = min(CHUNK, rows_total)
    c0 = corpus_chunk(0, c0_rows, dim, device, unit=unit, aniso=aniso, seed=seed)
    qpos = torch.arange(0, batch, 8)
    rows = (qpos * 7919) % c0_rows
    q[qpos] = c0[rows.to(device)].float().cpu() + 0.1 * q[qpos]
    q = q.to(torch.bfloat16).to(torch.float32)
    return q, qpos.numpy(), rows.numpy()


def build_shard(lrb, rows_total, dim, lo, hi, metric, dev, unit=True, aniso=None, seed=SEED_CORPUS, whiten=None):
    """This rank's rows [lo, hi) of the seeded corpus, added chunk by chunk on the device."""
    index = lrb.ExactIndex(dim, max(1, hi - lo), metric=metric, storage="bf16", device=dev.index, whiten=whiten)
    for c in range(lo // CHUNK, -(-hi // CHUNK)):
        c_lo, c_hi = c * CHUNK, min((c + 1) * CHUNK, rows_total)
        chunk = corpus_chunk(c, c_hi - c_lo, dim, dev, unit=unit, aniso=aniso, seed=seed)
        index.add(chunk[max(lo, c_lo) - c_lo: min(hi, c_hi) - c_lo])
        del chunk
    torch.cuda.synchronize(dev)
    assert index.size == hi - lo
    return index


# ---------------------------------------------------------------------------------------
# the reference's CPU path (oracle port of retrieval/bruteforce.py), bounded samples
# ---------------------------------------------------------------------------------------
class CpuSample:
    """One host buffer of bf16-representable Gaussian rows, viewed at whatever width a configuration
    needs (the CPU timings do not depend on the values).  Filled through the GPU when one is visible
    (torch's CPU normal generator needs ~4.5 s per million 384-d rows), by numpy threads otherwise."""

    def __init__(self, rows: int, dim: int = 384):
        try:
            import psutil

            avail = psutil.virtual_memory().available
        except Exception:
            avail = 32 << 30
        # the buffer + one normalised copy (cosine build) + 8 GB of score blocks must fit with headroom
        while rows > 1_000_000 and rows * dim * 4 * 2 + (10 << 30) > 0.6 * avail:
            rows //= 2
        self.rows, self.dim = rows, dim
        t0 = time.perf_counter()
        self.buf = torch.empty((rows, dim), dtype=torch.float32)
        if torch.cuda.is_available():
            dev = torch.device("cuda:0")
            for c, lo in enumerate(range(0, rows, CHUNK)):
                n = min(CHUNK, rows - lo)
                g = torch.Generator(device=dev).manual_seed(SEED_CORPUS + c)
                x = torch.randn((n, dim), generator=g, device=dev).to(torch.bfloat16).to(torch.float32)
                self.buf[lo:lo + n] = x.cpu()
        else:
            from concurrent.futures import ThreadPoolExecutor

            def fill(c_lo):
                c, lo = c_lo
                n = min(CHUNK, rows - lo)
                x = np.random.Generator(np.random.SFC64(SEED_CORPUS + c)).standard_normal((n, dim), dtype=np.float32)
                self.buf[lo:lo + n] = torch.from_numpy(x).to(torch.bfloat16).to(torch.float32)

            with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
                list(ex.map(fill, list(enumerate(range(0, rows, CHUNK)))))
        self.fill_s = time.perf_counter() - t0

    def view(self, dim: int, rows: int | None = None) -> torch.Tensor:
        v = self.buf.view(-1, dim)
        return v if rows is None else v[:rows]

    def queries(self, b: int, dim: int) -> torch.Tensor:
        q = torch.randn((b, dim), generator=torch.Generator().manual_seed(SEED_QUERIES))
        return q.to(torch.bfloat16).to(torch.float32)


def cpu_search_qps(emb_built, q, k, metric, steps, warmup, rows_full, chunk_bytes=8 << 30):
    """Wall-clock QPS of oracle.bruteforce_search (retrieval/bruteforce.py:58-83) on the sample, and
    the same scaled linearly in N to `rows_full` rows."""
    import oracle

    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        oracle.bruteforce_search(emb_built, q, k, metric, chunk_bytes=chunk_bytes)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    t = statistics.median(times)
    n_s = emb_built.size(0)
    return {"sample_qps": q.size(0) / t, "sample_s_per_step": t,
            "value": q.size(0) / t * min(1.0, n_s / rows_full), "rows": n_s, "queries": q.size(0)}


def cpu_baseline_c5(args, sample: CpuSample, steps: int, warmup: int):
    import oracle

    cores = torch.get_num_threads()
    emb = oracle.bruteforce_build(sample.view(args.dim), "cosine")  # retrieval/bruteforce.py:49-50
    q = sample.queries(min(args.cpu_queries, args.batch), args.dim)
    r = cpu_search_qps(emb, q, args.k, "cosine", steps, warmup, args.rows)
    extrap = "" if r["rows"] >= args.rows else f"; QPS scaled linearly in N to {args.rows} rows (extrapolated)"
    return {"value": r["value"], "unit": UNIT, "cores": cores, "kind": "port",
            "sample": (f"{r['queries']} queries x {r['rows']} rows x {args.dim} fp32 per step, torch CPU mm + topk in "
                       f"8 GB score blocks = the reference's retrieval/bruteforce.py:58-83 arithmetic{extrap}"),
            "sample_qps": r["sample_qps"], "sample_s_per_step": r["sample_s_per_step"]}


def run_reference(args, rank: int):
    """The reference arm: the oracle port of retrieval/bruteforce.py on the host cores, config 5's
    workload, a bounded sample per step (rank 0 only; other ranks exit without work)."""
    if rank != 0:
        return
    import oracle

    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    sample = CpuSample(min(args.cpu_rows, args.rows), args.dim)
    emb = oracle.bruteforce_build(sample.view(args.dim), "cosine")
    nq = min(args.cpu_queries, args.batch)
    K, W = max(1, args.steps), max(1, args.warmup)
    # size the step so that the whole K + W run ends within a few minutes
    q = sample.queries(nq, args.dim)
    t0 = time.perf_counter()
    oracle.bruteforce_search(emb, q, args.k, "cosine", chunk_bytes=8 << 30)
    t1 = time.perf_counter() - t0
    while nq > 8 and t1 * (K + W) * nq / q.size(0) > 150.0:
        nq //= 2
    q = q[:nq].contiguous()
    r = cpu_search_qps(emb, q, args.k, "cosine", K, W, args.rows)
    extrap = "" if r["rows"] >= args.rows else f"; QPS scaled linearly in N to {args.rows} rows (extrapolated)"
    base = {"value": r["value"], "unit": UNIT, "cores": cores, "kind": "port",
            "sample": (f"{nq} queries x {r['rows']} rows x {args.dim} fp32 per step, torch CPU mm + topk in 8 GB score "
                       f"blocks = the reference's retrieval/bruteforce.py:58-83 arithmetic (oracle port; FAISS-CPU is "
                       f"not in the image){extrap}"),
            "sample_qps": r["sample_qps"], "sample_s_per_step": r["sample_s_per_step"]}
    line = {
        "impl": "reference", "metric": METRIC_NAME, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": K, "warmup": W, "ms_per_step": base["sample_s_per_step"] * 1e3,
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


# ---------------------------------------------------------------------------------------
# the process-group context every measurement shares
# ---------------------------------------------------------------------------------------
class Ctx:
    def __init__(self, args, lrb, rank, world, local_rank):
        self.args, self.lrb, self.rank, self.world, self.local_rank = args, lrb, rank, world, local_rank
        self.dev = torch.device(f"cuda:{local_rank}")
        self.pk = peaks()
        self.keep = []

    def barrier(self):
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, *vals):
        t = torch.tensor(list(vals), dtype=torch.float64, device=self.dev)
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def all_ranks(self, val):
        t = torch.tensor([val], dtype=torch.float64, device=self.dev)
        if self.world == 1:
            return [float(val)]
        out = torch.empty(self.world, dtype=torch.float64, device=self.dev)
        dist.all_gather_into_tensor(out, t)
        return [float(v) for v in out]

    def timed(self, fn, arg, steps, warmup, warm_s=0.0):
        """(device ms per step by CUDA events on the launching stream, wall ms per step, last result);
        barrier + synchronize on both sides, max over ranks.  warm_s (one GPU only: the step count has to be the
        same on every rank of a sharded search): keep stepping, untimed, for that long after the `warmup` calls -- a
        sub-millisecond step does not bring the clocks back up after the GPU sat idle through a CPU baseline (config 1
        measured 0.47 .. 0.87 ms per step against 0.36 ms once warm)."""
        for _ in range(warmup):
            fn(arg)
        if warm_s > 0 and self.world == 1:
            t_end = time.perf_counter() + warm_s
            while time.perf_counter() < t_end:
                for _ in range(10):
                    fn(arg)
                torch.cuda.synchronize()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        out = None
        for _ in range(steps):
            out = fn(arg)
        e1.record()
        self.barrier()
        wall = time.perf_counter() - t0
        ms_dev, ms_wall = self.max_over_ranks(e0.elapsed_time(e1), wall * 1e3)
        return ms_dev / steps, ms_wall / steps, out

    def kernel_ms(self, index, q, k, lo, steps):
        """average device time of the fused search kernel alone (CUDA events on its stream): (max over
        ranks, per-rank list)."""
        index.set_timing(True)
        ts = []
        for _ in range(steps):
            index.search(q, k, idx_base=lo, device_out=True)
            ts.append(index.last_timing()[0])
        index.set_timing(False)
        mine = sum(ts) / len(ts)
        per_rank = self.all_ranks(mine)
        return max(per_rank), per_rank


class SearchCase:
    """One (index shard, metric, k) searched the way a user of the package does: on one GPU through
    ExactIndex.search (the C ABI with host or device buffers), on several through ShardedRetriever."""

    def __init__(self, ctx: Ctx, index, lo: int, metric: str, max_batch: int):
        self.ctx, self.index, self.lo = ctx, index, lo
        self.sharded = None
        if ctx.world > 1:
            self.sharded = ctx.lrb.ShardedRetriever(index, lo, metric, exchange=ctx.args.exchange, max_batch=max_batch)
            ctx.keep.append(self.sharded)  # peers keep this rank's exchange buffer mapped until the process ends

    def step_device(self, q, k):
        if self.sharded is None:
            return self.index.search(q, k, idx_base=self.lo, device_out=True)
        return self.sharded.search_tensors(q, k)

    def step_e2e(self, q_pinned, k):
        """What a user calls: host queries in, host results out (H2D of the queries and D2H of the
        results inside the call)."""
        if self.sharded is None:
            d, i = self.index.search(q_pinned, k)
            return torch.from_numpy(d), torch.from_numpy(i)
        d, i = self.sharded.search(q_pinned.to(self.ctx.dev, non_blocking=True), k)
        return torch.from_numpy(d), torch.from_numpy(i)

    def check(self):
        if self.sharded is not None:
            self.sharded.check()
        else:
            self.index.check()

    def measure(self, q_host, k, steps, warmup, rows_per_gpu, dim, side_bytes=4, warm_s=0.0):
        """kernel / device-step / end-to-end times of one batch + both roofline fractions."""
        ctx = self.ctx
        q_pin = q_host.pin_memory()
        q_dev = q_host.to(ctx.dev)
        b = q_host.size(0)
        step = lambda q: self.step_device(q, k)  # noqa: E731
        if warm_s > 0 and ctx.world == 1 and steps >= 10:
            # sub-millisecond steps are three to five launches each and a stall of the submitting thread shows up as idle
            # GPU time (config 1 measured 0.36 .. 0.87 ms per step from run to run with an unchanged 0.314 ms kernel): the
            # MEDIAN of five timed groups of steps / 5 ("median of >= 20"), not one group's mean
            per = max(2, steps // 5)
            first, _, out = ctx.timed(step, q_dev, per, warmup, warm_s)
            ms_dev = statistics.median([first] + [ctx.timed(step, q_dev, per, 0)[0] for _ in range(4)])
        else:
            ms_dev, _, out = ctx.timed(step, q_dev, steps, warmup, warm_s)
        _, ms_e2e, out_e2e = ctx.timed(lambda q: self.step_e2e(q, k), q_pin, max(2, steps // 2), 1)
        if warm_s > 0 and ctx.world == 1 and steps >= 50:  # (same reason: the median of five groups)
            ms_e2e = statistics.median([ms_e2e] + [ctx.timed(lambda q: self.step_e2e(q, k), q_pin, steps // 2, 0)[1]
                                                   for _ in range(4)])
        kms, per_rank = ctx.kernel_ms(self.index, q_dev, k, self.lo, max(2, steps // 2))
        self.check()
        flops = 2.0 * b * rows_per_gpu * dim
        byts = rows_per_gpu * dim * 2 + rows_per_gpu * side_bytes
        tf, gbs = flops / (kms * 1e-3) / 1e12, byts / (kms * 1e-3) / 1e9
        bound = "hbm" if b <= 128 else "tensor"
        rec = {
            "batch": b, "k": k, "qps": b / (ms_dev * 1e-3), "ms_per_step": ms_dev, "kernel_ms": kms,
            "e2e": {"value": b / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(q_host.numel() * 4), "d2h_bytes_per_step": int(b * k * 12)},
            "roofline": {"bound": bound, "achieved": gbs if bound == "hbm" else tf,
                         "peak": ctx.pk["hbm_gbs"] if bound == "hbm" else ctx.pk["bf16_tflops_sustained"],
                         "unit": "GB/s" if bound == "hbm" else "TFLOP/s",
                         "frac": gbs / ctx.pk["hbm_gbs"] if bound == "hbm" else tf / ctx.pk["bf16_tflops_sustained"],
                         "tensor_frac_of_burst_peak": tf / ctx.pk["bf16_tflops"],
                         "tflops": tf, "gbs_one_pass": gbs, "flops_per_launch": flops, "bytes_per_launch": byts,
                         "traffic": None},
        }
        if ctx.world > 1:
            rec["kernel_ms_per_rank"] = per_rank
        return rec, out, out_e2e


def oracle_parity_sharded(ctx: Ctx):
    """Oracle equality of the multi-GPU data path, run by this very process group: a 1M x 384 side
    corpus (cosine top-10) and a 250k x 768 one (euclidean top-100, config 4's selector) sharded over the
    ranks, 256 queries, against oracle.bruteforce_search (retrieval/bruteforce.py:58-83) on rank 0."""
    import oracle

    lrb, dev, world, rank = ctx.lrb, ctx.dev, ctx.world, ctx.rank
    res = {}
    for name, rows, dim, metric, k, unit in (("cosine_top10_1Mx384", 1_000_000, 384, "cosine", 10, True),
                                             ("euclidean_top100_250kx768", 250_000, 768, "euclidean", 100, False)):
        lo, hi = lrb.shard_bounds(rows, world)[rank]
        index = build_shard(lrb, rows, dim, lo, hi, metric, dev, unit=unit, seed=777)
        case = SearchCase(ctx, index, lo, metric, 256)
        q_host, qpos, planted = make_queries(256, dim, rows, dev, unit=unit, seed=777)
        d, i = case.step_device(q_host.to(dev), k)
        d1, i1 = case.step_device(q_host[:1].to(dev), k)  # the batch-1 route (single CTA per group, select merge)
        case.check()
        d, i, i1 = d.cpu().numpy(), i.cpu().numpy(), i1.cpu().numpy()
        ok, why = True, ""
        if rank == 0:
            emb = torch.cat([corpus_chunk(c, min(CHUNK, rows - c * CHUNK), dim, dev, unit=unit, seed=777).float().cpu()
                             for c in range(-(-rows // CHUNK))])
            d_ref, i_ref = oracle.bruteforce_search(oracle.bruteforce_build(emb, metric), q_host, k, metric)
            scale = oracle.euclidean_scale(emb, q_host) if metric == "euclidean" else None
            ok, why = oracle.topk_equivalent(d_ref, i_ref, d, i, rtol=1e-5, scale=scale)
            ok = bool(ok and (i[qpos, 0] == planted).all() and (i1[0] == i[0]).all())
        flag = ctx.max_over_ranks(0.0 if ok else 1.0)[0]
        res[name] = flag == 0.0
        if why and rank == 0:
            res[name + "_why"] = why
        index.close()
    return res


# ---------------------------------------------------------------------------------------
# side configurations (BASELINE.json configs 1-4)
# ---------------------------------------------------------------------------------------
def cfg_c1(ctx: Ctx, sample):
    """config 1: 10k queries x 20k docs x 384, cosine top-10 -- the reference's own CPU-runnable case."""
    import oracle

    lrb, dev = ctx.lrb, ctx.dev
    n, b, d, k = 20_000, 10_000, 384, 10
    emb = corpus_chunk(0, n, d, dev)
    q_host, qpos, planted = make_queries(b, d, n, dev)
    out = {"workload": f"cosine top-{k}, {b} queries x {n} x {d}"}
    for prec in ("bf16", "fp32"):
        index = lrb.ExactIndex(d, n, metric="cosine", storage=prec, device=dev.index)
        index.add(emb)
        case = SearchCase(ctx, index, 0, "cosine", b)
        # (30 warm-up calls: the GPU sat idle during the CPU baseline of the headline and a 0.4 ms step does not ramp
        #  the clocks by itself)
        # (100 timed steps: 3-5 launches per call, and a stream's backlog of ~1000 pending launches stalls the submitting
        #  thread -- 200 steps of the fp32 path measured 1.24 ms per step against 0.45 with 40)
        rec, res, _ = case.measure(q_host, k, 100, 30, n, d, warm_s=0.3)
        rec["planted_neighbours_found"] = bool((res[1].cpu().numpy()[qpos, 0] == planted).all())
        out[prec] = rec
        if prec == "bf16":  # the reference caller's loop: one retrieve() per query (main.py:270-271)
            r = lrb.BruteForceRetriever(emb, [""] * n, None, metric="cosine", device=dev.index)
            qs = [q_host[t] for t in range(200)]
            for t in range(20):
                r.retrieve(qs[t], top_k=k)
            t0 = time.perf_counter()
            for t in range(200):
                r.retrieve(qs[t], top_k=k)
            out["retrieve_loop"] = {"us_per_query": (time.perf_counter() - t0) / 200 * 1e6,
                                    "what": "BruteForceRetriever.retrieve(q, top_k=10), one host query per call, "
                                            "200 calls (single-launch small-batch kernel)"}
        index.close()
    if sample is not None:
        embc = oracle.bruteforce_build(emb.float().cpu(), "cosine")
        r = cpu_search_qps(embc, q_host, k, "cosine", 3, 1, n)
        out["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                               "sample": f"the whole workload: {b} queries x {n} rows, oracle port of bruteforce.py:58-83"}
        t0 = time.perf_counter()
        for t in range(100):
            oracle.bruteforce_search(embc, q_host[t], k, "cosine")
        out["retrieve_loop"]["cpu_us_per_query"] = (time.perf_counter() - t0) / 100 * 1e6
    return {"c1": out}


def cfg_c2(ctx: Ctx, sample, scale):
    """config 2: CAE latent corpus encode (384 -> 512 -> 64) + cosine top-10, 1M docs x 10k queries."""
    import oracle

    lrb, dev, pk = ctx.lrb, ctx.dev, ctx.pk
    n, b = int(1_000_000 * scale), 10_000
    m = n + b
    gold = os.path.join(ROOT, "tests", "golden")
    ae = lrb.load_autoencoder("cae", os.path.join(gold, "ae_weights_cae.npz"), device=dev.index)
    g = torch.Generator(device=dev).manual_seed(SEED_CORPUS)
    x = torch.randn((m, 384), generator=g, device=dev)
    x = x / x.norm(dim=1, keepdim=True)
    enc = {}
    z = None
    for prec in ("fp32", "bf16"):
        ae.set_precision(prec)
        ae.encode(x[:4096])
        torch.cuda.synchronize()
        ts = []
        for _ in range(7):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            zz = ae.encode(x)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = statistics.median(ts)
        if prec == "fp32":
            z = zz
        tf, gbs = m * AE_FLOPS_PER_VEC / (ms * 1e-3) / 1e12, m * AE_BYTES_PER_VEC / (ms * 1e-3) / 1e9
        enc[prec] = {"vectors": m, "ms": ms, "vectors_per_s": m / (ms * 1e-3),
                     "kernel": "split_rows_kernel + ae_umma_kernel" if prec == "fp32" else "ae_pair_kernel",
                     "operands": "split-bf16 (hi + lo planes, 3 MMAs per product, fp32-level results: 3x the tensor work "
                                 "by construction)" if prec == "fp32"
                     else "bf16 inputs / weights / hidden activations, fp32 accumulate (1 MMA per product; north_star's "
                          "stated precision); CTA pairs, fp32 rows converted in the kernel",
                     "roofline": {"bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s",
                                  "frac": gbs / pk["hbm_gbs"], "tflops": tf,
                                  "tensor_frac_of_burst_peak": tf / pk["bf16_tflops"],
                                  "bytes_per_vector": AE_BYTES_PER_VEC, "flops_per_vector": AE_FLOPS_PER_VEC,
                                  "traffic": None}}
    # end to end: pinned host rows in, host latents out (bounded: 262,144 rows)
    m_e = min(m, 262_144)
    xh = x[:m_e].cpu().pin_memory()
    ae.set_precision("fp32")
    ae.encode(xh)  # (the first call of a size allocates the staging buffers and the page-locked result block)
    t0 = time.perf_counter()
    for _ in range(5):
        ae.encode(xh)
    e2e_s = (time.perf_counter() - t0) / 5
    enc["e2e"] = {"value": m_e / e2e_s, "unit": "vectors/s", "vectors": m_e, "h2d_bytes_per_step": m_e * 384 * 4,
                  "d2h_bytes_per_step": m_e * 64 * 4}
    if sample is not None:
        w = oracle.load_encoder_weights(np.load(os.path.join(gold, "ae_weights_cae.npz")), "cae")
        xs = x[:200_000].cpu()
        oracle.ae_encode(xs[:10_000], w, "cae")
        t0 = time.perf_counter()
        for _ in range(3):
            oracle.ae_encode(xs, w, "cae")
        t = (time.perf_counter() - t0) / 3
        enc["cpu_baseline"] = {"value": len(xs) / t, "unit": "vectors/s", "cores": torch.get_num_threads(), "kind": "port",
                               "sample": "200,000 x 384 fp32 vectors, oracle port of ContrastiveAutoencoder.encode "
                                         "(torch CPU linear + relu + linear + normalize)"}
    # latent search
    index = lrb.ExactIndex(64, n, metric="cosine", storage="bf16", device=dev.index)
    index.add(z[:n])
    case = SearchCase(ctx, index, 0, "cosine", b)
    rec, _, _ = case.measure(z[n:].cpu(), 10, 10, 3, n, 64, warm_s=0.2)
    rec["workload"] = f"cosine top-10, {b} latent queries x {n} x 64"
    if sample is not None:
        embc = oracle.bruteforce_build(z[:n].cpu(), "cosine")
        r = cpu_search_qps(embc, z[n:n + 2048].cpu(), 10, "cosine", 2, 1, n)
        rec["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                               "sample": f"2048 queries x {n} rows x 64 fp32 (the encoder's own latents), oracle port of "
                                         "bruteforce.py:58-83"}
    index.close()
    return {"c2_encode": enc, "c2_search": rec}


def aniso_matrix(d, dev):
    """A = diag(linspace(0.2, 2, d)) R, R a seeded random orthogonal."""
    g = torch.Generator(device=dev).manual_seed(7)
    rot = torch.linalg.qr(torch.randn((d, d), generator=g, device=dev))[0]
    return torch.diag(torch.linspace(0.2, 2.0, d, device=dev)) @ rot


def cfg_c3(ctx: Ctx, sample, scale):
    """config 3: Mahalanobis top-10 on 384-d, full covariance, 10M docs, batch 1 / 64 / 4096."""
    import oracle
    from latent_rag_b200.retrieval.common import whitener_from_precision

    lrb, dev = ctx.lrb, ctx.dev
    n, d, k = int(10_000_000 * scale), 384, 10
    A = aniso_matrix(d, dev)
    # the MLE covariance of the corpus itself (sklearn's EmpiricalCovariance), from fp64 moment sums
    t0 = time.perf_counter()
    s1 = torch.zeros(d, dtype=torch.float64, device=dev)
    s2 = torch.zeros((d, d), dtype=torch.float64, device=dev)
    for c in range(-(-n // CHUNK)):
        x = corpus_chunk(c, min(CHUNK, n - c * CHUNK), d, dev, unit=False, aniso=A).to(torch.float64)
        s1 += x.sum(0)
        s2 += x.T @ x
    mean = s1 / n
    cov = s2 / n - torch.outer(mean, mean)
    prec = np.linalg.pinv((0.5 * (cov + cov.T)).cpu().numpy(), hermitian=True)
    index = build_shard(lrb, n, d, 0, n, "mahalanobis", dev, unit=False, aniso=A, whiten=whitener_from_precision(prec))
    build_s = time.perf_counter() - t0
    case = SearchCase(ctx, index, 0, "mahalanobis", 4096)
    out = {}
    for b in (1, 64, 4096):
        q_host, qpos, planted = make_queries(b, d, n, dev, unit=False, aniso=A)
        rec, res, _ = case.measure(q_host, k, 10 if b < 4096 else 4, 3, n, d, warm_s=0.2 if b < 4096 else 0.0)
        rec["planted_neighbours_found"] = bool((res[1].cpu().numpy()[qpos, 0] == planted).all())
        rec["workload"] = f"mahalanobis top-{k}, {b} queries x {n} x {d} (bf16-stored whitened rows), full covariance"
        rec["index_build_s"] = build_s
        out[f"c3_b{b}"] = rec
    index.close()
    # the same corpus with fp32-stored whitened rows (split-bf16 planes on the tensor cores): what exactness costs
    del case
    torch.cuda.empty_cache()
    index32 = lrb.ExactIndex(d, n, metric="mahalanobis", storage="fp32", device=dev.index,
                             whiten=whitener_from_precision(prec))
    for c in range(-(-n // CHUNK)):
        index32.add(corpus_chunk(c, min(CHUNK, n - c * CHUNK), d, dev, unit=False, aniso=A))
    case32 = SearchCase(ctx, index32, 0, "mahalanobis", 64)
    q_host, qpos, planted = make_queries(64, d, n, dev, unit=False, aniso=A)
    rec, res, _ = case32.measure(q_host, k, 6, 3, n, d)
    rec["planted_neighbours_found"] = bool((res[1].cpu().numpy()[qpos, 0] == planted).all())
    rec["workload"] = (f"mahalanobis top-{k}, 64 queries x {n} x {d}, fp32-stored whitened rows as split-bf16 planes "
                       "(3 MMAs per product; 2x the bytes, 3x the flops of the bf16 rows)")
    rec["roofline"]["bytes_per_launch"] = n * d * 4 + n * 4
    rec["roofline"]["gbs_one_pass"] = rec["roofline"]["bytes_per_launch"] / (rec["kernel_ms"] * 1e-3) / 1e9
    rec["roofline"]["achieved"] = rec["roofline"]["gbs_one_pass"]
    rec["roofline"]["frac"] = rec["roofline"]["gbs_one_pass"] / ctx.pk["hbm_gbs"]
    out["c3_b64_fp32"] = rec
    index32.close()
    if sample is not None:  # Mahalanobis on the CPU = whitening + the reference's euclidean path
        emb = sample.view(d)
        for b in (1, 64):
            r = cpu_search_qps(emb, sample.queries(b, d), k, "euclidean", 2, 1, n)
            out[f"c3_b{b}"]["cpu_baseline"] = {
                "value": r["value"], "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                "sample": f"{b} queries x {r['rows']} whitened rows x {d} fp32, oracle port of the euclidean path "
                          "(bruteforce.py:73-76) -- the reference has no Mahalanobis code"
                          + ("" if r["rows"] >= n else f"; scaled linearly to {n} rows (extrapolated)")}
        out["c3_b4096"]["cpu_baseline"] = dict(out["c3_b64"]["cpu_baseline"],
                                               sample=out["c3_b64"]["cpu_baseline"]["sample"] + "; the 64-query figure "
                                               "(the CPU path is GEMM-bound from there on)")
    return out


def cfg_c4(ctx: Ctx, sample, scale):
    """config 4: Euclidean top-100 over 10M x 768; on N > 1 GPUs row-sharded (ShardedRetriever: per-shard
    append-buffer selection + the fused peer exchange / k-merge with 100 candidates per rank and query)."""
    lrb, dev, world, rank = ctx.lrb, ctx.dev, ctx.world, ctx.rank
    n, d, k = int(10_000_000 * scale), 768, 100
    lo, hi = lrb.shard_bounds(n, world)[rank]
    index = build_shard(lrb, n, d, lo, hi, "euclidean", dev, unit=False)
    case = SearchCase(ctx, index, lo, "euclidean", 4096)
    rows_per_gpu = -(-n // world)
    out = {}
    for b in (1, 64, 4096):
        q_host, qpos, planted = make_queries(b, d, n, dev, unit=False)
        rec, res, res_e2e = case.measure(q_host, k, 10 if b < 4096 else 4, 3, rows_per_gpu, d,
                                            warm_s=0.2 if b < 4096 else 0.0)
        i_dev = res[1].cpu().numpy()
        rec["planted_neighbours_found"] = bool((i_dev[qpos, 0] == planted).all())
        rec["e2e_equals_device_path"] = bool((res_e2e[1].numpy() == i_dev).all())
        rec["workload"] = (f"euclidean top-{k}, {b} queries x {n} x {d} bf16"
                           + (f", row-sharded over {world} GPUs ({rows_per_gpu} rows each)" if world > 1 else ""))
        out[f"c4_b{b}"] = rec
    index.close()
    if sample is not None:
        emb = sample.view(d)  # the 384-d host buffer read as 768-d rows
        for b in (1, 64):
            r = cpu_search_qps(emb, sample.queries(b, d), k, "euclidean", 2, 1, n)
            out[f"c4_b{b}"]["cpu_baseline"] = {
                "value": r["value"], "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                "sample": f"{b} queries x {r['rows']} rows x {d} fp32, oracle port of bruteforce.py:73-83"
                          + ("" if r["rows"] >= n else f"; scaled linearly to {n} rows (extrapolated)")}
        out["c4_b4096"]["cpu_baseline"] = dict(out["c4_b64"]["cpu_baseline"],
                                               sample=out["c4_b64"]["cpu_baseline"]["sample"] + "; the 64-query figure")
    return out


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
    ctx = Ctx(args, lrb, rank, world, local_rank)
    pk = ctx.pk
    torch.set_num_threads(max(1, (os.cpu_count() or 1) // max(1, world)) if world > 1 else (os.cpu_count() or 1))

    # ---- build this rank's shard on the device -------------------------------------
    lo, hi = lrb.shard_bounds(args.rows, world)[rank]
    t_build = time.perf_counter()
    index = build_shard(lrb, args.rows, args.dim, lo, hi, "cosine", dev)
    build_s = time.perf_counter() - t_build
    case = SearchCase(ctx, index, lo, "cosine", max(args.batch, 1))

    q_host, qpos, planted = make_queries(args.batch, args.dim, args.rows, dev)
    q_pin = q_host.pin_memory()
    q_dev = q_host.to(dev)
    k = args.k

    K, W = args.steps, max(3, args.warmup)
    launches0 = lrb._native.launch_count()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms_dev, _, out = ctx.timed(lambda q: case.step_device(q, k), q_dev, K, W)
    launches = lrb._native.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    _, ms_e2e, out_e2e = ctx.timed(lambda q: case.step_e2e(q, k), q_pin, K, 1)
    kms, kms_ranks = ctx.kernel_ms(index, q_dev, k, lo, K)

    # batch-1 (HBM-streaming regime)
    q1_dev, q1_pin = q_dev[:1].contiguous(), q_pin[:1].clone().pin_memory()
    ms1_dev, _, _ = ctx.timed(lambda q: case.step_device(q, k), q1_dev, args.batch1_steps, W)
    _, ms1_e2e, _ = ctx.timed(lambda q: case.step_e2e(q, k), q1_pin, args.batch1_steps, 1)
    kms1, kms1_ranks = ctx.kernel_ms(index, q1_dev, k, lo, args.batch1_steps)

    # the batch sweep of config 5 between its two end points (powers of 4): device-timed step + search kernel
    sweep = {}
    rows_per_gpu_s = -(-args.rows // world)
    for bs in (4, 16, 64, 256, 1024):
        if bs >= args.batch:
            continue
        qs = q_dev[:bs].contiguous()
        ms_s, _, _ = ctx.timed(lambda q: case.step_device(q, k), qs, 5, 3)
        kms_s, _ = ctx.kernel_ms(index, qs, k, lo, 3)
        tf_s = 2.0 * bs * rows_per_gpu_s * args.dim / (kms_s * 1e-3) / 1e12
        gb_s = (rows_per_gpu_s * args.dim * 2 + rows_per_gpu_s * 4) / (kms_s * 1e-3) / 1e9
        sweep[f"b{bs}"] = {"qps": bs / (ms_s * 1e-3), "ms_per_step": ms_s, "kernel_ms": kms_s,
                           "hbm_frac_one_pass": gb_s / pk["hbm_gbs"], "tensor_frac_sustained": tf_s / pk["bf16_tflops_sustained"]}

    # ---- full-size correctness properties (outside the timed regions) -----------------
    case.check()  # no search kernel hit a pipeline timeout, no rank timed out waiting for a peer's candidates
    d_fin, i_fin = out[0].cpu().numpy(), out[1].cpu().numpy()
    planted_ok = bool((i_fin[qpos, 0] == planted).all())
    sorted_ok = bool((np.diff(d_fin, axis=1) <= 0).all())
    same_e2e = bool((out_e2e[1].numpy() == i_fin).all())
    index.close()
    del index, case
    torch.cuda.empty_cache()

    checks = {"planted_neighbours_found": planted_ok, "scores_sorted": sorted_ok, "e2e_equals_device_path": same_e2e}
    parity = oracle_parity_sharded(ctx)
    checks["oracle_parity_sharded"] = bool(all(v for kk, v in parity.items() if not kk.endswith("_why")))
    checks["oracle_parity_cases"] = parity

    # ---- the other BASELINE configurations + the CPU path beside them ------------------
    want = set() if args.configs == "none" else (
        {"c1", "c2", "c3", "c4"} if args.configs == "all" else set(args.configs.split(",")))
    if world > 1:
        want &= {"c4"}  # the one side configuration that shards (BASELINE.json config 4: 2 / 4 / 8 GPUs)
    sample = None
    base = None
    if rank == 0 and world == 1:
        sample = CpuSample(min(args.cpu_rows, args.rows), args.dim)
        base = cpu_baseline_c5(args, sample, 3, 1)
    configs = {}
    cfg_s = time.perf_counter()
    scale = args.config_scale
    if "c1" in want:
        configs.update(cfg_c1(ctx, sample))
    if "c2" in want:
        configs.update(cfg_c2(ctx, sample, scale))
    if "c3" in want:
        configs.update(cfg_c3(ctx, sample, scale))
    if "c4" in want:
        configs.update(cfg_c4(ctx, sample, scale))
    cfg_s = time.perf_counter() - cfg_s

    if rank == 0:
        rows_per_gpu = -(-args.rows // world)
        flops = 2.0 * args.batch * rows_per_gpu * args.dim
        bytes_per_launch = rows_per_gpu * args.dim * 2 + rows_per_gpu * 4  # bf16 rows + fp32 side values
        ach_tf = flops / (kms * 1e-3) / 1e12
        ach_gbs1 = bytes_per_launch / (kms1 * 1e-3) / 1e9
        r1, r4096 = traffic_ratios()
        line = {
            "metric": METRIC_NAME, "value": args.batch / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16", "dtype_detail": "bf16 rows and queries, fp32 accumulate and scores",
            "data": "synthetic",
            "config": workload_config(args, world),
            "e2e": {"value": args.batch / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(q_pin.numel() * 4), "d2h_bytes_per_step": int(args.batch * k * 12)},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "achieved": ach_tf, "peak": pk["bf16_tflops_sustained"],
                         "unit": "TFLOP/s", "frac": ach_tf / pk["bf16_tflops_sustained"],
                         "traffic": None if r4096 is None else r4096 * bytes_per_launch, "traffic_src": TRAFFIC_SRC,
                         "kernel": "umma_search_kernel", "kernel_ms": kms, "kernel_ms_per_rank": kms_ranks,
                         "flops_per_launch": flops,
                         "peak_src": pk["src"] + " sustained bf16 (kernel runs for hundreds of ms per launch)",
                         "frac_of_burst_peak": ach_tf / pk["bf16_tflops"]},
            "batch1": {"value": 1.0 / (ms1_dev * 1e-3), "unit": UNIT, "ms_per_query": ms1_dev,
                       "e2e": {"value": 1.0 / (ms1_e2e * 1e-3), "unit": UNIT, "ms_per_query": ms1_e2e,
                               "h2d_bytes_per_step": args.dim * 4, "d2h_bytes_per_step": k * 12},
                       "roofline": {"bound": "hbm", "achieved": ach_gbs1, "peak": pk["hbm_gbs"], "unit": "GB/s",
                                    "frac": ach_gbs1 / pk["hbm_gbs"],
                                    "traffic": None if r1 is None else r1 * bytes_per_launch, "traffic_src": TRAFFIC_SRC,
                                    "kernel_ms": kms1, "kernel_ms_per_rank": kms1_ranks,
                                    "bytes_per_launch": bytes_per_launch, "peak_src": pk["src"]}},
            "batch_sweep": sweep,
            "clocks": clocks,
            "checks": checks,
            "build_s": build_s,
            "configs": configs,
            "configs_wall_s": cfg_s,
        }
        if base is not None:
            line["cpu_baseline"] = base
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if not (planted_ok and sorted_ok and same_e2e and checks["oracle_parity_sharded"]):
        sys.exit(3)


if __name__ == "__main__":
    main()
