"""evaluate_retrieval on the device: drop-in for evaluation/retrieval_metrics.py:55-96 of the
reference (Recall@k, MRR@k, nDCG@k; per-metric mean and sample standard deviation), with the
per-query values computed by one kernel launch for the whole query batch instead of a Python
loop per query and metric.  The values are the reference's bit for bit (float64, same order
of summation); the mean / std over queries are taken with numpy exactly as the reference does.
"""
from __future__ import annotations

from ctypes import c_void_p
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _native as nat

ID = Union[int, str]
_KIND = {"recall": 0, "mrr": 1, "ndcg": 2}


def _parse_metric(m: str) -> Tuple[str, Optional[int]]:
    return (m.split("@")[0], int(m.split("@")[1])) if "@" in m else (m, None)


def _encode(retrieved_batch, relevant_batch):
    """ids (int or str) -> dense int64 codes shared by both sides; retrieved padded with -1."""
    codes: Dict[ID, int] = {}

    def code(x):
        c = codes.get(x)
        if c is None:
            c = codes[x] = len(codes)
        return c

    q = len(retrieved_batch)
    kr = max(1, max((len(r) for r in retrieved_batch), default=1))
    ret = np.full((q, kr), -1, dtype=np.int64)
    for i, r in enumerate(retrieved_batch):
        ret[i, : len(r)] = [code(x) for x in r]
    off = np.zeros(q + 1, dtype=np.int64)
    rel: List[int] = []
    for i, r in enumerate(relevant_batch):
        rel.extend(code(x) for x in r)
        off[i + 1] = len(rel)
    return ret, off, np.asarray(rel if rel else [0], dtype=np.int64)


def evaluate_retrieval(retrieved_batch, relevant_batch, metrics: Optional[List[str]] = None, *,
                       return_per_query: bool = False, device: Optional[int] = None):
    """Same contract as the reference's evaluate_retrieval (retrieval_metrics.py:55-96)."""
    single = isinstance(retrieved_batch[0], (str, int, np.integer))
    if single:
        retrieved_batch, relevant_batch = [retrieved_batch], [relevant_batch]
    assert len(retrieved_batch) == len(relevant_batch), \
        "retrieved_batch and relevant_batch must have the same length."
    if not metrics:
        raise ValueError("No metrics specified.")
    kinds, ks = [], []
    for m in metrics:
        name, k = _parse_metric(m)
        name = name.lower()
        if name not in _KIND or (name != "mrr" and k is None):
            raise ValueError(f"Metric '{name}' not found.")  # retrieval_metrics.py:53
        kinds.append(_KIND[name])
        ks.append(-1 if k is None else int(k))
    lib = nat.load()
    nat.require_device()
    if device is None:
        device = torch.cuda.current_device()
    dev = torch.device(f"cuda:{device}")
    ret, off, rel = _encode(retrieved_batch, relevant_batch)
    q, kr = ret.shape
    # the reference's own discounts, float64; the ideal DCG of "ndcg@k" sums min(len(relevant), k) of
    # them with the caller's k, which can exceed the retrieved length (retrieval_metrics.py:29)
    disc = 1.0 / np.log2(np.arange(max([kr] + ks), dtype=np.int64) + 2)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d_ret, d_off, d_rel, d_disc = t(ret), t(off), t(rel), t(disc)
    d_kind, d_k = t(np.asarray(kinds, dtype=np.int32)), t(np.asarray(ks, dtype=np.int32))
    out = torch.empty((q, len(metrics)), dtype=torch.float64, device=dev)
    nat.check(lib.lk_retrieval_metrics(device, c_void_p(d_ret.data_ptr()), q, kr, c_void_p(d_off.data_ptr()),
                                       c_void_p(d_rel.data_ptr()), c_void_p(d_kind.data_ptr()),
                                       c_void_p(d_k.data_ptr()), len(metrics), c_void_p(d_disc.data_ptr()),
                                       c_void_p(out.data_ptr()),
                                       c_void_p(int(torch.cuda.current_stream(device).cuda_stream))),
              "lk_retrieval_metrics")
    vals = out.cpu().numpy()
    summary: Dict[str, Dict[str, float]] = {}
    per_query: List[Dict[str, float]] = [{} for _ in range(q)]
    for j, m in enumerate(metrics):
        col = [float(v) for v in vals[:, j]]
        summary[m] = {"mean": float(np.mean(col)), "std": float(np.std(col, ddof=1)) if q > 1 else 0.0}
        for d, v in zip(per_query, col):
            d[m] = v
    if return_per_query:
        return summary, per_query
    if single:
        return {k: v["mean"] for k, v in summary.items()}
    return summary


def rank_positive(q: torch.Tensor, d: torch.Tensor, device: Optional[int] = None) -> torch.Tensor:
    """1-based rank of each query's paired document by cosine similarity: drop-in for
    `_rank_positive` of the reference (evaluation/embedding_visualization.py:34-37), without
    the [n, n, D] broadcast.  Returns an int64 tensor on the device of `q`."""
    if q.dim() != 2 or q.shape != d.shape:
        raise ValueError(f"expected two [n, D] tensors of the same shape, got {tuple(q.shape)} and {tuple(d.shape)}")
    lib = nat.load()
    nat.require_device()
    if device is None:
        device = q.device.index if q.is_cuda else torch.cuda.current_device()
    dev = torch.device(f"cuda:{device}")
    qd = q.detach().to(dev, torch.float32).contiguous()
    dd = d.detach().to(dev, torch.float32).contiguous()
    out = torch.empty(q.size(0), dtype=torch.int64, device=dev)
    nat.check(lib.lk_rank_positive(device, c_void_p(qd.data_ptr()), c_void_p(dd.data_ptr()), q.size(0), q.size(1),
                                   c_void_p(out.data_ptr()),
                                   c_void_p(int(torch.cuda.current_stream(device).cuda_stream))), "lk_rank_positive")
    return out if q.is_cuda else out.cpu()
