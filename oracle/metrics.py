"""CPU restatement of the reference's retrieval metrics.  TEST INFRASTRUCTURE.

Follows evaluation/retrieval_metrics.py:14-96 of the reference; pinned by the KATs
of test/test_evaluation.py:9-22 (recall 1/3, mrr 1/3, ndcg@3 0.23463936301137822).
"""
from __future__ import annotations

import math

import numpy as np
from typing import Dict, List, Optional, Sequence, Tuple, Union

ID = Union[int, str]


def recall_at_k(retrieved: Sequence[ID], relevant: Sequence[ID], k: int) -> float:
    """retrieval_metrics.py:14-17."""
    if not relevant:
        return 0.0
    return len(set(retrieved[:k]) & set(relevant)) / len(relevant)


def mrr(retrieved: Sequence[ID], relevant: Sequence[ID]) -> float:
    """retrieval_metrics.py:19-23: reciprocal rank of the first relevant hit."""
    for rank, d in enumerate(retrieved, 1):
        if d in relevant:
            return 1.0 / rank
    return 0.0


def ndcg_at_k(retrieved: Sequence[ID], relevant: Sequence[ID], k: int) -> float:
    """retrieval_metrics.py:25-31: binary gains, log2(i+2) discount.  numpy's log2 like the
    reference (math.log2 differs from it in the last bit for some arguments), summed left to right
    as Python floats, misses contributing 0.0."""
    dcg = sum(1.0 / np.log2(i + 2) if d in relevant else 0.0 for i, d in enumerate(retrieved[:k]))
    idcg = sum(1.0 / np.log2(i + 2) for i in range(min(len(relevant), k)))
    return dcg / idcg if idcg else 0.0


def _parse(metric: str) -> Tuple[str, Optional[int]]:
    if "@" in metric:
        name, kk = metric.split("@")
        return name, int(kk)
    return metric, None


def _one(retrieved, relevant, name: str, k: Optional[int]) -> float:
    name = name.lower()
    if name == "recall" and k is not None:
        return recall_at_k(retrieved, relevant, k)
    if name == "mrr":
        return mrr(retrieved[: (k or len(retrieved))], relevant)
    if name == "ndcg" and k is not None:
        return ndcg_at_k(retrieved, relevant, k)
    raise ValueError(f"Metric '{name}' not found.")


def evaluate_retrieval(
    retrieved_batch: List[Sequence[ID]], relevant_batch: List[Sequence[ID]], metrics: List[str]
) -> Dict[str, Dict[str, float]]:
    """retrieval_metrics.py:55-96, batch form: per metric np.mean and the sample standard
    deviation np.std(ddof=1) (0.0 for a single query)."""
    assert len(retrieved_batch) == len(relevant_batch)
    if not metrics:
        raise ValueError("No metrics specified.")
    q = len(retrieved_batch)
    out: Dict[str, Dict[str, float]] = {}
    for m in metrics:
        name, k = _parse(m)
        vals = [_one(r, rel, name, k) for r, rel in zip(retrieved_batch, relevant_batch)]
        # numpy's (pairwise) mean and std like retrieval_metrics.py:85-88, so the summary is the reference's bit for bit
        out[m] = {"mean": float(np.mean(vals)), "std": float(np.std(vals, ddof=1)) if q > 1 else 0.0}
    return out


def maxsim_rerank(scores_k, docids_k, top_k: int):
    """Document-level MaxSim aggregation of one query's candidates, as the reference's caller
    does it (main.py:273-282): best score per doc id, documents sorted by it (descending,
    Python's stable sort: ties stay in first-seen order), truncated to top_k.
    -> (ranked doc ids, their scores)."""
    agg = {}
    for did, sc in zip(docids_k, scores_k):
        prev = agg.get(did)
        if (prev is None) or (sc > prev):
            agg[did] = sc
    ranked = sorted(agg, key=agg.get, reverse=True)[:top_k]
    return ranked, [agg[d] for d in ranked]


def rank_positive(q, d):
    """1-based rank of each paired document by cosine similarity
    (evaluation/embedding_visualization.py:34-37): rank_i = 1 + #{j : sim[i, j] > sim[i, i]}
    (equal to the reference's double argsort whenever no other document ties with the pair)."""
    import torch
    import torch.nn.functional as F

    qn = F.normalize(q.detach().to("cpu", torch.float32), dim=-1, eps=1e-8)
    dn = F.normalize(d.detach().to("cpu", torch.float32), dim=-1, eps=1e-8)
    sim = qn @ dn.T
    return (sim > sim.diagonal().unsqueeze(1)).sum(dim=1) + 1
