"""Host-side helpers shared by the retriever classes: the StatsTracker contract of the
reference (retrieval/common.py:37-65) and Mahalanobis statistics.  No search arithmetic."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List

import numpy as np
import torch


@dataclass
class StatsTracker:
    """Same four keys, same accumulation rules as the reference's StatsTracker
    (retrieval/common.py:37-65); consumed by main.py:343-344 and
    utils/benchmark_utils.py:36-40 of the reference."""

    build_time_s: float = 0.0
    search_time_s: float = 0.0
    search_calls: int = 0
    per_query_ms: List[float] = field(default_factory=list)

    def add_build_time(self, seconds: float) -> None:
        self.build_time_s += float(seconds)

    def add_search_batch(self, batch_size: int, seconds: float) -> None:
        self.search_time_s += float(seconds)
        self.search_calls += 1
        self.per_query_ms.append((seconds / max(1, int(batch_size))) * 1000.0)

    def get_stats(self, reset: bool = False) -> Dict[str, object]:
        out = {
            "build_time_s": float(self.build_time_s),
            "search_time_s": float(self.search_time_s),
            "search_calls": int(self.search_calls),
            "per_query_ms": list(self.per_query_ms),
        }
        if reset:
            self.build_time_s = 0.0
            self.search_time_s = 0.0
            self.search_calls = 0
            self.per_query_ms.clear()
        return out


def empirical_precision(embeddings: torch.Tensor, chunk: int = 1 << 18) -> np.ndarray:
    """fp64 precision matrix of the MLE covariance (ddof=0, centred) of the corpus: what
    sklearn.covariance.EmpiricalCovariance -- imported and never called at
    retrieval/retriever.py:8 of the reference -- would give.  Index-build statistics on a
    [dim, dim] matrix; runs on whatever device the embeddings live on, in row chunks."""
    n, d = embeddings.shape
    dev = embeddings.device
    s1 = torch.zeros(d, dtype=torch.float64, device=dev)
    s2 = torch.zeros(d, d, dtype=torch.float64, device=dev)
    for lo in range(0, n, chunk):
        x = embeddings[lo : lo + chunk].to(torch.float64)
        s1 += x.sum(0)
        s2 += x.T @ x
    mean = s1 / n
    cov = s2 / n - torch.outer(mean, mean)
    cov = 0.5 * (cov + cov.T)
    return np.linalg.pinv(cov.cpu().numpy(), hermitian=True)


def whitener_from_precision(precision: np.ndarray) -> np.ndarray:
    """Lower Cholesky factor L (precision = L L^T), so that x -> x L turns the Mahalanobis
    form (q-e)^T P (q-e) into a squared Euclidean distance."""
    p = np.asarray(precision, dtype=np.float64)
    return np.linalg.cholesky(0.5 * (p + p.T))
