"""Host-side helpers shared by the retriever classes: the StatsTracker contract of the
reference (retrieval/common.py:37-65) and Mahalanobis statistics.  No search arithmetic."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

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


class BatchedRetrieveMixin:
    """`retrieve_batch`: the reference caller's hot loop (main.py:264-282) as two device calls --
    one search at `candidate_k` for the whole query batch and the document-level MaxSim
    aggregation kernel -- instead of one `retrieve` + dict + sort per query in Python.

    Needs `self.index` (ExactIndex), `self.doc_ids` / `self._doc_ids` (row -> document id) and the
    class's own query preparation through `_search_device(queries, k)`.
    """

    _row_doc_dev = None
    _row_doc_len = -1

    def _row_doc_ids(self):
        return self.doc_ids if hasattr(self, "doc_ids") else self._doc_ids

    def _row_doc_tensor(self, device):
        ids = self._row_doc_ids()
        if self._row_doc_dev is None or self._row_doc_len != len(ids) or self._row_doc_dev.device != device:
            self._row_doc_dev = torch.as_tensor(np.asarray(ids, dtype=np.int64)).to(device)
            self._row_doc_len = len(ids)
        return self._row_doc_dev

    def retrieve_batch(self, query_embeddings, top_k: int = 10, candidate_k: Optional[int] = None):
        """-> (doc_ids List[List[int]], scores List[List[float]]): per query the `top_k` documents
        ranked by their best chunk among the `candidate_k` nearest chunks (main.py:264-282:
        candidate_k defaults to top_k; the reference uses 3 * top_k when chunking)."""
        from ctypes import c_void_p

        from .. import _native as nat

        candidate_k = int(top_k if candidate_k is None else candidate_k)
        d, i = self._search_device(query_embeddings, candidate_k)  # CUDA tensors [B, ck]
        b, ck = d.shape
        top_k = min(int(top_k), ck)
        if b == 0 or ck == 0:
            return [[] for _ in range(b)], [[] for _ in range(b)]
        dev = d.device
        row_doc = self._row_doc_tensor(dev)
        out_s = torch.empty((b, top_k), dtype=torch.float32, device=dev)
        out_d = torch.empty((b, top_k), dtype=torch.int64, device=dev)
        lib = nat.load()
        nat.check(lib.lk_maxsim_rerank(dev.index, c_void_p(d.data_ptr()), c_void_p(i.data_ptr()), b, ck,
                                       c_void_p(row_doc.data_ptr()), row_doc.numel(), top_k,
                                       c_void_p(out_s.data_ptr()), c_void_p(out_d.data_ptr()),
                                       c_void_p(int(torch.cuda.current_stream(dev).cuda_stream))),
                  "lk_maxsim_rerank")
        ids, sc = out_d.cpu().numpy(), out_s.cpu().numpy()
        self.index.check()
        keep = np.isfinite(sc)  # padding: fewer documents than top_k among the candidates
        return ([ids[r][keep[r]].tolist() for r in range(b)], [sc[r][keep[r]].tolist() for r in range(b)])
