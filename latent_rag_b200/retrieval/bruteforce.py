"""BruteForceRetriever on the B200 engine: drop-in for the reference class of the same
name (retrieval/bruteforce.py:17-95) -- same constructor arguments, same `search`,
`retrieve` and `get_stats` results -- with the corpus resident in HBM and the fused
distance + top-k kernels of liblatentknn doing the work.  There is no CPU path."""
from __future__ import annotations

import time
from typing import List, Literal, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from ..engine import ExactIndex
from .common import BatchedRetrieveMixin, StatsTracker, empirical_precision, whitener_from_precision

Similarity = Literal["cosine", "euclidean", "mahalanobis"]


class BruteForceRetriever(BatchedRetrieveMixin):
    """Exact retriever with performance metrics (reference: retrieval/bruteforce.py:17-24).

    Positional arguments are the reference's: `embeddings [N, D]`, `texts`, `doc_ids`,
    `metric` ("cosine" | "euclidean", plus "mahalanobis", which the reference's README
    names and its code never implemented).  Keyword-only additions:
      precision        "bf16" (default; tensor-core path, results are those of the reference
                       run on bf16-rounded inputs) or "fp32" (exact fp32 path)
      device           CUDA device index (default: torch's current device)
      precision_matrix [D, D] inverse covariance for mahalanobis (default: estimated from
                       `embeddings` like sklearn's EmpiricalCovariance)
      keep_source      keep a REFERENCE (not a copy) to `embeddings` so that the reference's public
                       `.emb` attribute can be rebuilt on demand (default True: drop-in).  False lets
                       the caller's corpus tensor be freed -- the engine's own copy lives in HBM --
                       and makes `.emb` raise.
    """

    def __init__(
        self,
        embeddings: torch.Tensor,
        texts: Sequence[str],
        doc_ids: Sequence[int] | None = None,
        metric: Similarity = "cosine",
        *,
        precision: str = "bf16",
        device: Optional[int] = None,
        precision_matrix: Optional[np.ndarray] = None,
        keep_source: bool = True,
    ):
        if doc_ids is not None:
            assert len(texts) == len(doc_ids), "len mismatch (texts vs doc_ids)"  # bruteforce.py:33-34
        if metric not in ("cosine", "euclidean", "mahalanobis"):
            raise ValueError(f"Unsupported metric: {metric}")  # bruteforce.py:54

        self.texts = list(texts)
        self.doc_ids = list(doc_ids) if doc_ids is not None else list(range(len(texts)))
        self.metric = metric
        self.precision = precision
        self._stats = StatsTracker()
        self._src = embeddings if keep_source else None  # a reference, not a copy: only `.emb` ever reads it again

        if isinstance(embeddings, np.ndarray):
            embeddings = torch.from_numpy(embeddings)
        if embeddings.dim() != 2:
            raise ValueError(f"embeddings must be [N, D], got {tuple(embeddings.shape)}")
        if device is None:
            device = embeddings.device.index if embeddings.is_cuda else torch.cuda.current_device()

        t0 = time.perf_counter()
        whiten = None
        if metric == "mahalanobis":
            if precision_matrix is None:
                precision_matrix = empirical_precision(embeddings)
            self.precision_matrix = np.asarray(precision_matrix, dtype=np.float64)
            whiten = whitener_from_precision(self.precision_matrix)
        self.index = ExactIndex(embeddings.size(1), max(1, embeddings.size(0)), metric=metric, storage=precision,
                                device=device, whiten=whiten)
        if embeddings.size(0):
            self.index.add(embeddings)
        torch.cuda.synchronize(device)
        self._stats.add_build_time(time.perf_counter() - t0)

    # the reference exposes the (normalised) CPU corpus as `.emb` (bruteforce.py:49-53);
    # the engine's copy lives in HBM as bf16 tiles, so this one is rebuilt on demand
    @property
    def emb(self) -> torch.Tensor:
        if self._src is None:
            raise AttributeError("this retriever was built with keep_source=False: the corpus only exists as tiles in HBM")
        e = self._src if torch.is_tensor(self._src) else torch.from_numpy(np.asarray(self._src))
        e = e.detach().to("cpu", torch.float32)
        return F.normalize(e, p=2, dim=1).contiguous() if self.metric == "cosine" else e.contiguous()

    def __len__(self) -> int:
        return self.index.size

    # batch API (bruteforce.py:58-83): (scores float32 [B,k'], indices int64 [B,k']), k' = min(k, N)
    def search(self, queries: torch.Tensor, k: int) -> Tuple[np.ndarray, np.ndarray]:
        if isinstance(queries, np.ndarray):
            queries = torch.from_numpy(queries)
        if queries.dim() == 1:
            queries = queries.unsqueeze(0)
        n = self.index.size
        k = min(int(k), n)  # bruteforce.py:81
        if k < 1 or queries.size(0) == 0:
            return (np.empty((queries.size(0), max(k, 0)), dtype=np.float32),
                    np.empty((queries.size(0), max(k, 0)), dtype=np.int64))
        t0 = time.perf_counter()
        d, i = self.index.search(queries, k)
        self._stats.add_search_batch(batch_size=len(queries), seconds=time.perf_counter() - t0)
        return d, i

    def _search_device(self, queries: torch.Tensor, k: int):
        """`search` that leaves (scores, row ids) on the device: for retrieve_batch."""
        if isinstance(queries, np.ndarray):
            queries = torch.from_numpy(queries)
        if queries.dim() == 1:
            queries = queries.unsqueeze(0)
        k = min(int(k), self.index.size)
        dev = torch.device(f"cuda:{self.index.device}")
        if k < 1 or queries.size(0) == 0:
            return (torch.empty((queries.size(0), max(k, 0)), dtype=torch.float32, device=dev),
                    torch.empty((queries.size(0), max(k, 0)), dtype=torch.int64, device=dev))
        t0 = time.perf_counter()
        d, i = self.index.search(queries, k, device_out=True)
        self._stats.add_search_batch(batch_size=len(queries), seconds=time.perf_counter() - t0)
        return d, i

    # single-query convenience (bruteforce.py:86-92)
    def retrieve(self, query_emb: torch.Tensor, top_k: int = 10) -> Tuple[List[str], List[float], List[int]]:
        d, i = self.search(query_emb, top_k)
        idxs = i[0].tolist()
        texts = [self.texts[j] for j in idxs]
        scores = d[0].tolist()
        docids = [self.doc_ids[j] for j in idxs]
        return texts, scores, docids

    def get_stats(self, reset: bool = False):
        return self._stats.get_stats(reset=reset)
