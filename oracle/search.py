"""CPU restatement of the reference's exact-search arithmetic.  TEST INFRASTRUCTURE.

Follows `retrieval/bruteforce.py:26-83` and `retrieval/common.py:18-32` of the
reference (paths relative to /root/reference).  torch CPU ops are used on purpose:
the reference *is* `F.normalize` + `mm` + `torch.topk` on CPU, so the same ATen
calls reproduce it bit for bit.  The only structural change is that queries are
processed in row chunks so the `[B, N]` score matrix stays bounded; a row's neighbours
do not depend on the chunking (its scores can move in the last bit: ATen picks different
GEMM kernels for different batch heights, exactly as the reference's own `q @ emb.T` does).
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

_CHUNK_BYTES = 1 << 30  # bound on the fp32 [b, N] score block


def bf16_round(x: torch.Tensor) -> torch.Tensor:
    """fp32 -> bf16 (round to nearest even) -> fp32: the values the bf16 kernels see."""
    return x.to(torch.float32).to(torch.bfloat16).to(torch.float32)


# ----------------------------------------------------------------------------
# BruteForceRetriever
# ----------------------------------------------------------------------------
def bruteforce_build(embeddings: torch.Tensor, metric: str = "cosine") -> torch.Tensor:
    """Index build.  bruteforce.py:36-55: move to CPU, cosine => F.normalize(dim=1)
    (common.py:30-32, eps 1e-12), euclidean => as is, anything else => ValueError."""
    emb = embeddings.detach().to("cpu", torch.float32)
    if metric == "cosine":
        return F.normalize(emb, p=2, dim=1).contiguous()
    if metric == "euclidean":
        return emb.contiguous()
    raise ValueError(f"Unsupported metric: {metric}")


def bruteforce_search(
    emb: torch.Tensor, queries: torch.Tensor, k: int, metric: str = "cosine", chunk_bytes: int = _CHUNK_BYTES
) -> Tuple[np.ndarray, np.ndarray]:
    """bruteforce.py:58-83.  `emb` is the output of `bruteforce_build`.

    cosine   : scores = normalize(Q) @ emb.T                       (:66-69)
    euclidean: scores = -(|q|^2 + |e|^2 - 2 q.e)                   (:73-76)
    k = min(k, N); torch.topk sorted descending; numpy (D, I)      (:81-83)
    chunk_bytes bounds the fp32 [b, N] score block (BASELINE.md section 3 times the CPU path with
    8 GB blocks); results do not depend on it.
    """
    if queries.dim() == 1:
        queries = queries.unsqueeze(0)
    queries = queries.detach().to("cpu", torch.float32)
    n = emb.size(0)
    k = min(int(k), n)
    b = queries.size(0)
    out_d = np.empty((b, k), dtype=np.float32)
    out_i = np.empty((b, k), dtype=np.int64)
    if b == 0:
        return out_d, out_i
    step = max(1, min(b, int(chunk_bytes) // max(1, 4 * n)))
    if metric == "euclidean":
        e2 = (emb * emb).sum(dim=1).unsqueeze(0)
    for s in range(0, b, step):
        q = queries[s : s + step]
        if metric == "cosine":
            qn = F.normalize(q, p=2, dim=1)
            scores = qn @ emb.T
        elif metric == "euclidean":
            q2 = (q * q).sum(dim=1, keepdim=True)
            scores = -(q2 + e2 - 2.0 * (q @ emb.T))
        else:
            raise ValueError(f"Unsupported metric: {metric}")
        vals, idxs = torch.topk(scores, k=k, dim=1)
        out_d[s : s + step] = vals.numpy()
        out_i[s : s + step] = idxs.numpy()
    return out_d, out_i


# ----------------------------------------------------------------------------
# FAISSEmbeddingRetriever, index_type="flatip"
# ----------------------------------------------------------------------------
def _faiss_normalize_l2(x: np.ndarray) -> np.ndarray:
    """`faiss.normalize_L2` (called at common.py:25): each row is divided by its L2
    norm when the norm is > 0 and left untouched otherwise [upstream behaviour]."""
    x = np.ascontiguousarray(x, dtype=np.float32).copy()
    nrm = np.sqrt((x.astype(np.float32) ** 2).sum(axis=1, dtype=np.float32))
    nz = nrm > 0
    x[nz] /= nrm[nz, None]
    return x


def faiss_flatip_build(embeddings: torch.Tensor) -> np.ndarray:
    """FAISSEmbeddingRetriever.py:206-209,252-257: fp32 C-contiguous numpy copy,
    L2-normalise in place, `index.add`."""
    return _faiss_normalize_l2(embeddings.detach().cpu().numpy().astype("float32", copy=False))


def faiss_flatip_search(xb: np.ndarray, queries: torch.Tensor, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """FAISSEmbeddingRetriever.py:314-326: normalise queries, IndexFlatIP.search.
    [upstream] IndexFlatIP pads with id -1 / score -FLT_MAX when k > ntotal instead
    of clamping k the way bruteforce.py:81 does."""
    if queries.dim() == 1:
        queries = queries.unsqueeze(0)
    q = _faiss_normalize_l2(queries.detach().cpu().numpy().astype("float32", copy=False))
    n = xb.shape[0]
    kk = min(int(k), n)
    scores = torch.from_numpy(q) @ torch.from_numpy(xb).T
    vals, idxs = torch.topk(scores, k=kk, dim=1)
    d = np.full((q.shape[0], int(k)), -np.finfo(np.float32).max, dtype=np.float32)
    i = np.full((q.shape[0], int(k)), -1, dtype=np.int64)
    d[:, :kk] = vals.numpy()
    i[:, :kk] = idxs.numpy()
    return d, i


# ----------------------------------------------------------------------------
# Mahalanobis (PARITY UNPINNED: our definition)
# ----------------------------------------------------------------------------
def mahalanobis_precision(corpus: torch.Tensor) -> np.ndarray:
    """fp64 precision matrix of sklearn's EmpiricalCovariance (MLE, ddof=0, centred):
    the estimator the reference imports at retrieval/retriever.py:8 and never calls."""
    x = corpus.detach().cpu().numpy().astype(np.float64)
    xc = x - x.mean(axis=0, keepdims=True)
    cov = (xc.T @ xc) / x.shape[0]
    return np.linalg.pinv(cov, hermitian=True)


def mahalanobis_whitener(precision: np.ndarray) -> np.ndarray:
    """Lower Cholesky factor L of the precision (P = L L^T): x' = x L turns
    (q-e)^T P (q-e) into |q' - e'|^2."""
    return np.linalg.cholesky(np.asarray(precision, dtype=np.float64))


def mahalanobis_search(
    corpus: torch.Tensor, queries: torch.Tensor, k: int, precision: Optional[np.ndarray] = None
) -> Tuple[np.ndarray, np.ndarray]:
    """score(q, e) = -(q-e)^T P (q-e) in fp64, top-k descending, (D float32, I int64)."""
    if queries.dim() == 1:
        queries = queries.unsqueeze(0)
    if precision is None:
        precision = mahalanobis_precision(corpus)
    e = corpus.detach().cpu().numpy().astype(np.float64)
    q = queries.detach().cpu().numpy().astype(np.float64)
    ep = e @ precision
    e_pe = np.einsum("nd,nd->n", ep, e)
    q_pq = np.einsum("bd,bd->b", q @ precision, q)
    scores = -(q_pq[:, None] + e_pe[None, :] - 2.0 * (q @ ep.T))
    kk = min(int(k), e.shape[0])
    vals, idxs = torch.topk(torch.from_numpy(scores), k=kk, dim=1)
    return vals.numpy().astype(np.float32), idxs.numpy().astype(np.int64)


# ----------------------------------------------------------------------------
# Row-sharded search: per-shard top-k + k-way merge (net-new)
# ----------------------------------------------------------------------------
def shard_bounds(n: int, world: int) -> Sequence[Tuple[int, int]]:
    """Contiguous row blocks: rank r owns [r*ceil(n/W), min(n, (r+1)*ceil(n/W)))."""
    per = -(-n // world) if world > 0 else n
    return [(min(n, r * per), min(n, (r + 1) * per)) for r in range(world)]


def merge_topk(cand_d: np.ndarray, cand_i: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """Merge `[B, L]` candidate (score, global index) lists into the top-k, best first,
    ties by lower index; candidates with index < 0 are padding."""
    cand_d = np.asarray(cand_d).reshape(len(cand_d), -1)  # [B, L] or [B, lists, len]
    cand_i = np.asarray(cand_i).reshape(len(cand_i), -1)
    d = np.where(cand_i < 0, -np.inf, cand_d.astype(np.float32))
    order = np.lexsort((cand_i, -d), axis=1)[:, :k]
    return np.take_along_axis(d, order, 1).astype(np.float32), np.take_along_axis(cand_i, order, 1)


# ----------------------------------------------------------------------------
# Tie-aware comparison (SURVEY.md section 8c "parity protocol")
# ----------------------------------------------------------------------------
def euclidean_scale(emb: torch.Tensor, queries: torch.Tensor) -> np.ndarray:
    """Per-query magnitude of the terms the euclidean score cancels: |q|^2 + max|e|^2.
    `-(q2 + e2 - 2qe)` is a difference of numbers of this size, so fp32 results (the
    reference's included: it returns 1.5e-5 for an exact self-match at d=64) carry an
    absolute error of a few ulp of THIS magnitude, not of the score."""
    if queries.dim() == 1:
        queries = queries.unsqueeze(0)
    e2 = (emb.double() ** 2).sum(1).max().item() if emb.numel() else 0.0
    return ((queries.double() ** 2).sum(1) + e2).numpy()


def topk_equivalent(d_ref, i_ref, d_got, i_got, rtol: float = 1e-5, scale=None) -> Tuple[bool, str]:
    """True when `got` is the same top-k as `ref` up to score ties: rows must match
    index for index, except inside groups whose reference scores agree within
    `rtol * max(1, |s|, scale)`, where only the index *sets* (and the scores) must
    agree.  `scale` (per query) is the magnitude of the operands when the score is a
    cancelling difference (see `euclidean_scale`); None for cosine.
    torch.topk's tie order is unspecified, so tie order is never asserted."""
    d_ref = np.asarray(d_ref); i_ref = np.asarray(i_ref)
    d_got = np.asarray(d_got); i_got = np.asarray(i_got)
    if d_ref.shape != d_got.shape or i_ref.shape != i_got.shape:
        return False, f"shape mismatch {d_ref.shape}/{i_ref.shape} vs {d_got.shape}/{i_got.shape}"
    floor = np.ones((d_ref.shape[0], 1)) if scale is None else np.maximum(1.0, np.asarray(scale, dtype=np.float64).reshape(-1, 1))
    tol = rtol * np.maximum(floor, np.abs(d_ref))
    bad = np.abs(d_ref - d_got) > tol
    if bad.any():
        r, c = np.argwhere(bad)[0]
        return False, f"score mismatch at ({r},{c}): ref {d_ref[r, c]!r} got {d_got[r, c]!r}"
    rows = np.nonzero((i_ref != i_got).any(axis=1))[0]
    for r in rows:
        kk = d_ref.shape[1]
        c = 0
        while c < kk:
            e = c + 1
            while e < kk and abs(d_ref[r, e] - d_ref[r, e - 1]) <= tol[r, e]:
                e += 1
            ref_set, got_set = set(i_ref[r, c:e].tolist()), set(i_got[r, c:e].tolist())
            if ref_set != got_set:
                # a tie group cut by the k boundary may legitimately differ in its last
                # members: accept when every differing index scores within tol of the cut
                if e == kk:
                    ok = all(
                        abs(d_got[r, c + j] - d_ref[r, kk - 1]) <= tol[r, kk - 1]
                        for j, ix in enumerate(i_got[r, c:e].tolist())
                        if ix not in ref_set
                    )
                    if ok:
                        c = e
                        continue
                return False, f"row {r} cols [{c},{e}): ref {sorted(ref_set)} got {sorted(got_set)}"
            c = e
    return True, ""
