"""Row-sharded exact search over several GPUs (one process per GPU, torch.distributed).

Net-new relative to the reference, which is single-process: rank r
owns a contiguous block of corpus rows, runs the fused search on its shard and returns
GLOBAL row ids (local id + row offset, added in the merge kernel); one all-gather of the
[B, k] candidates (NCCL over NVLink) and the k-way merge kernel give every rank the global
top-k.  Queries are replicated.  The only exchange is B * k * 12 bytes per rank.

Two exchange paths: "p2p" (default on the native path) -- latent_rag_b200.exchange.PeerExchange,
one kernel per rank that stores the candidates into every peer's buffer over NVLink, waits for
the peers' flags and merges; "nccl" -- all_gather_into_tensor + the merge kernel (also the path
of the CPU tests, over gloo).
"""
from __future__ import annotations

import time
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

from .retrieval.common import StatsTracker, whitener_from_precision


def shard_bounds(n: int, world: int) -> List[Tuple[int, int]]:
    """rank r owns rows [r*ceil(n/W), min(n, (r+1)*ceil(n/W)))."""
    per = -(-n // world) if world > 0 else n
    return [(min(n, r * per), min(n, (r + 1) * per)) for r in range(world)]


def distributed_precision(local_embeddings: torch.Tensor, group=None) -> np.ndarray:
    """Global MLE covariance -> precision from per-shard fp64 moment sums (one all-reduce)."""
    x = local_embeddings
    d = x.size(1)
    buf = torch.zeros(d * d + d + 1, dtype=torch.float64, device=x.device)
    for lo in range(0, x.size(0), 1 << 18):
        c = x[lo : lo + (1 << 18)].to(torch.float64)
        buf[: d * d] += (c.T @ c).reshape(-1)
        buf[d * d : d * d + d] += c.sum(0)
    buf[-1] = float(x.size(0))
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, group=group)
    n = buf[-1]
    mean = buf[d * d : d * d + d] / n
    cov = buf[: d * d].reshape(d, d) / n - torch.outer(mean, mean)
    cov = 0.5 * (cov + cov.T)
    return np.linalg.pinv(cov.cpu().numpy(), hermitian=True)


class ShardedRetriever:
    """One shard of a row-partitioned corpus; `search` returns the GLOBAL top-k on every rank.

    local_embeddings  this rank's rows [n_local, D], or an ExactIndex already holding them (a shard
                      built chunk by chunk on the device; its metric / storage are kept)
    row_offset        global index of this rank's first row
    texts / doc_ids   optional GLOBAL lists for `retrieve` (host side, indexed by global row)
    local_search      TEST-ONLY seam: (queries, k) -> (scores [B,k], global ids [B,k]) torch tensors.
                      The product path is the default -- the native engine on this rank's GPU; the
                      gloo CPU tests (tests/test_dist_gloo.py) inject the oracle here to drive the
                      host-side sharding / padding / exchange logic without a device.
    merge             TEST-ONLY seam: (cand_scores [B,W,k], cand_ids [B,W,k], k) -> (scores, ids);
                      default: the native merge kernel.
    """

    def __init__(
        self,
        local_embeddings: torch.Tensor,
        row_offset: int,
        metric: str = "cosine",
        *,
        texts: Optional[Sequence[str]] = None,
        doc_ids: Optional[Sequence[int]] = None,
        precision: str = "bf16",
        device: Optional[int] = None,
        group=None,
        precision_matrix: Optional[np.ndarray] = None,
        local_search: Optional[Callable] = None,
        merge: Optional[Callable] = None,
        exchange: str = "auto",
        max_batch: int = 4096,
    ):
        if metric not in ("cosine", "euclidean", "mahalanobis"):
            raise ValueError(f"Unsupported metric: {metric}")
        self.metric = metric
        self.group = group
        self.row_offset = int(row_offset)
        prebuilt = None
        if not torch.is_tensor(local_embeddings) and hasattr(local_embeddings, "search") and hasattr(local_embeddings, "size"):
            prebuilt = local_embeddings  # an ExactIndex
            if prebuilt.metric != metric:
                raise ValueError(f"the prebuilt index is {prebuilt.metric}, not {metric}")
        self.n_local = int(prebuilt.size if prebuilt is not None else local_embeddings.size(0))
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.texts = list(texts) if texts is not None else None
        self.doc_ids = list(doc_ids) if doc_ids is not None else None
        self._stats = StatsTracker()
        self._merge = merge
        self.index = None
        t0 = time.perf_counter()
        if local_search is not None:
            self._local_search = local_search
            self.comm_device = torch.device("cpu")
        elif prebuilt is not None:
            self.index = prebuilt
            self.comm_device = torch.device(f"cuda:{prebuilt.device}")
            self._local_search = self._native_local_search
        else:
            from .engine import ExactIndex  # needs the native library + a GPU

            if device is None:
                device = torch.cuda.current_device()
            whiten = None
            if metric == "mahalanobis":
                if precision_matrix is None:
                    precision_matrix = distributed_precision(local_embeddings, group)
                whiten = whitener_from_precision(precision_matrix)
            self.index = ExactIndex(local_embeddings.size(1), max(1, self.n_local), metric=metric, storage=precision,
                                    device=device, whiten=whiten)
            if self.n_local:
                self.index.add(local_embeddings)
            torch.cuda.synchronize(device)
            self.comm_device = torch.device(f"cuda:{device}")
            self._local_search = self._native_local_search
        if exchange not in ("auto", "p2p", "nccl"):
            raise ValueError(f"Unknown exchange: {exchange}")
        self._xchg = None
        if self.index is not None and self.world > 1 and exchange in ("auto", "p2p"):
            from .exchange import PeerExchange

            self._xchg = PeerExchange(self.comm_device.index, self.rank, self.world, max_b=max_batch).connect(group)
        n_total = torch.tensor([self.n_local], dtype=torch.int64, device=self.comm_device)
        if self.world > 1:
            dist.all_reduce(n_total, group=group)
        self.n_total = int(n_total.item())
        self._stats.add_build_time(time.perf_counter() - t0)

    def _native_local_search(self, queries: torch.Tensor, k: int):
        return self.index.search(queries, k, idx_base=self.row_offset, device_out=True)

    def search(self, queries: torch.Tensor, k: int) -> Tuple[np.ndarray, np.ndarray]:
        """Global top-k as numpy (scores float32 [B,k'], ids int64 [B,k']), k' = min(k, N_total)."""
        t0 = time.perf_counter()
        out_d, out_i = self.search_tensors(queries, k)
        if torch.is_tensor(out_d):
            out_d, out_i = out_d.cpu().numpy(), out_i.cpu().numpy()
        if self.index is not None:
            self.index.check()  # device-output searches do not synchronise by themselves
        if self._xchg is not None:
            self._xchg.check()  # a peer that never published shows up here, not as a hang
        b = 1 if queries.dim() == 1 else queries.size(0)
        self._stats.add_search_batch(batch_size=b, seconds=time.perf_counter() - t0)
        return np.asarray(out_d, dtype=np.float32), np.asarray(out_i, dtype=np.int64)

    def search_tensors(self, queries: torch.Tensor, k: int):
        """Same as `search` but leaves the result where the merge produced it (CUDA tensors on
        the native path: no host synchronisation).  Nothing is checked here: a caller that consumes
        these tensors must call `check()` at its own synchronisation point -- a search kernel whose
        pipeline timed out or a peer that never published leaves id -1 / score -inf entries and is
        reported there, not as an exception from this call."""
        if queries.dim() == 1:
            queries = queries.unsqueeze(0)
        b = queries.size(0)
        k = min(int(k), self.n_total)
        # the peer exchange carries up to 128 candidates per rank; deeper results take the all-gather
        xchg = self._xchg if self._xchg is not None and k <= self._xchg.max_k else None
        # 1. local top-k with global ids, padded to k when the shard is smaller than k
        kl = min(k, self.n_local)
        if kl == k and b > 0 and xchg is not None:
            d, i = self._local_search(queries, k)
        else:
            d = torch.full((b, k), float("-inf"), dtype=torch.float32, device=self.comm_device)
            i = torch.full((b, k), -1, dtype=torch.int64, device=self.comm_device)
            if kl > 0 and b > 0:
                dl, il = self._local_search(queries, kl)
                d[:, :kl] = torch.as_tensor(dl, device=self.comm_device)
                i[:, :kl] = torch.as_tensor(il, device=self.comm_device)
        # 2+3 fused: candidates go straight into every peer's buffer, flags, wait, merge
        if xchg is not None:
            return xchg.exchange_merge(d, i, k)
        # 2. the one exchange step: all-gather of the candidates
        if self.world > 1:
            gd = torch.empty((self.world * b, k), dtype=torch.float32, device=self.comm_device)
            gi = torch.empty((self.world * b, k), dtype=torch.int64, device=self.comm_device)
            dist.all_gather_into_tensor(gd, d, group=self.group)  # rank-major concatenation
            dist.all_gather_into_tensor(gi, i, group=self.group)
            cd = gd.view(self.world, b, k).permute(1, 0, 2).contiguous()
            ci = gi.view(self.world, b, k).permute(1, 0, 2).contiguous()
        else:
            cd, ci = d.unsqueeze(1), i.unsqueeze(1)
        # 3. k-way merge
        if self._merge is not None:
            out_d, out_i = self._merge(cd, ci, k)
        else:
            from .engine import merge_topk

            out_d, out_i = merge_topk(cd, ci, k)
        return out_d, out_i

    def check(self) -> None:
        """Synchronise and raise if a search or an exchange since the last check timed out."""
        if self.index is not None:
            self.index.check()
        if self._xchg is not None:
            self._xchg.check()

    def retrieve(self, query_emb: torch.Tensor, top_k: int = 10):
        d, i = self.search(query_emb, top_k)
        idxs = i[0].tolist()
        texts = [self.texts[j] for j in idxs] if self.texts is not None else [""] * len(idxs)
        docids = [self.doc_ids[j] for j in idxs] if self.doc_ids is not None else idxs
        return texts, d[0].tolist(), docids

    def get_stats(self, reset: bool = False):
        return self._stats.get_stats(reset=reset)
