"""ExactIndex: the Python handle on a native lk_index (include/latentknn.h).

Thin by design: it converts tensors to (pointer, dtype, memory space) triples, picks the
CUDA stream torch is using and calls the C ABI.  No arithmetic happens here.
"""
from __future__ import annotations

import ctypes
from ctypes import byref, c_float, c_int, c_int64, c_void_p
from typing import Optional, Tuple, Union

import numpy as np
import torch

from . import _native as nat

ArrayLike = Union[torch.Tensor, np.ndarray]


def _as_tensor(x: ArrayLike) -> torch.Tensor:
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    if not torch.is_tensor(x):
        raise TypeError(f"expected a torch.Tensor or numpy array, got {type(x).__name__}")
    x = x.detach()
    if x.dtype not in (torch.float32, torch.bfloat16):
        x = x.to(torch.float32)
    return x.contiguous()


def _triple(x: torch.Tensor, device: int) -> Tuple[int, int, int, torch.Tensor]:
    """(pointer, lk_dtype, lk_mem, keep-alive tensor) for a contiguous f32/bf16 tensor."""
    if x.is_cuda and x.device.index != device:
        x = x.to(f"cuda:{device}")
    dtype = nat.LK_F32 if x.dtype == torch.float32 else nat.LK_BF16
    mem = nat.LK_DEVICE if x.is_cuda else nat.LK_HOST
    return x.data_ptr(), dtype, mem, x


def _stream(device: int) -> int:
    return int(torch.cuda.current_stream(device).cuda_stream)


class ExactIndex:
    """A row-block-tiled corpus resident in one GPU's HBM plus the fused search.

    metric   "cosine" | "euclidean" | "mahalanobis"
    storage  "bf16" (tcgen05 path; scores are those of the reference fed bf16-rounded
             inputs) | "fp32" (fp32-level scores: split-bf16 planes on the tensor cores for
             b * rows >= 2^20, the fp32 FMA kernel below that)
    whiten   for mahalanobis: [dim, dim] fp64 L with precision = L @ L.T
    """

    def __init__(self, dim: int, capacity: int, metric: str = "cosine", storage: str = "bf16",
                 device: int = 0, whiten: Optional[np.ndarray] = None):
        if metric not in nat.METRICS:
            raise ValueError(f"Unsupported metric: {metric}")
        if storage not in nat.STORAGE:
            raise ValueError(f"Unsupported storage precision: {storage}")
        self._lib = nat.load()
        nat.require_device()
        self.dim, self.metric, self.storage, self.device = int(dim), metric, storage, int(device)
        self._h = c_void_p()
        wptr = None
        if whiten is not None:
            w = np.ascontiguousarray(np.asarray(whiten, dtype=np.float64))
            if w.shape != (dim, dim):
                raise ValueError(f"whitening matrix must be [{dim}, {dim}], got {w.shape}")
            self._whiten = w
            wptr = w.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        nat.check(
            self._lib.lk_index_create(byref(self._h), self.device, int(max(1, capacity)), self.dim,
                                      nat.METRICS[metric], nat.STORAGE[storage], wptr),
            "lk_index_create",
        )
        self.capacity = int(max(1, capacity))
        self._pinned: Optional[Tuple[torch.Tensor, torch.Tensor]] = None

    # -- lifetime ---------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.lk_index_destroy(self._h)
            self._h = c_void_p()

    def __del__(self):  # pragma: no cover - best effort
        try:
            self.close()
        except Exception:
            pass

    # -- build ------------------------------------------------------------------------
    @property
    def size(self) -> int:
        n = c_int64(0)
        nat.check(self._lib.lk_index_size(self._h, byref(n), None), "lk_index_size")
        return int(n.value)

    ntotal = size  # faiss spelling

    def reserve(self, capacity: int) -> None:
        if capacity > self.capacity:
            nat.check(self._lib.lk_index_reserve(self._h, int(capacity), c_void_p(_stream(self.device))),
                      "lk_index_reserve")
            self.capacity = int(capacity)

    def add(self, rows: ArrayLike) -> None:
        x = _as_tensor(rows)
        if x.dim() != 2 or x.size(1) != self.dim:
            raise ValueError(f"expected [n, {self.dim}] rows, got {tuple(x.shape)}")
        need = self.size + x.size(0)
        if need > self.capacity:
            self.reserve(max(need, 2 * self.capacity))
        ptr, dtype, mem, keep = _triple(x, self.device)
        nat.check(self._lib.lk_index_add(self._h, c_void_p(ptr), dtype, mem, x.size(0), c_void_p(_stream(self.device))),
                  "lk_index_add")
        del keep

    # -- search -----------------------------------------------------------------------
    def set_timing(self, enabled: bool) -> None:
        nat.check(self._lib.lk_index_set_timing(self._h, int(bool(enabled))), "lk_index_set_timing")

    def last_timing(self) -> Tuple[float, float]:
        """(search kernel ms, whole device side ms) of the last search, CUDA events."""
        a, b = c_float(0), c_float(0)
        nat.check(self._lib.lk_index_last_timing(self._h, byref(a), byref(b)), "lk_index_last_timing")
        return float(a.value), float(b.value)

    def _host_out(self, b: int, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
        if self._pinned is None or self._pinned[0].numel() < b * k:
            n = max(b * k, 4096)
            self._pinned = (torch.empty(n, dtype=torch.float32, pin_memory=True),
                            torch.empty(n, dtype=torch.int64, pin_memory=True))
        return self._pinned[0][: b * k].view(b, k), self._pinned[1][: b * k].view(b, k)

    def search(self, queries: ArrayLike, k: int, *, idx_base: int = 0, kernel: str = "auto",
               device_out: bool = False):
        """Top-k of every query row: (scores float32 [B,k], indices int64 [B,k]), best first.

        Returns numpy arrays (host), or CUDA tensors when `device_out` (no host sync).
        Slots that cannot be filled (k > rows) hold index -1 / score -inf.
        """
        q = _as_tensor(queries)
        if q.dim() == 1:
            q = q.unsqueeze(0)
        if q.dim() != 2 or q.size(1) != self.dim:
            raise ValueError(f"expected [b, {self.dim}] queries, got {tuple(q.shape)}")
        if not 1 <= k <= nat.LK_MAX_K:
            # known incompatibility: the reference ranks a materialised [B, N] matrix and so
            # takes any k (retrieval/bruteforce.py:81-82); the engine's deepest search is LK_MAX_K
            raise ValueError(f"top-k of {k} is outside 1..{nat.LK_MAX_K} (LK_MAX_K), the deepest search the B200 "
                             "engine runs; ask for fewer neighbours or search row slabs and merge them with merge_topk")
        b = q.size(0)
        ptr, dtype, mem, keep = _triple(q, self.device)
        if device_out:
            d = torch.empty((b, k), dtype=torch.float32, device=f"cuda:{self.device}")
            i = torch.empty((b, k), dtype=torch.int64, device=f"cuda:{self.device}")
            out_mem = nat.LK_DEVICE
        else:
            d, i = self._host_out(b, k)
            out_mem = nat.LK_HOST
        nat.check(
            self._lib.lk_index_search(self._h, c_void_p(ptr), dtype, mem, b, int(k), c_void_p(d.data_ptr()),
                                      c_void_p(i.data_ptr()), out_mem, int(idx_base), nat.KERNELS[kernel],
                                      c_void_p(_stream(self.device))),
            "lk_index_search",
        )
        del keep
        if device_out:
            return d, i
        return d.numpy().copy(), i.numpy().copy()

    def check(self) -> None:
        """Synchronise and raise if a search with device outputs hit a pipeline timeout."""
        nat.check(self._lib.lk_index_check(self._h), "lk_index_check")

    # -- persistence ------------------------------------------------------------------
    def export_bytes(self) -> Tuple[np.ndarray, np.ndarray]:
        tb, sb = c_int64(0), c_int64(0)
        nat.check(self._lib.lk_index_storage_bytes(self._h, byref(tb), byref(sb)), "lk_index_storage_bytes")
        tiles = np.empty(tb.value, dtype=np.uint8)
        side = np.empty(sb.value // 4, dtype=np.float32)
        nat.check(self._lib.lk_index_export(self._h, c_void_p(tiles.ctypes.data), c_void_p(side.ctypes.data),
                                            c_void_p(_stream(self.device))), "lk_index_export")
        return tiles, side

    def import_bytes(self, tiles: np.ndarray, side: np.ndarray, n_rows: int) -> None:
        tiles = np.ascontiguousarray(tiles, dtype=np.uint8)
        side = np.ascontiguousarray(side, dtype=np.float32)
        n_rows = int(n_rows)
        want_t, want_s = self.image_bytes(n_rows)
        if n_rows < 0 or tiles.nbytes != want_t or side.nbytes != want_s:
            raise ValueError(f"index image does not hold {n_rows} rows of dim {self.dim}: {tiles.nbytes} tile bytes "
                             f"(need {want_t}), {side.nbytes} side bytes (need {want_s})")
        self.reserve(n_rows)
        nat.check(self._lib.lk_index_import(self._h, c_void_p(tiles.ctypes.data), int(tiles.nbytes),
                                            c_void_p(side.ctypes.data), int(side.nbytes), n_rows,
                                            c_void_p(_stream(self.device))), "lk_index_import")

    def image_bytes(self, n_rows: int) -> Tuple[int, int]:
        """(tile bytes, side bytes) of the persisted image of `n_rows` rows: whole 128-row blocks of
        K blocks of 128 bytes per row (include/latentknn.h, lk_index_storage_bytes)."""
        elem = 2 if self.storage == "bf16" else 4
        per_kb = 128 // elem
        kblocks = -(-self.dim // per_kb)
        nblk = -(-max(0, int(n_rows)) // 128)
        return nblk * kblocks * 128 * 128, nblk * 128 * 4


def merge_topk(cand_scores: ArrayLike, cand_idx: ArrayLike, k: int, device: int = 0):
    """k-way merge of [B, L, len] candidate lists (index < 0 = padding) on the GPU.
    numpy in -> numpy out; CUDA tensors in -> CUDA tensors out."""
    lib = nat.load()
    nat.require_device()
    s = cand_scores if torch.is_tensor(cand_scores) else torch.from_numpy(np.ascontiguousarray(cand_scores))
    i = cand_idx if torch.is_tensor(cand_idx) else torch.from_numpy(np.ascontiguousarray(cand_idx))
    s = s.to(torch.float32).contiguous()
    i = i.to(torch.int64).contiguous()
    if s.dim() == 2:
        s, i = s.unsqueeze(1), i.unsqueeze(1)
    b, n_lists, ln = s.shape
    on_dev = s.is_cuda
    if on_dev:
        device = s.device.index
        out_s = torch.empty((b, k), dtype=torch.float32, device=s.device)
        out_i = torch.empty((b, k), dtype=torch.int64, device=s.device)
    else:
        out_s = torch.empty((b, k), dtype=torch.float32)
        out_i = torch.empty((b, k), dtype=torch.int64)
    nat.check(
        lib.lk_merge_topk(device, c_void_p(s.data_ptr()), c_void_p(i.data_ptr()), b, n_lists, ln, int(k),
                          c_void_p(out_s.data_ptr()), c_void_p(out_i.data_ptr()),
                          nat.LK_DEVICE if on_dev else nat.LK_HOST, c_void_p(_stream(device))),
        "lk_merge_topk",
    )
    if on_dev:
        return out_s, out_i
    return out_s.numpy(), out_i.numpy()
