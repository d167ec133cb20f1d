"""FAISSEmbeddingRetriever on the B200 engine: drop-in for the reference class
(retrieval/FAISSEmbeddingRetriever.py:20-345) for `index_type="flatip"` (exact inner
product over L2-normalised vectors == cosine).  The FAISS library itself is not used:
`index.add` / `index.search` are the liblatentknn kernels, persistence is this
package's own tiled image next to the same `<path>.meta.json` the reference writes.

`index_type="hnsw"` / `"ivfpq"` are approximate indexes in the reference; here they are
accepted for configuration compatibility and served by the same exact search (a superset
of their recall).  Anything else raises the reference's ValueError.
"""
from __future__ import annotations

import json
import os
import time
from pathlib import Path
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from ..engine import ExactIndex
from .common import BatchedRetrieveMixin, StatsTracker

_INDEX_TYPES = ("flatip", "hnsw", "ivfpq")
_MAGIC = b"LKNNIDX1"


class FAISSEmbeddingRetriever(BatchedRetrieveMixin):
    def __init__(
        self,
        embedding_dim: int,
        index_path: Optional[str | Path] = None,
        index_type: str = "hnsw",  # "flatip" | "hnsw" | "ivfpq"
        use_gpu: bool = False,
        *,
        hnsw_M: int = 32,
        ef_construction: int = 200,
        ef_search: int = 64,
        precision: str = "bf16",
        device: Optional[int] = None,
    ):
        self.d = int(embedding_dim)
        self.index_type = index_type
        self.path = Path(index_path) if index_path else None
        self.use_gpu = use_gpu
        self.hnsw_M, self.ef_construction, self.ef_search = int(hnsw_M), int(ef_construction), int(ef_search)
        self.precision = precision
        self.device = torch.cuda.current_device() if device is None else int(device)

        self._texts: List[str] = []
        self._doc_ids: List[int] = []
        self.meta_fp: Dict[str, Any] = {}
        self._stats = StatsTracker()

        self.index = self._build_index(self.d, self.index_type)
        # the engine only exists on the GPU: unlike the reference (FAISSEmbeddingRetriever.py:
        # 77-86) there is no CPU index to fall back to, so gpu_enabled is always True
        self.gpu_enabled = True

        if self.path and self.path.exists():
            try:
                self._load_index()
                self._load_metadata()
            except Exception:
                # corrupted or incompatible file -> start clean (FAISSEmbeddingRetriever.py:70-73)
                self.index = self._build_index(self.d, self.index_type)
                self._texts, self._doc_ids, self.meta_fp = [], [], {}

    # ------------------------------ helpers -------------------------------------------
    def _build_index(self, d: int, kind: str) -> ExactIndex:
        if kind not in _INDEX_TYPES:
            raise ValueError(f"Index type not supported: {kind}")  # FAISSEmbeddingRetriever.py:103
        return ExactIndex(d, 1024, metric="cosine", storage=self.precision, device=self.device)

    def _meta_path(self) -> Path:
        assert self.path is not None
        return self.path.with_suffix(self.path.suffix + ".meta.json")

    def _save_metadata(self) -> None:  # FAISSEmbeddingRetriever.py:110-124
        if not self.path:
            return
        meta = {"texts": self._texts, "doc_ids": self._doc_ids, "fingerprint": self.meta_fp}
        self._meta_path().parent.mkdir(parents=True, exist_ok=True)
        with self._meta_path().open("w", encoding="utf-8") as f:
            json.dump(meta, f, ensure_ascii=False)

    def _load_metadata(self) -> None:  # FAISSEmbeddingRetriever.py:126-137
        if not self.path:
            return
        mp = self._meta_path()
        if not mp.exists():
            self._texts, self._doc_ids, self.meta_fp = [], [], {}
            return
        with mp.open("r", encoding="utf-8") as f:
            meta = json.load(f)
        self._texts = list(meta.get("texts", []))
        self._doc_ids = list(meta.get("doc_ids", []))
        self.meta_fp = dict(meta.get("fingerprint", {}))

    def _save_index(self) -> None:
        """Own on-disk image: magic, JSON header, raw tiles, raw side values.  Written to a
        temporary file and renamed into place, so a crash mid-save never leaves a truncated image
        under the index path."""
        tiles, side = self.index.export_bytes()
        header = json.dumps({"d": self.d, "n": self.index.size, "precision": self.precision,
                             "tile_bytes": int(tiles.nbytes), "side_bytes": int(side.nbytes)}).encode()
        self.path.parent.mkdir(parents=True, exist_ok=True)
        tmp = self.path.with_name(self.path.name + f".tmp{os.getpid()}")
        try:
            with tmp.open("wb") as f:
                f.write(_MAGIC)
                f.write(len(header).to_bytes(8, "little"))
                f.write(header)
                f.write(tiles.tobytes())
                f.write(side.tobytes())
                f.flush()
                os.fsync(f.fileno())
            os.replace(tmp, self.path)
        finally:
            if tmp.exists():
                tmp.unlink()

    def _load_index(self) -> None:
        """Raises ValueError on anything that is not a complete image of this geometry (the
        constructor then starts clean, FAISSEmbeddingRetriever.py:70-73): the header is never
        trusted further than the bytes actually present."""
        with self.path.open("rb") as f:
            if f.read(8) != _MAGIC:
                raise ValueError("not a latentknn index file")
            hlen = int.from_bytes(f.read(8), "little")
            if not 0 < hlen <= 1 << 20:
                raise ValueError("corrupt index header")
            hdr = json.loads(f.read(hlen))
            if hdr["d"] != self.d or hdr["precision"] != self.precision:
                raise ValueError("index file does not match this retriever")
            n = int(hdr["n"])
            index = self._build_index(self.d, self.index_type)
            want_t, want_s = index.image_bytes(n)
            if n < 0 or int(hdr["tile_bytes"]) != want_t or int(hdr["side_bytes"]) != want_s:
                raise ValueError("index header is inconsistent (rows vs payload sizes)")
            tiles = np.frombuffer(f.read(want_t), dtype=np.uint8)
            side_raw = f.read(want_s)
            if tiles.nbytes != want_t or len(side_raw) != want_s:
                raise ValueError("index file is truncated")
            side = np.frombuffer(side_raw, dtype=np.float32)
        index.import_bytes(tiles, side, n)  # checks the lengths again on both sides of the C ABI
        self.index = index

    @staticmethod
    def _fingerprint(*, d: int, embedding_model: Optional[str], ae_type: Optional[str], latent_dim: Optional[int],
                     chunking_cfg: Optional[Dict[str, Any]], metric: str = "ip", normalize_l2: bool = True,
                     version: int = 1) -> Dict[str, Any]:
        # same keys and defaults as FAISSEmbeddingRetriever.py:139-167
        ch = chunking_cfg or {}
        return {
            "d": int(d),
            "embedding_model": embedding_model,
            "ae_type": ae_type,
            "latent_dim": int(latent_dim) if latent_dim is not None else None,
            "chunking": {
                "enabled": bool(ch.get("enabled", False)),
                "mode": ch.get("mode", "sliding"),
                "max_tokens": int(ch.get("max_tokens", 128)) if ch.get("max_tokens") is not None else None,
                "stride": int(ch.get("stride", 64)) if ch.get("stride") is not None else None,
                "min_tokens": int(ch.get("min_tokens", 48)) if ch.get("min_tokens") is not None else None,
            },
            "metric": metric,
            "normalize_l2": bool(normalize_l2),
            "version": int(version),
        }

    def _compatible(self, current_fp: Dict[str, Any]) -> bool:  # FAISSEmbeddingRetriever.py:169-179
        m = self.meta_fp or {}
        for key in ["d", "embedding_model", "ae_type", "latent_dim", "metric", "normalize_l2", "version"]:
            if m.get(key) != current_fp.get(key):
                return False
        mch, cch = (m.get("chunking") or {}), (current_fp.get("chunking") or {})
        return all(mch.get(key) == cch.get(key) for key in ["enabled", "mode", "max_tokens", "stride", "min_tokens"])

    # ------------------------------- build --------------------------------------------
    def build(
        self,
        embeddings: torch.Tensor,
        texts: Sequence[str],
        doc_ids: Sequence[int] | None = None,
        train: bool = True,
        *,
        embedding_model_name: Optional[str] = None,
        ae_type: Optional[str] = None,
        latent_dim: Optional[int] = None,
        chunking_cfg: Optional[Dict[str, Any]] = None,
    ) -> None:
        """Build (or extend) the index and attach metadata (FAISSEmbeddingRetriever.py:190-311).
        `train` is accepted for signature compatibility; an exact index has nothing to train."""
        assert len(embeddings) == len(texts), "len mismatch (embeddings vs texts)"
        if doc_ids is not None:
            assert len(texts) == len(doc_ids), "len mismatch (texts vs doc_ids)"
        if isinstance(embeddings, np.ndarray):
            embeddings = torch.from_numpy(embeddings)

        cur_fp = self._fingerprint(d=int(embeddings.shape[1]), embedding_model=embedding_model_name, ae_type=ae_type,
                                   latent_dim=latent_dim, chunking_cfg=chunking_cfg)
        rebuild = self.index.dim != cur_fp["d"]
        if (self.path and self.path.exists()) and not self._compatible(cur_fp):
            rebuild = True
        if rebuild:
            self.d = cur_fp["d"]
            self.index = self._build_index(self.d, self.index_type)
            self._texts, self._doc_ids, self.meta_fp = [], [], {}

        # normalise + add: the engine keeps the rows and 1/|row| and applies the norm in the
        # kernel epilogue, which is faiss.normalize_L2 + IndexFlatIP.add in one pass
        t0 = time.perf_counter()
        first_new = self.index.size
        self.index.add(embeddings)
        torch.cuda.synchronize(self.device)
        self._stats.add_build_time(time.perf_counter() - t0)

        # minimal sanity check: self-search of the first vector (FAISSEmbeddingRetriever.py:259-292)
        self._sanity_ok = True
        if len(embeddings) and first_new == 0 and float(embeddings[0].float().abs().sum()) > 0:
            _, i_chk = self.index.search(embeddings[:1], 1)
            self._sanity_ok = bool(i_chk[0, 0] == 0)
            if not self._sanity_ok:
                print("[ERROR] index sanity check failed: row 0 is not its own nearest neighbour")

        self._texts.extend(list(texts))
        self._doc_ids.extend(list(doc_ids) if doc_ids is not None else [-1] * len(texts))
        self.meta_fp = cur_fp

        if self.path:
            self._save_index()
            self._save_metadata()
        print(f"[LKNN] type={self.index_type}->exact d={cur_fp['d']} ntotal={self.index.size} metric=IP normL2=True")

    # ------------------------------ search --------------------------------------------
    def search(self, queries: torch.Tensor, k: int) -> Tuple[np.ndarray, np.ndarray]:
        """(D, I) like IndexFlatIP.search (FAISSEmbeddingRetriever.py:314-326): when k exceeds
        the number of rows the tail is padded with index -1 / score -FLT_MAX [upstream]."""
        if isinstance(queries, np.ndarray):
            queries = torch.from_numpy(queries)
        if queries.dim() == 1:
            queries = queries.unsqueeze(0)
        k = int(k)
        b, n = queries.size(0), self.index.size
        d = np.full((b, k), -np.finfo(np.float32).max, dtype=np.float32)
        i = np.full((b, k), -1, dtype=np.int64)
        kk = min(k, n)
        if b == 0 or kk < 1:
            return d, i
        t0 = time.perf_counter()
        dd, ii = self.index.search(queries, kk)
        self._stats.add_search_batch(batch_size=b, seconds=time.perf_counter() - t0)
        d[:, :kk], i[:, :kk] = dd, ii
        return d, i

    def _search_device(self, queries: torch.Tensor, k: int):
        """`search` that leaves (scores, row ids) on the device, clamped to the rows present: for
        retrieve_batch."""
        if isinstance(queries, np.ndarray):
            queries = torch.from_numpy(queries)
        if queries.dim() == 1:
            queries = queries.unsqueeze(0)
        if not self._doc_ids:
            self._load_metadata()
        kk = min(int(k), self.index.size)
        dev = torch.device(f"cuda:{self.index.device}")
        if queries.size(0) == 0 or kk < 1:
            return (torch.empty((queries.size(0), max(kk, 0)), dtype=torch.float32, device=dev),
                    torch.empty((queries.size(0), max(kk, 0)), dtype=torch.int64, device=dev))
        t0 = time.perf_counter()
        d, i = self.index.search(queries, kk, device_out=True)
        self._stats.add_search_batch(batch_size=queries.size(0), seconds=time.perf_counter() - t0)
        return d, i

    def retrieve(self, query_emb: torch.Tensor, top_k: int = 10) -> Tuple[List[str], List[float], List[int]]:
        d, i = self.search(query_emb, top_k)
        idxs = i[0].tolist()
        if not self._texts or not self._doc_ids:  # lazy metadata load (FAISSEmbeddingRetriever.py:333-334)
            self._load_metadata()
        texts = [self._texts[j] for j in idxs]
        scores = d[0].tolist()
        docids = [self._doc_ids[j] for j in idxs]
        return texts, scores, docids

    def get_stats(self, reset: bool = False) -> Dict[str, Any]:
        return self._stats.get_stats(reset=reset)
