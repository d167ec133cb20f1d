"""build_retriever: the switch the reference's pipeline flips (retrieval/retriever.py:17-34,
called at main.py:248).  Same cfg keys: backend, index_path, index_type, use_gpu."""
from __future__ import annotations

from typing import Sequence

import torch

from .bruteforce import BruteForceRetriever
from .FAISSEmbeddingRetriever import FAISSEmbeddingRetriever


def build_retriever(embeddings: torch.Tensor, texts: Sequence[str], doc_ids: Sequence[int], cfg: dict):
    if cfg.get("backend", "faiss") == "faiss":
        ret = FAISSEmbeddingRetriever(
            embedding_dim=embeddings.size(1),
            index_path=cfg.get("index_path"),
            index_type=cfg.get("index_type", "hnsw"),
            use_gpu=cfg.get("use_gpu", False),
            **({"precision": cfg["precision"]} if "precision" in cfg else {}),
        )
        ret.build(embeddings, texts, doc_ids, train=True)
        return ret
    extra = {key: cfg[key] for key in ("metric", "precision") if key in cfg}
    return BruteForceRetriever(embeddings, texts, doc_ids, **extra)
