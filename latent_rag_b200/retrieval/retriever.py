"""build_retriever: the switch the reference's pipeline flips (retrieval/retriever.py:17-34,
called at main.py:248).  Same cfg keys and defaults -- backend ("faiss"), index_path (None),
index_type ("hnsw"), use_gpu (False) -- plus the optional `metric` / `precision` of this engine."""
from __future__ import annotations

from typing import Any, Dict, Mapping, Sequence

import torch

from .bruteforce import BruteForceRetriever
from .FAISSEmbeddingRetriever import FAISSEmbeddingRetriever

# cfg key -> (constructor keyword, default) of the FAISS mirror
_FAISS_KEYS = {"index_path": ("index_path", None), "index_type": ("index_type", "hnsw"), "use_gpu": ("use_gpu", False)}


def _picked(cfg: Mapping[str, Any], names: Sequence[str]) -> Dict[str, Any]:
    return {name: cfg[name] for name in names if name in cfg}


def build_retriever(embeddings: torch.Tensor, texts: Sequence[str], doc_ids: Sequence[int], cfg: dict):
    backend = cfg.get("backend", "faiss")
    if backend != "faiss":  # the reference sends every other value to the brute-force class
        return BruteForceRetriever(embeddings, texts, doc_ids, **_picked(cfg, ("metric", "precision")))
    kwargs = {kw: cfg.get(key, default) for key, (kw, default) in _FAISS_KEYS.items()}
    retriever = FAISSEmbeddingRetriever(embeddings.size(1), **kwargs, **_picked(cfg, ("precision",)))
    retriever.build(embeddings, texts, doc_ids, train=True)
    return retriever
