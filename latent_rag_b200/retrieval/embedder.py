"""EmbeddingCompressor: drop-in for retrieval/embedder.py:7-48 of the reference -- base sentence
embeddings, optionally compressed by an autoencoder, returned as a float32 CPU tensor [N, D].

The autoencoder forward runs on the fused B200 encoder kernel (latent_rag_b200.autoencoders).
The sentence encoder (SBERT all-MiniLM-L6-v2) is a third-party transformer with downloaded
weights: pass as `model` any object with the sentence-transformers `encode(texts, batch_size=...,
convert_to_tensor=True, normalize_embeddings=True)` method -- latent_rag_b200.SentenceEncoder
runs that forward on the B200 kernels from a BertModel state_dict and a tokenizer.  With `model`
None, a `base_model_name` that is a local checkpoint directory is loaded into a SentenceEncoder;
any other name constructs `SentenceTransformer(base_model_name)` when that package is installed.
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch


class EmbeddingCompressor:
    def __init__(self, base_model_name: str = "sentence-transformers/all-MiniLM-L6-v2", autoencoder=None,
                 device: Optional[str] = None, *, model=None):
        self.device = device or "cuda"
        if model is None and os.path.isdir(base_model_name) and os.path.exists(os.path.join(base_model_name, "config.json")):
            from ..sbert import SentenceEncoder

            model = SentenceEncoder.from_pretrained(base_model_name, device=torch.device(self.device).index)
        if model is None:
            try:
                from sentence_transformers import SentenceTransformer
            except ImportError as e:  # no silent substitute for the base encoder
                raise ImportError("EmbeddingCompressor needs sentence-transformers for the base embeddings, or a "
                                  "`model=` object with its encode() method") from e
            model = SentenceTransformer(base_model_name, device=self.device)
        self.model = model
        # embedder.py:20-22: autoencoder.to(device), eval()
        self.autoencoder = autoencoder.to(self.device) if autoencoder is not None else None
        if self.autoencoder is not None:
            self.autoencoder.eval()

    def encode_text(self, texts: List[str], compress: bool = True) -> torch.Tensor:
        """embedder.py:24-48: normalised base embeddings -> autoencoder.encode (a tuple means
        (mu, logvar): the mean is the latent code) -> float32 [N, D] on the CPU."""
        with torch.no_grad():
            embeddings = self.model.encode(texts, batch_size=64, convert_to_tensor=True, normalize_embeddings=True)
            embeddings = torch.as_tensor(embeddings).to(self.device)
            if self.autoencoder is not None and compress:
                encoded = self.autoencoder.encode(embeddings)
                if isinstance(encoded, tuple):
                    encoded = encoded[0]
                return encoded.cpu()
            return embeddings.float().cpu()
