"""latent_rag_b200 -- B200-native exact nearest-neighbour search and autoencoder encoder
forward behind latent-rag's retriever API.  Hand-written sm_100a CUDA (liblatentknn.so)
called through a C ABI; Python/PyTorch only for device memory, streams and
torch.distributed.  No CPU fallback."""
from . import _native
from ._native import NativeError
from .engine import ExactIndex, merge_topk
from .retrieval import (BruteForceRetriever, EmbeddingCompressor, FAISSEmbeddingRetriever, StatsTracker,
                        build_retriever)
from .autoencoders import (ContrastiveAutoencoder, DenoisingAutoencoder, VariationalAutoencoder,
                           load_autoencoder)
from .sharded import ShardedRetriever, shard_bounds
from .exchange import PeerExchange
from .evaluation import evaluate_retrieval, rank_positive
from .sbert import SentenceEncoder

__all__ = [
    "ExactIndex", "merge_topk", "BruteForceRetriever", "EmbeddingCompressor", "FAISSEmbeddingRetriever", "StatsTracker",
    "build_retriever", "ContrastiveAutoencoder", "DenoisingAutoencoder", "VariationalAutoencoder",
    "load_autoencoder", "ShardedRetriever", "shard_bounds", "NativeError", "PeerExchange", "evaluate_retrieval", "rank_positive", "SentenceEncoder",
]
