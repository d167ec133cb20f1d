from .bruteforce import BruteForceRetriever  # noqa: F401
from .FAISSEmbeddingRetriever import FAISSEmbeddingRetriever  # noqa: F401
from .retriever import build_retriever  # noqa: F401
from .common import StatsTracker  # noqa: F401
from .embedder import EmbeddingCompressor  # noqa: F401
