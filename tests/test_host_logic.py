"""Host-side logic that needs no GPU: stats, error conventions, fingerprints, and the rule
that the product fails loudly (never falls back) without a device."""
import os

import numpy as np
import pytest
import torch

import latent_rag_b200 as lrb
from latent_rag_b200 import _native
from latent_rag_b200.retrieval.common import StatsTracker, empirical_precision, whitener_from_precision
from latent_rag_b200.retrieval.FAISSEmbeddingRetriever import FAISSEmbeddingRetriever

NO_GPU = not torch.cuda.is_available()


def test_stats_tracker_contract():
    """retrieval/common.py:37-65 of the reference."""
    s = StatsTracker()
    s.add_build_time(0.5)
    s.add_search_batch(batch_size=4, seconds=0.2)
    s.add_search_batch(batch_size=0, seconds=0.1)
    out = s.get_stats()
    assert set(out) == {"build_time_s", "search_time_s", "search_calls", "per_query_ms"}
    assert out["build_time_s"] == 0.5 and out["search_calls"] == 2
    assert abs(out["search_time_s"] - 0.3) < 1e-12
    assert out["per_query_ms"] == [50.0, 100.0]
    s.get_stats(reset=True)
    assert s.get_stats() == {"build_time_s": 0.0, "search_time_s": 0.0, "search_calls": 0, "per_query_ms": []}


def test_len_mismatch_is_an_assertion_error():
    """test/test_retrieval.py:122-128 of the reference."""
    emb = torch.randn(10, 8)
    with pytest.raises(AssertionError):
        lrb.BruteForceRetriever(emb, ["a"] * 10, doc_ids=[0, 1])


def test_unsupported_metric_is_a_value_error():
    with pytest.raises(ValueError, match="Unsupported metric"):
        lrb.BruteForceRetriever(torch.randn(4, 8), ["a"] * 4, None, metric="manhattan")


@pytest.mark.skipif(not NO_GPU, reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback_without_a_device():
    with pytest.raises(_native.NativeError, match="no CUDA device"):
        lrb.ExactIndex(8, 16)
    with pytest.raises(_native.NativeError):
        lrb.merge_topk(np.zeros((1, 2, 3), np.float32), np.zeros((1, 2, 3), np.int64), 2)
    ae = lrb.DenoisingAutoencoder(8, 2, 4)
    ae.load_state_dict({"encoder.0.weight": np.zeros((4, 8)), "encoder.0.bias": np.zeros(4),
                        "encoder.2.weight": np.zeros((2, 4)), "encoder.2.bias": np.zeros(2)})
    with pytest.raises(_native.NativeError):
        ae.encode(torch.zeros(3, 8))


def test_product_never_imports_the_oracle():
    import os
    import re

    pkg = os.path.dirname(lrb.__file__)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"


def test_fingerprint_and_compat_rules():
    """FAISSEmbeddingRetriever.py:139-179 of the reference."""
    fp = FAISSEmbeddingRetriever._fingerprint(d=64, embedding_model="m", ae_type="vae", latent_dim=64,
                                              chunking_cfg={"enabled": True, "max_tokens": 128})
    assert fp["metric"] == "ip" and fp["normalize_l2"] is True and fp["version"] == 1
    # absent keys stay None in the reference (`... if ch.get(key) is not None else None`)
    assert fp["chunking"] == {"enabled": True, "mode": "sliding", "max_tokens": 128, "stride": None,
                              "min_tokens": None}
    r = FAISSEmbeddingRetriever.__new__(FAISSEmbeddingRetriever)
    r.meta_fp = fp
    assert r._compatible(dict(fp))
    other = dict(fp, ae_type="dae")
    assert not r._compatible(other)
    other = dict(fp, chunking=dict(fp["chunking"], stride=32))
    assert not r._compatible(other)


def test_autoencoder_state_dict_contract():
    ae = lrb.VariationalAutoencoder(16, 4, 8)
    with pytest.raises(KeyError):
        ae.load_state_dict({"encoder.0.weight": np.zeros((8, 16))})
    with pytest.raises(RuntimeError, match="size mismatch"):
        ae.load_state_dict({"encoder.0.weight": np.zeros((8, 15)), "encoder.0.bias": np.zeros(8),
                            "mu_layer.weight": np.zeros((4, 8)), "mu_layer.bias": np.zeros(4)})
    with pytest.raises(ValueError):
        lrb.load_autoencoder("gan", {})
    with pytest.raises(FileNotFoundError):
        lrb.load_autoencoder("vae", "/nonexistent/ckpt.pth")


def test_empirical_precision_matches_oracle():
    import oracle

    rng = np.random.default_rng(0)
    x = torch.from_numpy(rng.standard_normal((300, 12)).astype(np.float32) * np.linspace(0.5, 2, 12, dtype=np.float32))
    p = empirical_precision(x, chunk=64)
    np.testing.assert_allclose(p, oracle.mahalanobis_precision(x), rtol=1e-8, atol=1e-10)
    lw = whitener_from_precision(p)
    np.testing.assert_allclose(lw @ lw.T, p, rtol=1e-9, atol=1e-10)


def test_shard_bounds():
    import oracle

    for n, w in [(10, 3), (1000, 8), (7, 8), (0, 2), (128, 1)]:
        b = lrb.shard_bounds(n, w)
        assert b == list(oracle.shard_bounds(n, w))
        assert b[0][0] == 0 and b[-1][1] == n
        assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))


def test_metric_input_encoding_and_names():
    """Host half of the device metrics (latent_rag_b200/evaluation.py): ids of any hashable kind are
    coded densely and consistently on both sides, ragged lists are padded with -1, the relevant
    lists become a CSR; metric names parse like evaluation/retrieval_metrics.py:38-39."""
    from latent_rag_b200.evaluation import _encode, _parse_metric

    ret, off, rel = _encode([["a", "b", "a"], ["c"], []], [["b", "z"], [], ["a"]])
    assert ret.shape == (3, 3) and ret.dtype == np.int64
    assert ret[0].tolist() == [0, 1, 0] and ret[1].tolist() == [2, -1, -1] and ret[2].tolist() == [-1, -1, -1]
    assert off.tolist() == [0, 2, 2, 3]
    assert rel.tolist()[0] == 1 and rel.tolist()[2] == 0  # "b" and "a" keep their codes; "z" gets a new one
    assert _parse_metric("Recall@10") == ("Recall", 10) and _parse_metric("mrr") == ("mrr", None)
    with pytest.raises(ValueError, match="No metrics"):
        lrb.evaluate_retrieval([[1]], [[1]], [])


@pytest.mark.skipif(not NO_GPU, reason="only meaningful on a box without a GPU")
def test_new_entry_points_fail_loudly_without_a_device():
    with pytest.raises(_native.NativeError):
        lrb.evaluate_retrieval([[1, 2]], [[2]], ["mrr"])
    with pytest.raises(_native.NativeError):
        lrb.rank_positive(torch.randn(4, 8), torch.randn(4, 8))
    with pytest.raises(_native.NativeError):
        lrb.PeerExchange(0, 0, 1)


@pytest.mark.skipif(not NO_GPU, reason="only meaningful on a box without a GPU")
def test_sentence_encoder_fails_loudly_without_a_device():
    from tests.golden import inputs

    with pytest.raises(_native.NativeError, match="no CUDA device"):
        lrb.SentenceEncoder(inputs.sbert_weights(inputs.SBERT_SMALL), heads=4)


def test_sentence_encoder_state_dict_prefixes():
    """transformers BertModel keys, bare or as sentence-transformers / BertFor* checkpoints carry them."""
    from latent_rag_b200.sbert import SentenceEncoder

    sd = {"embeddings.word_embeddings.weight": 1, "encoder.layer.0.output.dense.bias": 2, "pooler.dense.weight": 3}
    assert SentenceEncoder._strip_prefix(sd) == sd
    for prefix in ("bert.", "0.auto_model.", "auto_model."):
        got = SentenceEncoder._strip_prefix({prefix + k: v for k, v in sd.items()} | {"cls.predictions.bias": 0})
        assert got == sd
    with pytest.raises(KeyError, match="word_embeddings"):
        SentenceEncoder._strip_prefix({"encoder.layer.0.output.dense.bias": 2})


def test_sentence_batches_are_length_sorted():
    from latent_rag_b200.sbert import length_sorted_chunks

    texts = ["bb", "a", "dddd", "ccc", "", "eeeee", "ff"]
    chunks = length_sorted_chunks(texts, 3)
    assert chunks == [[5, 2, 3], [0, 6, 1], [4]]  # longest first, ties in input order
    assert sorted(i for c in chunks for i in c) == list(range(len(texts)))
    assert length_sorted_chunks([], 4) == []


def test_sentence_encoder_reads_checkpoint_directories(tmp_path):
    """config.json / sentence_bert_config.json / 1_Pooling/config.json of a sentence-transformers
    directory: what is read, and what is refused because the kernels do not implement it."""
    import json

    from latent_rag_b200.sbert import SentenceEncoder
    from tests.golden import inputs

    cfg = inputs.SBERT_SMALL
    d = str(tmp_path / "ckpt")
    inputs.write_sbert_checkpoint_dir(d, cfg, {"embeddings.word_embeddings.weight": torch.zeros(2, 2)}, max_seq_length=16)
    info = SentenceEncoder.read_checkpoint_dir(d)
    assert (info["heads"], info["ln_eps"], info["max_seq_length"]) == (cfg["heads"], cfg["eps"], 16)
    assert info["weights"].endswith("pytorch_model.bin")
    from transformers import AutoTokenizer

    tok = AutoTokenizer.from_pretrained(d, local_files_only=True)
    out = tok(["the quick brown foxs jump", "dog"], padding=True, truncation=True, max_length=6, return_tensors="pt")
    assert out["input_ids"].shape == (2, 6) and out["attention_mask"][1].tolist() == [1, 1, 1, 0, 0, 0]
    edit = lambda name, **kw: json.dump({**json.load(open(os.path.join(d, name))), **kw}, open(os.path.join(d, name), "w"))  # noqa: E731
    edit("config.json", hidden_act="relu")
    with pytest.raises(ValueError, match="GELU"):
        SentenceEncoder.read_checkpoint_dir(d)
    edit("config.json", hidden_act="gelu")
    edit(os.path.join("1_Pooling", "config.json"), pooling_mode_cls_token=True, pooling_mode_mean_tokens=False)
    with pytest.raises(ValueError, match="mean pooling"):
        SentenceEncoder.read_checkpoint_dir(d)
    edit(os.path.join("1_Pooling", "config.json"), pooling_mode_cls_token=False, pooling_mode_mean_tokens=True)
    os.remove(os.path.join(d, "pytorch_model.bin"))
    with pytest.raises(FileNotFoundError):
        SentenceEncoder.read_checkpoint_dir(d)


def test_embedding_compressor_needs_a_base_encoder():
    """Without sentence-transformers and without `model=` (e.g. a latent_rag_b200.SentenceEncoder)
    the class says so instead of substituting something else."""
    try:
        import sentence_transformers  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError, match="sentence-transformers"):
            lrb.EmbeddingCompressor()


def test_build_retriever_reads_the_reference_cfg_keys(monkeypatch):
    """retrieval/retriever.py:17-34: backend defaults to faiss (index_path None, index_type hnsw,
    use_gpu False, build(train=True)); any other backend builds the brute-force class."""
    import latent_rag_b200.retrieval.retriever as mod

    calls = []

    class Rec:
        def __init__(self, *a, **k):
            calls.append((type(self).__name__, a[0] if not torch.is_tensor(a[0]) else "emb", k))

        def build(self, *a, **k):
            calls.append(("build", len(a), k))

    monkeypatch.setattr(mod, "FAISSEmbeddingRetriever", type("FA", (Rec,), {}))
    monkeypatch.setattr(mod, "BruteForceRetriever", type("BR", (Rec,), {}))
    emb = torch.zeros(3, 8)
    mod.build_retriever(emb, ["a"] * 3, [1, 2, 3], {})
    mod.build_retriever(emb, ["a"] * 3, [1, 2, 3], {"backend": "faiss", "index_type": "flatip", "index_path": "/x",
                                                     "use_gpu": True, "precision": "fp32"})
    mod.build_retriever(emb, ["a"] * 3, [1, 2, 3], {"backend": "bruteforce", "metric": "euclidean"})
    assert calls == [
        ("FA", 8, {"index_path": None, "index_type": "hnsw", "use_gpu": False}), ("build", 3, {"train": True}),
        ("FA", 8, {"index_path": "/x", "index_type": "flatip", "use_gpu": True, "precision": "fp32"}),
        ("build", 3, {"train": True}),
        ("BR", "emb", {"metric": "euclidean"}),
    ]


def test_sharded_retriever_rejects_unknown_exchange():
    import oracle

    emb = torch.randn(20, 8)
    fake = lambda q, k: oracle.bruteforce_search(oracle.bruteforce_build(emb), q, k)
    with pytest.raises(ValueError, match="Unknown exchange"):
        lrb.ShardedRetriever(emb, 0, "cosine", local_search=fake, merge=oracle.merge_topk, exchange="mpi")
