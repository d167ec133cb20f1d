"""Seeded input generators shared by make_golden.py (which stores the reference's
outputs for them) and by the tests (which feed the same inputs to the oracle and to
the CUDA path).  numpy's PCG64 stream + float32 casts: deterministic on this image."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np
import torch

SMALL_CASES = [(100, 10, 64), (1000, 50, 32)]  # test/test_retrieval.py:61


def reference_test_embeddings(num: int, dim: int = 64, seed: int = 7) -> torch.Tensor:
    """The generator of the reference's own retrieval test (test/test_retrieval.py:33-38):
    default_rng(seed).standard_normal -> fp32 -> rows / (|row| + 1e-12)."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((num, dim)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True) + 1e-12
    return torch.from_numpy(x)


def bf16_round(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32)


def edge_case_inputs() -> Tuple[torch.Tensor, torch.Tensor]:
    """7 x 16 corpus whose row 3 is all zeros; 3 queries whose row 2 is all zeros.
    Values are multiples of 1/8 in [-2, 2): exact in bf16, so the fp32 reference and
    the bf16 engine see the same numbers."""
    rng = np.random.default_rng(11)
    emb = rng.integers(-16, 16, size=(7, 16)).astype(np.float32) / 8.0
    emb[3] = 0.0
    q = rng.integers(-16, 16, size=(3, 16)).astype(np.float32) / 8.0
    q[2] = 0.0
    return torch.from_numpy(emb), torch.from_numpy(q)


def mid_case_inputs(n: int = 4096, d: int = 384, b: int = 64) -> Tuple[torch.Tensor, torch.Tensor]:
    """bf16-representable Gaussian corpus/queries (not normalised: the cosine path
    normalises, the euclidean path must not); a quarter of the queries are perturbed
    corpus rows so near neighbours exist."""
    rng = np.random.default_rng(1234)
    emb = rng.standard_normal((n, d)).astype(np.float32)
    q = np.random.default_rng(4321).standard_normal((b, d)).astype(np.float32)
    pick = np.arange(0, b, 4)
    q[pick] = emb[(pick * 37) % n] + 0.1 * q[pick]
    return bf16_round(torch.from_numpy(emb)), bf16_round(torch.from_numpy(q))


def ae_input(m: int = 48, d: int = 384) -> torch.Tensor:
    """Unit-norm fp32 rows, like SBERT's normalize_embeddings=True output
    (retrieval/embedder.py:35-40)."""
    rng = np.random.default_rng(99)
    x = rng.standard_normal((m, d)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return torch.from_numpy(x)


def metrics_case(q: int = 40, k: int = 10, universe: int = 60) -> Tuple[List[List[int]], List[List[int]]]:
    rng = np.random.default_rng(5)
    retrieved = [rng.permutation(universe)[:k].tolist() for _ in range(q)]
    relevant = [rng.permutation(universe)[: int(rng.integers(1, 4))].tolist() for _ in range(q)]
    return retrieved, relevant


def metrics_ragged_case(q: int = 60, universe: int = 30) -> Tuple[List[List[int]], List[List[int]]]:
    """Ragged retrieved lists (1..12 ids, shorter than the cut-off 10 for most) against 0..11 relevant
    ids: covers len(retrieved) < k < len(relevant), where the ideal DCG keeps the caller's k
    (evaluation/retrieval_metrics.py:29)."""
    rng = np.random.default_rng(23)
    retrieved = [rng.integers(0, universe, int(rng.integers(1, 13))).tolist() for _ in range(q)]
    relevant = [rng.integers(0, universe, int(rng.integers(0, 12))).tolist() for _ in range(q)]
    retrieved[0], relevant[0] = [4, 9, 1, 7, 3], [9, 3, 11, 12, 13, 14, 15, 16]
    return retrieved, relevant


def maxsim_case(q: int = 30, ck: int = 30, docs: int = 12, seed: int = 91):
    """Candidates of `q` queries for the document-level MaxSim aggregation (main.py:264-282):
    descending scores with some exact ties, chunk -> doc ids with repeats.
    -> (scores [q, ck] float32, doc ids [q, ck] int64)"""
    rng = np.random.default_rng(seed)
    sc = -np.sort(-rng.standard_normal((q, ck)).astype(np.float32), axis=1)
    sc[::3, 4] = sc[::3, 3]  # equal scores: the stable sort keeps first-seen order
    did = rng.integers(0, docs, size=(q, ck)).astype(np.int64)
    did[1] = np.arange(ck) % 3  # only three documents among the candidates: fewer than top_k
    return sc, did


def rank_case(n: int = 200, d: int = 32, seed: int = 17):
    """Paired (query, document) embeddings for the positive-rank helper
    (evaluation/embedding_visualization.py:34-37): documents are noisy copies of the queries."""
    rng = np.random.default_rng(seed)
    q = torch.from_numpy(rng.standard_normal((n, d)).astype(np.float32))
    noise = torch.from_numpy(rng.standard_normal((n, d)).astype(np.float32))
    return q, q + 1.5 * noise


# ---------------------------------------------------------------------------------------
# sentence encoder: seeded BertModel weights and token batches
# ---------------------------------------------------------------------------------------
SBERT_SMALL = dict(vocab=200, max_pos=64, hidden=128, heads=4, ffn=256, layers=2, eps=1e-12)


def sbert_weights(cfg: dict, seed: int = 5):
    """A transformers BertModel state_dict (no pooler) with every tensor random: linear weights
    N(0, 1/sqrt(fan_in)) so activations keep unit scale through the stack, biases and LayerNorm
    offsets N(0, 0.1), LayerNorm gains 1 + N(0, 0.1).  torch CPU generator: deterministic on this image."""
    g = torch.Generator().manual_seed(seed)
    h, f = cfg["hidden"], cfg["ffn"]
    rn = lambda *shape, std=1.0: torch.randn(*shape, generator=g) * std  # noqa: E731
    w = {
        "embeddings.word_embeddings.weight": rn(cfg["vocab"], h),
        "embeddings.position_embeddings.weight": rn(cfg["max_pos"], h, std=0.5),
        "embeddings.token_type_embeddings.weight": rn(2, h, std=0.5),
        "embeddings.LayerNorm.weight": 1 + rn(h, std=0.1),
        "embeddings.LayerNorm.bias": rn(h, std=0.1),
    }
    for l in range(cfg["layers"]):
        p = f"encoder.layer.{l}."
        for name, (o, i) in {"attention.self.query": (h, h), "attention.self.key": (h, h),
                             "attention.self.value": (h, h), "attention.output.dense": (h, h),
                             "intermediate.dense": (f, h), "output.dense": (h, f)}.items():
            w[p + name + ".weight"] = rn(o, i, std=i ** -0.5)
            w[p + name + ".bias"] = rn(o, std=0.1)
        for name in ("attention.output.LayerNorm", "output.LayerNorm"):
            w[p + name + ".weight"] = 1 + rn(h, std=0.1)
            w[p + name + ".bias"] = rn(h, std=0.1)
    return w


def sbert_tokens(cfg: dict, n_sent: int, seq_len: int, seed: int = 6):
    """(input_ids int64 [n, s], attention_mask int64 [n, s]): ragged lengths 1..s, one full row."""
    rng = np.random.default_rng(seed)
    ids = rng.integers(0, cfg["vocab"], size=(n_sent, seq_len))
    lens = rng.integers(1, seq_len + 1, size=n_sent)
    lens[0] = seq_len
    mask = (np.arange(seq_len)[None, :] < lens[:, None]).astype(np.int64)
    ids = ids * mask  # padding id 0, like the WordPiece [PAD]
    return torch.from_numpy(ids.astype(np.int64)), torch.from_numpy(mask)


def write_sbert_checkpoint_dir(path: str, cfg: dict, weights, max_seq_length: int = 16, safetensors: bool = False):
    """A sentence-transformers style checkpoint directory (what `SentenceTransformer(name)` downloads):
    BertConfig, weights, a WordPiece vocabulary of cfg["vocab"] entries, the pooling / max-length side files."""
    import json
    import os

    os.makedirs(os.path.join(path, "1_Pooling"), exist_ok=True)
    json.dump({"architectures": ["BertModel"], "model_type": "bert", "vocab_size": cfg["vocab"],
               "hidden_size": cfg["hidden"], "num_hidden_layers": cfg["layers"], "num_attention_heads": cfg["heads"],
               "intermediate_size": cfg["ffn"], "max_position_embeddings": cfg["max_pos"], "layer_norm_eps": cfg["eps"],
               "hidden_act": "gelu", "position_embedding_type": "absolute", "type_vocab_size": 2},
              open(os.path.join(path, "config.json"), "w"))
    json.dump({"max_seq_length": max_seq_length, "do_lower_case": False}, open(os.path.join(path, "sentence_bert_config.json"), "w"))
    json.dump({"word_embedding_dimension": cfg["hidden"], "pooling_mode_cls_token": False, "pooling_mode_mean_tokens": True,
               "pooling_mode_max_tokens": False}, open(os.path.join(path, "1_Pooling", "config.json"), "w"))
    words = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]", "the", "quick", "brown", "fox", "##s", "jump", "over", "lazy", "dog"]
    words += [f"w{i}" for i in range(cfg["vocab"] - len(words))]
    open(os.path.join(path, "vocab.txt"), "w").write("\n".join(words) + "\n")
    json.dump({"tokenizer_class": "BertTokenizer", "do_lower_case": True, "model_max_length": cfg["max_pos"]},
              open(os.path.join(path, "tokenizer_config.json"), "w"))
    if safetensors:
        from safetensors.torch import save_file

        save_file({k: v.contiguous() for k, v in weights.items()}, os.path.join(path, "model.safetensors"))
    else:
        torch.save(dict(weights), os.path.join(path, "pytorch_model.bin"))
