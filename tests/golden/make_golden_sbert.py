"""Golden outputs of the sentence-encoder forward, produced by transformers' OWN BertModel.

Run in the build container (transformers 5.5, CPU):

    python tests/golden/make_golden_sbert.py

Writes sbert_golden.npz: for two architectures -- a small one and all-MiniLM-L6-v2's
(config.json: hidden 384, 12 heads, ffn 1536, 6 layers, 512 positions, eps 1e-12) -- the
seeded weights of tests/golden/inputs.py are loaded into BertModel(add_pooling_layer=False),
run on the seeded token batches, mean-pooled over the attention mask the way
sentence_transformers.models.Pooling does and L2-normalised
(retrieval/embedder.py:35-40: normalize_embeddings=True).  Weights and tokens are regenerated
from their seeds by the tests, so only outputs are stored.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import torch.nn.functional as F
from transformers import BertConfig, BertModel

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from tests.golden import inputs  # noqa: E402

MINILM_L6 = dict(vocab=30522, max_pos=512, hidden=384, heads=12, ffn=1536, layers=6, eps=1e-12)
CASES = {"small": (inputs.SBERT_SMALL, 9, 40), "minilm": (MINILM_L6, 6, 24)}


def hf_encode(cfg, w, ids, mask):
    hf_cfg = BertConfig(vocab_size=cfg["vocab"], hidden_size=cfg["hidden"], num_hidden_layers=cfg["layers"],
                        num_attention_heads=cfg["heads"], intermediate_size=cfg["ffn"],
                        max_position_embeddings=cfg["max_pos"], layer_norm_eps=cfg["eps"], hidden_act="gelu",
                        type_vocab_size=2)
    model = BertModel(hf_cfg, add_pooling_layer=False).eval()
    missing, unexpected = model.load_state_dict(w, strict=False)
    assert not unexpected and all("position_ids" in k or "token_type_ids" in k for k in missing), (missing, unexpected)
    with torch.no_grad():
        hidden = model(input_ids=ids, attention_mask=mask).last_hidden_state
    m = mask.to(torch.float32).unsqueeze(-1)
    pooled = (hidden * m).sum(1) / torch.clamp(m.sum(1), min=1e-9)
    return hidden, pooled, F.normalize(pooled, p=2, dim=1)


def main():
    out = {}
    for name, (cfg, n, s) in CASES.items():
        w = inputs.sbert_weights(cfg)
        ids, mask = inputs.sbert_tokens(cfg, n, s)
        hidden, pooled, emb = hf_encode(cfg, w, ids, mask)
        out[f"{name}_hidden_row0"] = hidden[0].numpy()  # the unpadded sentence's token states
        out[f"{name}_pooled"] = pooled.numpy()
        out[f"{name}_emb"] = emb.numpy()
    np.savez_compressed(os.path.join(HERE, "sbert_golden.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
