"""SentenceEncoder: the forward of the sentence-embedding model on the B200 kernels.

Stands where the reference holds `SentenceTransformer("all-MiniLM-L6-v2")`
(retrieval/embedder.py:17-18) and calls `.encode(texts, batch_size=64, convert_to_tensor=True,
normalize_embeddings=True)` (embedder.py:35-40): a BERT encoder (transformers BertModel
state_dict), masked mean pooling, L2 normalisation -- liblatentknn's lk_bert_* (tcgen05 linear
layers with split-bf16 operands, fp32 attention / LayerNorm / pooling kernels).

Tokenisation is host work with a third-party vocabulary: pass `tokenizer`, a callable
`texts -> {"input_ids": [n, s], "attention_mask": [n, s]}` (e.g. a transformers tokenizer
partially applied with padding=True, truncation=True, max_length=256, return_tensors="pt"), or
feed token ids to `encode_tokens` directly.  There is no CPU path.
"""
from __future__ import annotations

import ctypes
import json
import os
from ctypes import POINTER, Structure, byref, c_float, c_void_p
from typing import Callable, Dict, List, Mapping, Optional, Sequence, Union

import numpy as np
import torch

from . import _native as nat

MINILM_L6 = dict(vocab=30522, max_pos=512, hidden=384, heads=12, ffn=1536, layers=6, eps=1e-12)

_FP = POINTER(c_float)


class _LayerWeights(Structure):
    _fields_ = [(n, _FP) for n in ("wq", "bq", "wk", "bk", "wv", "bv", "wo", "bo", "ln1_g", "ln1_b", "w1", "b1",
                                   "w2", "b2", "ln2_g", "ln2_b")]


class _Weights(Structure):
    _fields_ = [("word_emb", _FP), ("pos_emb", _FP), ("type_emb", _FP), ("emb_ln_g", _FP), ("emb_ln_b", _FP),
                ("layers", POINTER(_LayerWeights))]


_LAYER_KEYS = {
    "wq": "attention.self.query.weight", "bq": "attention.self.query.bias",
    "wk": "attention.self.key.weight", "bk": "attention.self.key.bias",
    "wv": "attention.self.value.weight", "bv": "attention.self.value.bias",
    "wo": "attention.output.dense.weight", "bo": "attention.output.dense.bias",
    "ln1_g": "attention.output.LayerNorm.weight", "ln1_b": "attention.output.LayerNorm.bias",
    "w1": "intermediate.dense.weight", "b1": "intermediate.dense.bias",
    "w2": "output.dense.weight", "b2": "output.dense.bias",
    "ln2_g": "output.LayerNorm.weight", "ln2_b": "output.LayerNorm.bias",
}
_EMB_KEYS = {
    "word_emb": "embeddings.word_embeddings.weight", "pos_emb": "embeddings.position_embeddings.weight",
    "type_emb": "embeddings.token_type_embeddings.weight", "emb_ln_g": "embeddings.LayerNorm.weight",
    "emb_ln_b": "embeddings.LayerNorm.bias",
}


def _f32(t) -> np.ndarray:
    if torch.is_tensor(t):
        t = t.detach().to("cpu", torch.float32).numpy()
    return np.ascontiguousarray(np.asarray(t, dtype=np.float32))


def length_sorted_chunks(texts: Sequence[str], step: int) -> List[List[int]]:
    """Indices of `texts`, longest first (stable), in runs of `step`: every run is tokenised and padded
    on its own, so a run's padding is bounded by its own longest sentence (sentence-transformers
    sorts the same way inside `encode`)."""
    order = sorted(range(len(texts)), key=lambda i: -len(texts[i]))
    return [order[lo: lo + step] for lo in range(0, len(order), step)]


class SentenceEncoder:
    """state_dict: the tensors of a transformers BertModel (keys may carry a `bert.` or `0.auto_model.`
    prefix, as sentence-transformers checkpoints do); the architecture is read from their shapes.

    heads            attention heads (the kernels implement head dimension 32: all-MiniLM-L6-v2 has 12)
    precision        "fp32" (default; split-bf16 operands, fp32-level results) | "bf16"
    max_seq_length   `encode` truncates to it (sentence-transformers ships all-MiniLM-L6-v2 with 256)
    """

    def __init__(self, state_dict: Mapping[str, object], *, heads: Optional[int] = None, ln_eps: float = 1e-12,
                 device: Optional[int] = None, tokenizer: Optional[Callable] = None, max_seq_length: int = 256,
                 precision: str = "fp32"):
        self._lib = nat.load()
        nat.require_device()
        sd = self._strip_prefix(state_dict)
        for key in _EMB_KEYS.values():
            if key not in sd:
                raise KeyError(f"Missing key in state_dict: {key}")
        word = _f32(sd[_EMB_KEYS["word_emb"]])
        self.vocab, self.hidden = int(word.shape[0]), int(word.shape[1])
        self.max_pos = int(sd[_EMB_KEYS["pos_emb"]].shape[0])
        self.layers = 0
        while f"encoder.layer.{self.layers}.{_LAYER_KEYS['wq']}" in sd:
            self.layers += 1
        if self.layers == 0:
            raise KeyError("Missing key in state_dict: encoder.layer.0.attention.self.query.weight")
        self.ffn = int(sd[f"encoder.layer.0.{_LAYER_KEYS['w1']}"].shape[0])
        self.heads = int(heads) if heads is not None else self.hidden // 32
        self.ln_eps = float(ln_eps)
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.tokenizer = tokenizer
        self.max_seq_length = min(int(max_seq_length), self.max_pos)

        keep: List[np.ndarray] = []  # host arrays stay alive until lk_bert_create has copied them

        def ptr(t, shape):
            a = _f32(t)
            if a.shape != tuple(shape):
                raise ValueError(f"weight of shape {a.shape}, expected {tuple(shape)}")
            keep.append(a)
            return a.ctypes.data_as(_FP)

        h, f = self.hidden, self.ffn
        lw = (_LayerWeights * self.layers)()
        shapes = {"wq": (h, h), "wk": (h, h), "wv": (h, h), "wo": (h, h), "w1": (f, h), "w2": (h, f), "b1": (f,)}
        for l in range(self.layers):
            for field, key in _LAYER_KEYS.items():
                full = f"encoder.layer.{l}.{key}"
                if full not in sd:
                    raise KeyError(f"Missing key in state_dict: {full}")
                setattr(lw[l], field, ptr(sd[full], shapes.get(field, (h,))))
        w = _Weights()
        w.word_emb = ptr(word, (self.vocab, h))
        w.pos_emb = ptr(sd[_EMB_KEYS["pos_emb"]], (self.max_pos, h))
        type_emb = _f32(sd[_EMB_KEYS["type_emb"]])
        w.type_emb = ptr(type_emb[0], (h,))
        w.emb_ln_g = ptr(sd[_EMB_KEYS["emb_ln_g"]], (h,))
        w.emb_ln_b = ptr(sd[_EMB_KEYS["emb_ln_b"]], (h,))
        w.layers = ctypes.cast(lw, POINTER(_LayerWeights))
        self._h = c_void_p()
        nat.check(self._lib.lk_bert_create(byref(self._h), self.device, self.vocab, self.max_pos, h, self.heads, f,
                                           self.layers, c_float(self.ln_eps), byref(w)), "lk_bert_create")
        del keep
        self.set_precision(precision)

    # -- checkpoint directories ------------------------------------------------------------
    @staticmethod
    def read_checkpoint_dir(path: str) -> dict:
        """What a local sentence-transformers / transformers checkpoint directory says about the model:
        {"heads", "ln_eps", "max_seq_length", "weights": path} -- and a ValueError for anything the
        kernels do not implement (a non-GELU activation, relative positions, a pooling other than the mean)."""
        with open(os.path.join(path, "config.json")) as f:
            cfg = json.load(f)
        if cfg.get("hidden_act", "gelu") != "gelu":
            raise ValueError(f"hidden_act={cfg['hidden_act']!r}: only the erf GELU of BERT is implemented")
        if cfg.get("position_embedding_type", "absolute") != "absolute":
            raise ValueError("only absolute position embeddings are implemented")
        max_seq = int(cfg.get("max_position_embeddings", 512))
        sb = os.path.join(path, "sentence_bert_config.json")
        if os.path.exists(sb):
            with open(sb) as f:
                max_seq = min(max_seq, int(json.load(f).get("max_seq_length") or max_seq))
        pool = os.path.join(path, "1_Pooling", "config.json")
        if os.path.exists(pool):
            with open(pool) as f:
                pc = json.load(f)
            others = [k for k, v in pc.items() if k.startswith("pooling_mode_") and v and k != "pooling_mode_mean_tokens"]
            if not pc.get("pooling_mode_mean_tokens", True) or others:
                raise ValueError(f"only mean pooling is implemented (1_Pooling/config.json: {pc})")
        for name in ("model.safetensors", "pytorch_model.bin"):
            if os.path.exists(os.path.join(path, name)):
                weights = os.path.join(path, name)
                break
        else:
            raise FileNotFoundError(f"no model.safetensors / pytorch_model.bin under {path}")
        return {"heads": int(cfg["num_attention_heads"]), "ln_eps": float(cfg.get("layer_norm_eps", 1e-12)),
                "max_seq_length": max_seq, "weights": weights}

    @classmethod
    def from_pretrained(cls, path: str, *, device: Optional[int] = None, precision: str = "fp32",
                        tokenizer: Optional[Callable] = None) -> "SentenceEncoder":
        """A LOCAL checkpoint directory (config.json + model.safetensors / pytorch_model.bin, as
        `SentenceTransformer(name)` downloads it; nothing is fetched).  Without `tokenizer` the
        directory's own one is loaded through transformers and applied with padding to the longest
        sentence of a call and truncation to the model's max_seq_length."""
        info = cls.read_checkpoint_dir(path)
        if info["weights"].endswith(".safetensors"):
            from safetensors.torch import load_file

            sd = load_file(info["weights"])
        else:
            sd = torch.load(info["weights"], map_location="cpu", weights_only=True)
        if tokenizer is None:
            from transformers import AutoTokenizer

            tok = AutoTokenizer.from_pretrained(path, local_files_only=True)
            max_len = info["max_seq_length"]

            def tokenizer(texts):
                return tok(list(texts), padding=True, truncation=True, max_length=max_len, return_tensors="pt")

        return cls(sd, heads=info["heads"], ln_eps=info["ln_eps"], device=device, tokenizer=tokenizer,
                   max_seq_length=info["max_seq_length"], precision=precision)

    @staticmethod
    def _strip_prefix(state_dict: Mapping[str, object]) -> Dict[str, object]:
        for prefix in ("", "bert.", "0.auto_model.", "auto_model."):
            if prefix + _EMB_KEYS["word_emb"] in state_dict:
                return {k[len(prefix):]: v for k, v in state_dict.items() if k.startswith(prefix)}
        raise KeyError(f"Missing key in state_dict: {_EMB_KEYS['word_emb']}")

    # -- lifetime ----------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.lk_bert_destroy(self._h)
            self._h = c_void_p()

    def __del__(self):  # pragma: no cover - best effort
        try:
            self.close()
        except Exception:
            pass

    def set_precision(self, precision: str) -> "SentenceEncoder":
        if precision not in ("fp32", "bf16"):
            raise ValueError(f"Unsupported precision: {precision}")
        nat.check(self._lib.lk_bert_set_precision(self._h, nat.LK_F32 if precision == "fp32" else nat.LK_BF16),
                  "lk_bert_set_precision")
        self.precision = precision
        return self

    def get_sentence_embedding_dimension(self) -> int:  # sentence-transformers spelling
        return self.hidden

    # -- forward -----------------------------------------------------------------------
    def encode_tokens(self, input_ids, attention_mask, normalize_embeddings: bool = True) -> torch.Tensor:
        """[n, s] token ids + mask (torch or numpy, host or this GPU) -> float32 [n, hidden] on the GPU."""
        ids = torch.as_tensor(input_ids)
        mask = torch.as_tensor(attention_mask)
        if ids.dim() != 2 or ids.shape != mask.shape:
            raise ValueError(f"expected [n, s] ids and mask, got {tuple(ids.shape)} and {tuple(mask.shape)}")
        dev = torch.device(f"cuda:{self.device}")
        if ids.is_cuda != mask.is_cuda:
            ids, mask = ids.to(dev), mask.to(dev)
        ids = ids.to(torch.int32).contiguous()
        mask = mask.to(torch.int32).contiguous()
        n, s = ids.shape
        out = torch.empty((n, self.hidden), dtype=torch.float32, device=dev)
        mem = nat.LK_DEVICE if ids.is_cuda else nat.LK_HOST
        nat.check(self._lib.lk_bert_encode(self._h, c_void_p(ids.data_ptr()), c_void_p(mask.data_ptr()), mem, n, s,
                                           int(bool(normalize_embeddings)), c_void_p(out.data_ptr()), nat.LK_DEVICE,
                                           c_void_p(int(torch.cuda.current_stream(self.device).cuda_stream))),
                  "lk_bert_encode")
        return out

    def check(self) -> None:
        """Synchronise and raise if a linear-layer kernel hit a pipeline timeout."""
        nat.check(self._lib.lk_bert_check(self._h), "lk_bert_check")

    def encode(self, sentences: Union[str, Sequence[str]], batch_size: int = 64, show_progress_bar: bool = False,
               convert_to_tensor: bool = False, convert_to_numpy: bool = True, normalize_embeddings: bool = False,
               device=None, **_ignored):
        """sentence-transformers' `encode` for the arguments the reference passes (embedder.py:35-40).
        Sentences are sorted by length and padded per batch, like upstream, so padding work stays small;
        `batch_size` here only bounds one tokenizer call (the kernels take the whole padded batch)."""
        if self.tokenizer is None:
            raise RuntimeError("SentenceEncoder.encode needs a tokenizer (texts -> input_ids / attention_mask); "
                               "pass tokenizer=... or call encode_tokens")
        single = isinstance(sentences, str)
        texts = [sentences] if single else list(sentences)
        out = torch.empty((len(texts), self.hidden), dtype=torch.float32, device=f"cuda:{self.device}")
        for pick in length_sorted_chunks(texts, max(int(batch_size), 1) * 16):
            tok = self.tokenizer([texts[i] for i in pick])
            ids = torch.as_tensor(tok["input_ids"])[:, : self.max_seq_length]
            mask = torch.as_tensor(tok["attention_mask"])[:, : self.max_seq_length]
            out[torch.as_tensor(pick, device=out.device)] = self.encode_tokens(ids, mask, normalize_embeddings)
        self.check()
        if single:
            out = out[0]
        if convert_to_tensor:
            return out
        return out.cpu().numpy() if convert_to_numpy else [row for row in out.cpu()]
