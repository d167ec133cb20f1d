"""Parity of the sentence-encoder forward (lk_bert_*, lk_linear_forward) with the CPU oracle and
with the committed outputs of transformers' own BertModel (tests/golden/sbert_golden.npz).
Run on a B200: pytest -m gpu.

Tolerance: the linear layers carry fp32 operands as two bf16 planes (16 mantissa bits, three
MMAs per product, fp32 accumulate): 2e-5 of the output row's scale per layer; the L2-normalised
sentence embeddings agree with the fp32 reference to 1e-4 absolute (cosine > 0.99999).  With
precision="bf16" (operands rounded to bf16 once per layer) the stated bar is cosine > 0.999.
"""
from ctypes import c_void_p

import numpy as np
import pytest
import torch

import oracle
from tests.golden import inputs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lrb():
    import latent_rag_b200 as m

    m._native.require_device()
    return m


def _linear(lrb, x, w, bias, residual, act, precision="fp32"):
    nat = lrb._native
    m, k = x.shape
    n = w.shape[0]
    y = np.empty((m, n), dtype=np.float32)
    ptr = lambda a: c_void_p(a.ctypes.data) if a is not None else None  # noqa: E731
    nat.check(nat.load().lk_linear_forward(0, ptr(x), m, k, ptr(w), n, ptr(bias), ptr(residual), act,
                                           nat.LK_F32 if precision == "fp32" else nat.LK_BF16, ptr(y)), "lk_linear_forward")
    return y


@pytest.mark.parametrize("m,k,n", [(1, 64, 128), (128, 384, 384), (300, 384, 1152), (1000, 1536, 384), (257, 384, 1536),
                                   (20000, 128, 256)])
def test_linear_layer_matches_fp32_reference(lrb, m, k, n):
    rng = np.random.default_rng(m + k + n)
    x = rng.standard_normal((m, k)).astype(np.float32)
    w = (rng.standard_normal((n, k)) / np.sqrt(k)).astype(np.float32)
    b = rng.standard_normal(n).astype(np.float32)
    r = rng.standard_normal((m, n)).astype(np.float32)
    ref = torch.from_numpy(x).double() @ torch.from_numpy(w).double().T + torch.from_numpy(b).double()
    y = _linear(lrb, x, w, b, None, 0)
    scale = ref.abs().amax(dim=1, keepdim=True).numpy()
    assert (np.abs(y - ref.numpy()) / scale).max() < 2e-5
    y = _linear(lrb, x, w, None, r, 0)
    assert (np.abs(y - (ref - torch.from_numpy(b).double() + torch.from_numpy(r).double()).numpy()) / scale).max() < 2e-5
    y = _linear(lrb, x, w, b, r, 1)
    ref_g = torch.nn.functional.gelu(ref) + torch.from_numpy(r).double()
    assert (np.abs(y - ref_g.numpy()) / scale).max() < 2e-5
    y16 = _linear(lrb, x, w, b, None, 0, precision="bf16")
    ref16 = (oracle.bf16_round(torch.from_numpy(x)).double() @ oracle.bf16_round(torch.from_numpy(w)).double().T
             + torch.from_numpy(b).double())
    assert (np.abs(y16 - ref16.numpy()) / scale).max() < 2e-5


def test_linear_layer_rejects_unsupported_shapes(lrb):
    x = np.zeros((4, 100), np.float32)
    with pytest.raises(lrb.NativeError):
        _linear(lrb, x, np.zeros((128, 100), np.float32), None, None, 0)
    with pytest.raises(lrb.NativeError):
        _linear(lrb, np.zeros((4, 64), np.float32), np.zeros((100, 64), np.float32), None, None, 0)


CASES = {"small": (inputs.SBERT_SMALL, 9, 40), "minilm": (oracle.MINILM_L6, 6, 24)}


@pytest.mark.parametrize("name", ["small", "minilm"])
def test_encoder_matches_transformers_outputs(lrb, golden, name):
    """Seeded weights and tokens; the committed outputs come from transformers' BertModel + mean pooling +
    normalisation (tests/golden/make_golden_sbert.py)."""
    cfg, n, s = CASES[name]
    g = golden("sbert_golden.npz")
    w = inputs.sbert_weights(cfg)
    ids, mask = inputs.sbert_tokens(cfg, n, s)
    enc = lrb.SentenceEncoder(w, heads=cfg["heads"], ln_eps=cfg["eps"])
    assert (enc.hidden, enc.ffn, enc.layers, enc.vocab, enc.max_pos) == (cfg["hidden"], cfg["ffn"], cfg["layers"],
                                                                          cfg["vocab"], cfg["max_pos"])
    emb = enc.encode_tokens(ids, mask, normalize_embeddings=True).cpu().numpy()
    enc.check()
    assert np.abs(emb - g[f"{name}_emb"]).max() < 1e-4
    assert (emb * g[f"{name}_emb"]).sum(1).min() > 0.99999
    pooled = enc.encode_tokens(ids.cuda(), mask.cuda(), normalize_embeddings=False).cpu().numpy()
    scale = np.abs(g[f"{name}_pooled"]).max(axis=1, keepdims=True)
    assert (np.abs(pooled - g[f"{name}_pooled"]) / scale).max() < 2e-4
    ref = oracle.sbert_encode(w, cfg, ids, mask).numpy()
    assert np.abs(emb - ref).max() < 1e-4
    # padding does not leak: a sentence encodes the same alone, without its padding
    ln = int(mask[1].sum())
    alone = enc.encode_tokens(ids[1:2, :ln], mask[1:2, :ln]).cpu().numpy()
    assert np.abs(alone - emb[1:2]).max() < 2e-5
    enc16 = enc.set_precision("bf16").encode_tokens(ids, mask).cpu().numpy()
    assert (enc16 * g[f"{name}_emb"]).sum(1).min() > 0.999


def test_encoder_batches_longer_than_one_pass(lrb):
    """More tokens than one pass of the layer stack takes (65536): passes are independent."""
    cfg = inputs.SBERT_SMALL
    w = inputs.sbert_weights(cfg)
    ids, mask = inputs.sbert_tokens(cfg, 1500, 64, seed=8)
    enc = lrb.SentenceEncoder(w, heads=cfg["heads"])
    emb = enc.encode_tokens(ids, mask).cpu().numpy()
    enc.check()
    ref = oracle.sbert_encode(w, cfg, ids, mask).numpy()
    assert np.abs(emb - ref).max() < 1e-4
    host = enc.encode_tokens(ids.numpy(), mask.numpy()).cpu().numpy()
    np.testing.assert_array_equal(host, emb)


@pytest.mark.parametrize("s", [64, 100, 128, 200, 256, 300])
def test_attention_kernels_over_sequence_lengths(lrb, s, monkeypatch):
    """64..256 tokens take the register-tiled attention kernel, shorter and longer sentences the
    warp-per-query one; both against the oracle, ragged masks, lengths that are no multiple of 16."""
    cfg = dict(inputs.SBERT_SMALL, max_pos=320)
    w = inputs.sbert_weights(cfg)
    ids, mask = inputs.sbert_tokens(cfg, 7, s, seed=s)
    enc = lrb.SentenceEncoder(w, heads=cfg["heads"])
    ref = oracle.sbert_encode(w, cfg, ids, mask).numpy()
    emb = enc.encode_tokens(ids, mask).cpu().numpy()
    enc.check()
    assert np.abs(emb - ref).max() < 1e-4
    if s <= 256:
        monkeypatch.setenv("LK_ATTN", "warp")
        emb_w = enc.encode_tokens(ids, mask).cpu().numpy()
        assert np.abs(emb_w - ref).max() < 1e-4
        assert np.abs(emb_w - emb).max() < 2e-5


def test_encode_contract_of_the_reference_caller(lrb):
    """EmbeddingCompressor.encode_text (retrieval/embedder.py:24-48) drives the encoder through
    `encode(texts, batch_size=64, convert_to_tensor=True, normalize_embeddings=True)`."""
    cfg = inputs.SBERT_SMALL
    w = {"bert." + k: v for k, v in inputs.sbert_weights(cfg).items()}  # prefixed checkpoints load too

    def toy_tokenizer(texts):  # whitespace words -> ids; padded to the longest of the call
        rows = [[1 + (sum(map(ord, wd)) % (cfg["vocab"] - 1)) for wd in t.split()] or [1] for t in texts]
        s = max(map(len, rows))
        ids = np.zeros((len(rows), s), np.int64)
        mask = np.zeros((len(rows), s), np.int64)
        for r, row in enumerate(rows):
            ids[r, : len(row)] = row
            mask[r, : len(row)] = 1
        return {"input_ids": ids, "attention_mask": mask}

    enc = lrb.SentenceEncoder(w, heads=cfg["heads"], tokenizer=toy_tokenizer, max_seq_length=32)
    texts = ["the quick brown fox", "jumps", "over the lazy dog again and again", "", "a b c d e f g h i j k l m n o p q r s t u v w x y z " * 3]
    out = enc.encode(texts, batch_size=64, convert_to_tensor=True, normalize_embeddings=True)
    assert out.is_cuda and out.shape == (5, cfg["hidden"]) and out.dtype == torch.float32
    np.testing.assert_allclose(out.norm(dim=1).cpu().numpy(), 1.0, atol=1e-5)
    plain = {k[5:]: v for k, v in w.items()}
    for r, t in enumerate(texts):
        tok = toy_tokenizer([t])
        ids, mask = torch.from_numpy(tok["input_ids"])[:, :32], torch.from_numpy(tok["attention_mask"])[:, :32]
        ref = oracle.sbert_encode(plain, cfg, ids, mask).numpy()
        assert np.abs(out[r].cpu().numpy() - ref[0]).max() < 1e-4
    comp = lrb.EmbeddingCompressor(model=enc, device="cuda")
    z = comp.encode_text(texts, compress=False)
    assert z.device.type == "cpu" and z.shape == (5, cfg["hidden"])
    np.testing.assert_allclose(z.numpy(), out.cpu().numpy(), atol=1e-6)
    assert enc.encode("one sentence").shape == (cfg["hidden"],)
    with pytest.raises(RuntimeError):
        lrb.SentenceEncoder(plain, heads=cfg["heads"]).encode(["no tokenizer"])


@pytest.mark.parametrize("safetensors", [False, True])
def test_from_pretrained_directory(lrb, tmp_path, safetensors):
    """A local sentence-transformers checkpoint directory (config, weights, WordPiece vocabulary) loads into
    the encoder with its own tokenizer; EmbeddingCompressor takes the directory as `base_model_name`, where the
    reference passes the model name to SentenceTransformer (retrieval/embedder.py:17-18)."""
    from transformers import AutoTokenizer

    cfg = inputs.SBERT_SMALL
    w = inputs.sbert_weights(cfg)
    d = str(tmp_path / "ckpt")
    inputs.write_sbert_checkpoint_dir(d, cfg, w, max_seq_length=16, safetensors=safetensors)
    enc = lrb.SentenceEncoder.from_pretrained(d)
    assert (enc.heads, enc.max_seq_length, enc.layers, enc.vocab) == (cfg["heads"], 16, cfg["layers"], cfg["vocab"])
    texts = ["the quick brown foxs jump over the lazy dog", "dog", "w7 w8 w9 unknownword the", "lazy " * 40]
    out = enc.encode(texts, batch_size=64, convert_to_tensor=True, normalize_embeddings=True)
    tok = AutoTokenizer.from_pretrained(d, local_files_only=True)
    for r, t in enumerate(texts):
        one = tok([t], padding=True, truncation=True, max_length=16, return_tensors="pt")
        ref = oracle.sbert_encode(w, cfg, one["input_ids"], one["attention_mask"]).numpy()
        assert np.abs(out[r].cpu().numpy() - ref[0]).max() < 1e-4
    comp = lrb.EmbeddingCompressor(base_model_name=d, device="cuda")
    z = comp.encode_text(texts, compress=False)
    np.testing.assert_allclose(z.numpy(), out.cpu().numpy(), atol=1e-6)


def test_tokens_to_neighbours_pipeline(lrb):
    """The reference's whole embedding + retrieval chain on the device (main.py: corpus chunks ->
    EmbeddingCompressor.encode_text = SBERT forward + autoencoder.encode -> retriever; queries
    the same way -> retrieve): sentence encoder (MiniLM architecture, seeded weights), the shipped
    contrastive autoencoder, cosine search over the latents -- against the oracle chain."""
    import os

    cfg = oracle.MINILM_L6
    w = inputs.sbert_weights(cfg)
    ids, mask = inputs.sbert_tokens(cfg, 600, 48, seed=11)
    enc = lrb.SentenceEncoder(w, heads=cfg["heads"])
    gold = os.path.join(os.path.dirname(__file__), "golden", "ae_weights_cae.npz")
    ae = lrb.load_autoencoder("cae", gold, device=0)
    emb = enc.encode_tokens(ids, mask, normalize_embeddings=True)
    z = ae.encode(emb)  # [600, 64] unit-norm latents, on the device
    enc.check()
    r = lrb.BruteForceRetriever(z, [f"chunk {j}" for j in range(600)], None, metric="cosine", precision="fp32")
    d, i = r.search(z[:40], 5)
    np.testing.assert_array_equal(i[:, 0], np.arange(40))
    emb_ref = oracle.sbert_encode(w, cfg, ids, mask)
    assert (emb.cpu() - emb_ref).abs().max() < 1e-4
    z_ref = oracle.ae_encode(emb_ref, oracle.load_encoder_weights(np.load(gold), "cae"), "cae")
    assert (z.cpu() - z_ref).abs().max() < 2e-4
    d_ref, i_ref = oracle.bruteforce_search(oracle.bruteforce_build(z_ref), z_ref[:40], 5)
    ok, why = oracle.topk_equivalent(d_ref, i_ref, d, i, rtol=1e-3)  # inputs differ by the encoders' 1e-4
    assert ok, why


def test_encoder_rejects_what_the_kernels_do_not_implement(lrb):
    cfg = dict(inputs.SBERT_SMALL, heads=2)  # head dimension 64
    with pytest.raises(lrb.NativeError):
        lrb.SentenceEncoder(inputs.sbert_weights(inputs.SBERT_SMALL), heads=cfg["heads"])
    enc = lrb.SentenceEncoder(inputs.sbert_weights(inputs.SBERT_SMALL), heads=4)
    with pytest.raises(lrb.NativeError):  # more tokens than position embeddings
        enc.encode_tokens(np.zeros((1, 65), np.int64), np.ones((1, 65), np.int64))
    with pytest.raises(KeyError):
        lrb.SentenceEncoder({"embeddings.word_embeddings.weight": torch.zeros(4, 128)})
