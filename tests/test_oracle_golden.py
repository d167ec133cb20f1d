"""Pin the CPU oracle against outputs of the real reference (tests/golden/*, written by
tests/golden/make_golden.py) and against the KATs in the reference's own tests."""
import json
import os

import numpy as np
import pytest
import torch

import oracle
from tests.golden import inputs

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("n,nq,dim", inputs.SMALL_CASES)
@pytest.mark.parametrize("metric", ["cosine", "euclidean"])
def test_oracle_matches_reference_test_vectors(golden, n, nq, dim, metric):
    g = golden("retrieval_small.npz")
    emb = inputs.reference_test_embeddings(n, dim)
    idx = oracle.bruteforce_build(emb, metric)
    d, i = oracle.bruteforce_search(idx, emb[:nq], 5, metric)
    np.testing.assert_array_equal(i, g[f"{n}_{nq}_{dim}_{metric}_I"])
    np.testing.assert_allclose(d, g[f"{n}_{nq}_{dim}_{metric}_D"], rtol=1e-6, atol=1e-6)
    assert d.dtype == np.float32 and i.dtype == np.int64
    # implied KAT of test/test_retrieval.py:61-83: query i's best hit is row i
    np.testing.assert_array_equal(i[:, 0], np.arange(nq))
    if metric == "cosine":
        np.testing.assert_allclose(d[:, 0], 1.0, atol=1e-6)


def test_oracle_survey_kat_first_query(golden):
    """SURVEY.md section 8c (observed by running the reference): (100,64) q0 ->
    [0,97,25,17,66], scores [1.0000002, 0.390545, 0.32753217, 0.23455197, 0.22197998]."""
    emb = inputs.reference_test_embeddings(100, 64)
    d, i = oracle.bruteforce_search(oracle.bruteforce_build(emb), emb[0], 5)
    assert i[0].tolist() == [0, 97, 25, 17, 66]
    np.testing.assert_allclose(d[0], [1.0000002, 0.390545, 0.32753217, 0.23455197, 0.22197998], rtol=1e-6)
    g = golden("retrieval_small.npz")
    assert g["100_10_64_retrieve_ids"].tolist() == [0, 97, 25, 17, 66]


@pytest.mark.parametrize("metric", ["cosine", "euclidean"])
def test_oracle_edge_cases(golden, metric):
    g = golden("retrieval_edge.npz")
    emb, q = inputs.edge_case_inputs()
    idx = oracle.bruteforce_build(emb, metric)
    d, i = oracle.bruteforce_search(idx, q, 50, metric)  # k > N clamps to N
    assert d.shape == (3, 7)
    ok, why = oracle.topk_equivalent(g[f"clamp_{metric}_D"], g[f"clamp_{metric}_I"], d, i)
    assert ok, why
    d1, i1 = oracle.bruteforce_search(idx, q[1], 3, metric)  # 1-D query
    assert d1.shape == (1, 3)
    ok, why = oracle.topk_equivalent(g[f"oned_{metric}_D"], g[f"oned_{metric}_I"], d1, i1)
    assert ok, why
    if metric == "cosine":
        assert np.all(d[2] == 0.0)  # zero query scores 0 against everything


@pytest.mark.parametrize("metric", ["cosine", "euclidean"])
def test_oracle_mid_case(golden, metric):
    g = golden("retrieval_mid.npz")
    emb, q = inputs.mid_case_inputs()
    d, i = oracle.bruteforce_search(oracle.bruteforce_build(emb, metric), q, 10, metric)
    ok, why = oracle.topk_equivalent(g[f"{metric}_D"], g[f"{metric}_I"], d, i)
    assert ok, why


def test_oracle_unsupported_metric():
    with pytest.raises(ValueError):
        oracle.bruteforce_build(torch.zeros(2, 4), "manhattan")


@pytest.mark.parametrize("n,nq,dim", inputs.SMALL_CASES)
def test_faiss_flatip_restatement_equals_bruteforce(n, nq, dim):
    """test/test_retrieval.py:61-83 defines FlatIP == brute force on ordered ids."""
    emb = inputs.reference_test_embeddings(n, dim)
    xb = oracle.faiss_flatip_build(emb)
    df, if_ = oracle.faiss_flatip_search(xb, emb[:nq], 5)
    db, ib = oracle.bruteforce_search(oracle.bruteforce_build(emb), emb[:nq], 5)
    np.testing.assert_array_equal(if_, ib)
    np.testing.assert_allclose(df, db, rtol=1e-6, atol=1e-6)
    # [upstream] k > ntotal pads with -1 instead of clamping
    dpad, ipad = oracle.faiss_flatip_search(xb[:3], emb[:2], 5)
    assert ipad.shape == (2, 5) and (ipad[:, 3:] == -1).all()


@pytest.mark.parametrize("kind", ["vae", "dae", "cae"])
def test_oracle_autoencoders_match_reference(golden, kind):
    w = oracle.load_encoder_weights(golden(f"ae_weights_{kind}.npz"), kind)
    assert tuple(w["w0"].shape) == (512, 384) and tuple(w["w1"].shape) == (64, 512)
    z = oracle.ae_encode(inputs.ae_input(), w, kind).numpy()
    np.testing.assert_allclose(z, golden("ae_golden.npz")[f"{kind}_z"], rtol=1e-5, atol=1e-6)
    if kind == "cae":  # test/test_models.py:26-36
        np.testing.assert_allclose(np.linalg.norm(z, axis=1), 1.0, atol=1e-6)


def test_oracle_metrics_kats_and_reference_outputs():
    with open(os.path.join(GOLDEN, "metrics_golden.json")) as f:
        g = json.load(f)
    # test/test_evaluation.py:9-22
    assert oracle.recall_at_k([1, 2, 3, 4, 5], [3, 4, 6], 3) == 1 / 3 == g["kat"]["recall_at_3"]
    assert oracle.mrr([1, 2, 3, 4, 5], [3, 4, 6]) == 1 / 3 == g["kat"]["mrr"]
    assert oracle.ndcg_at_k([1, 2, 3, 4, 5], [3, 4, 6], 3) == g["kat"]["ndcg_at_3"]
    assert abs(g["kat"]["ndcg_at_3"] - 0.23463936301137822) < 1e-15
    retrieved, relevant = inputs.metrics_case()
    res = oracle.evaluate_retrieval(retrieved, relevant, list(g["evaluate_retrieval"]))
    for name, ref in g["evaluate_retrieval"].items():
        assert res[name]["mean"] == ref["mean"] and res[name]["std"] == ref["std"], name
    # per-query values of the reference on ragged lists (len(retrieved) < k < len(relevant) included): bit for bit
    rr, rl = inputs.metrics_ragged_case()
    rg = g["ragged_per_query"]
    for t, (r, l) in enumerate(zip(rr, rl)):
        assert oracle.ndcg_at_k(r, l, 10) == rg["ndcg@10"][t] and oracle.ndcg_at_k(r, l, 3) == rg["ndcg@3"][t]
        assert oracle.recall_at_k(r, l, 10) == rg["recall@10"][t] and oracle.mrr(r, l) == rg["mrr"][t]


def test_mahalanobis_oracle_self_consistency():
    """Unpinned by the reference: check the two forms of our own definition agree
    (precision form vs Cholesky-whitened euclidean), SURVEY.md section 8c."""
    rng = np.random.default_rng(3)
    a = np.diag(np.linspace(0.2, 2.0, 24)) @ np.linalg.qr(rng.standard_normal((24, 24)))[0]
    e = torch.from_numpy((rng.standard_normal((500, 24)) @ a).astype(np.float32))
    q = torch.from_numpy((rng.standard_normal((9, 24)) @ a).astype(np.float32))
    p = oracle.mahalanobis_precision(e)
    d, i = oracle.mahalanobis_search(e, q, 5, p)
    lw = oracle.mahalanobis_whitener(p)
    ew = torch.from_numpy(e.numpy().astype(np.float64) @ lw)
    qw = torch.from_numpy(q.numpy().astype(np.float64) @ lw)
    dist = ((qw[:, None, :] - ew[None, :, :]) ** 2).sum(-1)
    vals, idx = torch.topk(-dist, 5, dim=1)
    np.testing.assert_array_equal(idx.numpy(), i)
    np.testing.assert_allclose(vals.numpy(), d, rtol=1e-5)


def test_merge_topk_equals_global_topk():
    rng = np.random.default_rng(8)
    emb = torch.from_numpy(rng.standard_normal((1000, 32)).astype(np.float32))
    q = torch.from_numpy(rng.standard_normal((6, 32)).astype(np.float32))
    full_d, full_i = oracle.bruteforce_search(oracle.bruteforce_build(emb, "euclidean"), q, 10, "euclidean")
    for world in (1, 2, 3, 8):
        cd, ci = [], []
        for lo, hi in oracle.shard_bounds(1000, world):
            d, i = oracle.bruteforce_search(emb[lo:hi].contiguous(), q, 10, "euclidean")
            cd.append(d)
            ci.append(i + lo)
        d, i = oracle.merge_topk(np.concatenate(cd, 1), np.concatenate(ci, 1), 10)
        np.testing.assert_array_equal(i, full_i)
        np.testing.assert_allclose(d, full_d, rtol=1e-6)


def test_oracle_maxsim_matches_reference_outputs():
    """oracle.maxsim_rerank against the ranked doc ids produced by the reference's own source
    lines (main.py:273-282; tests/golden/make_golden.py::maxsim)."""
    import json
    import os

    with open(os.path.join(os.path.dirname(__file__), "golden", "maxsim_golden.json")) as f:
        gold = json.load(f)
    sc, did = inputs.maxsim_case()
    for g in gold:
        ranked, scores = oracle.maxsim_rerank(sc[g["row"]].tolist(), did[g["row"]].tolist(), g["top_k"])
        assert ranked == g["ranked_docids"]
        assert scores == sorted(scores, reverse=True)


def test_oracle_rank_positive_matches_reference_outputs():
    import json
    import os

    with open(os.path.join(os.path.dirname(__file__), "golden", "rank_golden.json")) as f:
        gold = json.load(f)["ranks"]
    q, d = inputs.rank_case()
    assert oracle.rank_positive(q, d).tolist() == gold


# ---------------------------------------------------------------------------------------
# sentence encoder: the restatement against transformers' own BertModel outputs
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,n,s", [("small", 9, 40), ("minilm", 6, 24)])
def test_sbert_oracle_matches_transformers_outputs(golden, name, n, s):
    from tests.golden import inputs

    cfg = inputs.SBERT_SMALL if name == "small" else oracle.MINILM_L6
    g = golden("sbert_golden.npz")
    w = inputs.sbert_weights(cfg)
    ids, mask = inputs.sbert_tokens(cfg, n, s)
    hidden = oracle.bert_hidden_states(w, cfg, ids, mask)
    np.testing.assert_allclose(hidden[0].numpy(), g[f"{name}_hidden_row0"], atol=2e-5)
    np.testing.assert_allclose(oracle.sbert_encode(w, cfg, ids, mask, normalize=False).numpy(), g[f"{name}_pooled"], atol=1e-5)
    emb = oracle.sbert_encode(w, cfg, ids, mask).numpy()
    np.testing.assert_allclose(emb, g[f"{name}_emb"], atol=1e-6)
    np.testing.assert_allclose(np.linalg.norm(emb, axis=1), 1.0, atol=1e-6)


def test_oracle_search_does_not_depend_on_the_score_block_bound():
    """bench.py times the CPU path with 8 GB score blocks, the tests with 1 GB: the chunking of
    the queries must not change a neighbour; the scores may move in the last bit (ATen picks different GEMM kernels
    and summation orders for different batch heights -- the reference's own `q @ emb.T` does the same)."""
    rng = np.random.default_rng(3)
    emb = torch.from_numpy(rng.standard_normal((3000, 48)).astype(np.float32))
    q = torch.from_numpy(rng.standard_normal((37, 48)).astype(np.float32))
    for metric in ("cosine", "euclidean"):
        built = oracle.bruteforce_build(emb, metric)
        d0, i0 = oracle.bruteforce_search(built, q, 7, metric)
        for cb in (4 * 3000 * 1, 4 * 3000 * 5, 1 << 34):  # one query per block, five, everything at once
            d, i = oracle.bruteforce_search(built, q, 7, metric, chunk_bytes=cb)
            np.testing.assert_array_equal(i, i0)
            np.testing.assert_allclose(d, d0, rtol=2e-6, atol=3e-7)  # a few ulp
