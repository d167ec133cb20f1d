"""Parity of the CUDA path (through the retriever classes -> ctypes -> C ABI) with the CPU
oracle and with the committed outputs of the reference.  Run on a B200: pytest -m gpu.

Tolerance (BASELINE.json north_star): identical top-k indices except at score ties within
1e-5 relative; scores within 1e-5 * max(1, |s|) -- for the euclidean / mahalanobis scores,
which are cancelling differences -(q2 + e2 - 2qe), relative to the magnitude of the terms
(q2 + max e2): the reference's own fp32 result for an exact self-match is 1.5e-5, not 0.
bf16 storage is compared with the reference/oracle fed the SAME bf16-rounded values
(SURVEY.md section 8c).
"""
import os

import numpy as np
import pytest
import torch

import oracle
from tests.golden import inputs

pytestmark = pytest.mark.gpu

RTOL = 1e-5


@pytest.fixture(scope="module")
def lrb():
    import latent_rag_b200 as m

    m._native.require_device()
    return m


def _assert_topk(d_ref, i_ref, d, i, rtol=RTOL, l2=None):
    """l2 = (emb, queries) for the L2-type metrics: tolerance relative to q2 + max e2."""
    scale = oracle.euclidean_scale(*l2) if l2 is not None else None
    ok, why = oracle.topk_equivalent(d_ref, i_ref, d, i, rtol=rtol, scale=scale)
    assert ok, why


def _l2(metric, emb, q):
    return (emb, q) if metric != "cosine" else None


def _oracle(emb, q, k, metric):
    return oracle.bruteforce_search(oracle.bruteforce_build(emb, metric), q, k, metric)


def test_native_library_is_loaded(lrb):
    lrb.ExactIndex(8, 8).close()
    with open("/proc/self/maps") as f:
        assert "liblatentknn.so" in f.read()
    assert lrb._native.launch_count() >= 0


# ---------------------------------------------------------------------------------------
# the reference's own test vectors (test/test_retrieval.py:61-83), exact fp32 path
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,nq,dim", inputs.SMALL_CASES)
@pytest.mark.parametrize("metric", ["cosine", "euclidean"])
def test_fp32_path_matches_reference_outputs(lrb, golden, n, nq, dim, metric):
    g = golden("retrieval_small.npz")
    emb = inputs.reference_test_embeddings(n, dim)
    r = lrb.BruteForceRetriever(emb.clone(), [f"doc_{j}" for j in range(n)], list(range(n)), metric=metric,
                                precision="fp32")
    d, i = r.search(emb[:nq], 5)
    assert d.dtype == np.float32 and i.dtype == np.int64 and d.shape == (nq, 5)
    _assert_topk(g[f"{n}_{nq}_{dim}_{metric}_D"], g[f"{n}_{nq}_{dim}_{metric}_I"], d, i, l2=_l2(metric, emb, emb[:nq]))
    np.testing.assert_array_equal(i[:, 0], np.arange(nq))
    texts, scores, ids = r.retrieve(emb[0], top_k=5)  # bruteforce.py:86-92
    assert texts == [f"doc_{j}" for j in ids] and len(scores) == 5
    if metric == "cosine":
        assert ids == g[f"{n}_{nq}_{dim}_retrieve_ids"].tolist()
        np.testing.assert_allclose(scores, g[f"{n}_{nq}_{dim}_retrieve_scores"], rtol=1e-5, atol=1e-6)
    st = r.get_stats()
    assert st["search_calls"] == 2 and len(st["per_query_ms"]) == 2 and st["build_time_s"] > 0


@pytest.mark.parametrize("precision,kernel", [("fp32", "simt"), ("bf16", "simt"), ("bf16", "umma")])
@pytest.mark.parametrize("metric", ["cosine", "euclidean"])
def test_edge_cases(lrb, golden, precision, kernel, metric, monkeypatch):
    """k > N clamps (bruteforce.py:81), 1-D query (:59-60), zero corpus row / zero query."""
    monkeypatch.setenv("LK_FORCE_KERNEL", kernel)
    g = golden("retrieval_edge.npz")
    emb, q = inputs.edge_case_inputs()  # exactly representable in bf16
    r = lrb.BruteForceRetriever(emb, [""] * len(emb), None, metric=metric, precision=precision)
    d, i = r.search(q, 50)
    assert d.shape == (3, 7)
    _assert_topk(g[f"clamp_{metric}_D"], g[f"clamp_{metric}_I"], d, i, l2=_l2(metric, emb, q))
    d1, i1 = r.search(q[1], 3)
    assert d1.shape == (1, 3)
    _assert_topk(g[f"oned_{metric}_D"], g[f"oned_{metric}_I"], d1, i1, l2=_l2(metric, emb, q[1]))
    if metric == "cosine":
        assert np.all(d[2] == 0.0)


# ---------------------------------------------------------------------------------------
# bf16 storage: both kernels against the reference run on the same bf16-rounded values
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("kernel", ["simt", "umma"])
@pytest.mark.parametrize("metric", ["cosine", "euclidean"])
def test_bf16_mid_case_matches_reference_outputs(lrb, golden, kernel, metric, monkeypatch):
    monkeypatch.setenv("LK_FORCE_KERNEL", kernel)
    g = golden("retrieval_mid.npz")
    emb, q = inputs.mid_case_inputs()
    r = lrb.BruteForceRetriever(emb, [""] * len(emb), None, metric=metric)
    d, i = r.search(q, 10)
    _assert_topk(g[f"{metric}_D"], g[f"{metric}_I"], d, i, l2=_l2(metric, emb, q))
    # CUDA-resident inputs take the same path
    d2, i2 = r.search(q.cuda(), 10)
    np.testing.assert_array_equal(i, i2)
    np.testing.assert_array_equal(d, d2)


SHAPES = [
    # n, b, dim, k
    (1, 1, 64, 1),
    (100, 10, 64, 5),
    (127, 3, 32, 10),
    (129, 130, 64, 10),
    (1000, 50, 32, 5),
    (5000, 257, 100, 7),
    (20000, 300, 384, 10),
    (20000, 77, 384, 32),
    (3000, 40, 768, 10),
    (3000, 200, 768, 30),
    (40000, 1, 384, 10),
    (2500, 1500, 64, 10),
    # a query tile that does not fit in shared memory (768-d: 5 K blocks resident, 7 streamed) with MANY query-tile
    # pairs over a small corpus: every CTA pair changes segment (reloads the resident blocks, qfull / qempty parity)
    # several times inside its range
    (3000, 3000, 768, 10),
    (1500, 2500, 640, 100),
    # k > 32: append-buffer selector (lists that never fill, fill once, compact many times)
    (3000, 40, 768, 100),
    (20000, 150, 384, 128),
    (500, 7, 64, 64),
    (300, 3, 32, 100),
    (200000, 8, 64, 100),
    (60000, 2500, 64, 100),
    (150000, 130, 128, 33),
    (400000, 40, 64, 100),   # large enough for threshold seeding from a corpus-prefix search
    (330000, 300, 32, 128),
]


@pytest.mark.parametrize("n,b,dim,k", SHAPES)
@pytest.mark.parametrize("metric", ["cosine", "euclidean"])
def test_umma_kernel_matches_oracle(lrb, n, b, dim, k, metric, monkeypatch):
    """tcgen05 kernel over ragged shapes: partial row blocks, partial query tiles, several
    query tiles per CTA range, K padding (dim 32/100), streamed-Q (dim 768), k in both
    register-list sizes and in the append-buffer selector (k > 32)."""
    monkeypatch.setenv("LK_FORCE_KERNEL", "umma")
    rng = np.random.default_rng(n * 7 + b)
    emb = oracle.bf16_round(torch.from_numpy(rng.standard_normal((n, dim)).astype(np.float32)))
    q = oracle.bf16_round(torch.from_numpy(rng.standard_normal((b, dim)).astype(np.float32)))
    if n >= 8 and b >= 4:
        q[::4] = oracle.bf16_round(emb[(np.arange(0, b, 4) * 13) % n] + 0.05 * q[::4])
    r = lrb.BruteForceRetriever(emb, [""] * n, None, metric=metric)
    d, i = r.search(q, k)
    d_ref, i_ref = _oracle(emb, q, k, metric)
    _assert_topk(d_ref, i_ref, d, i, l2=_l2(metric, emb, q))


@pytest.mark.parametrize("n,b,dim,k", [(315, 1, 64, 10), (20000, 1, 384, 10), (9000, 4, 384, 30), (5003, 3, 100, 128),
                                       (100, 2, 32, 5), (39000, 1, 768, 100)])
@pytest.mark.parametrize("metric", ["cosine", "euclidean"])
def test_single_launch_small_batch_path(lrb, n, b, dim, k, metric, monkeypatch):
    """Up to 4 queries over a small corpus (how the reference's caller drives retrieve(), main.py:270-271)
    take the single-launch path: raw queries in, final top-k out of one kernel.  Same results as
    the oracle, as the general path (bit for bit against the SIMT kernel it shares its arithmetic
    with), and as itself from host and device queries; repeated calls reuse the completion ticket."""
    rng = np.random.default_rng(n * 3 + b)
    emb = oracle.bf16_round(torch.from_numpy(rng.standard_normal((n, dim)).astype(np.float32)))
    q = oracle.bf16_round(emb[rng.integers(0, n, b)] + 0.2 * torch.from_numpy(rng.standard_normal((b, dim)).astype(np.float32)))
    r = lrb.BruteForceRetriever(emb, [""] * n, None, metric=metric)
    kk = min(k, n)
    d_ref, i_ref = _oracle(emb, q, kk, metric)
    launches0 = lrb._native.launch_count()
    d, i = r.search(q, k)
    assert lrb._native.launch_count() - launches0 == 1  # ONE kernel
    _assert_topk(d_ref, i_ref, d, i, l2=_l2(metric, emb, q))
    for _ in range(3):
        d2, i2 = r.search(q.cuda(), k)
        np.testing.assert_array_equal(i2, i)
        np.testing.assert_array_equal(d2, d)
    monkeypatch.setenv("LK_FUSED", "0")
    ds, is_ = r.index.search(q, kk, kernel="simt")
    np.testing.assert_array_equal(is_, i)
    np.testing.assert_array_equal(ds, d)
    if b == 1:
        texts, scores, docids = r.retrieve(q[0], top_k=kk)
        assert docids == i[0].tolist()


@pytest.mark.parametrize("n,b,dim,k", [(20000, 700, 384, 10), (20000, 700, 64, 50), (5000, 130, 768, 10), (300, 1, 64, 128)])
def test_scratch_buffers_are_not_relied_on(lrb, n, b, dim, k, monkeypatch):
    """The tcgen05 path clears neither the padded query rows nor the partial lists before a launch
    (query tiles that span fewer scheduling groups than others leave list slots the kernel itself
    has to mark empty).  LK_DBG=8 poisons all of it with huge scores, valid-looking ids and counts."""
    monkeypatch.setenv("LK_FORCE_KERNEL", "umma")
    monkeypatch.setenv("LK_DBG", "8")
    rng = np.random.default_rng(n + b)
    emb = oracle.bf16_round(torch.from_numpy(rng.standard_normal((n, dim)).astype(np.float32)))
    q = oracle.bf16_round(torch.from_numpy(rng.standard_normal((b, dim)).astype(np.float32)))
    r = lrb.BruteForceRetriever(emb, [""] * n, None, metric="euclidean")
    for _ in range(2):  # the second call reuses the (now dirty) workspaces
        d, i = r.search(q, k)
        d_ref, i_ref = _oracle(emb, q, k, "euclidean")
        _assert_topk(d_ref, i_ref, d, i, l2=(emb, q))


@pytest.mark.parametrize("n,b,dim,k", [(1000, 50, 32, 5), (20000, 33, 384, 10), (3000, 9, 768, 100), (700, 5, 48, 128)])
@pytest.mark.parametrize("precision", ["bf16", "fp32"])
@pytest.mark.parametrize("metric", ["cosine", "euclidean"])
def test_simt_kernel_matches_oracle(lrb, n, b, dim, k, precision, metric, monkeypatch):
    monkeypatch.setenv("LK_FORCE_KERNEL", "simt")
    rng = np.random.default_rng(n + b)
    emb = torch.from_numpy(rng.standard_normal((n, dim)).astype(np.float32))
    q = torch.from_numpy(rng.standard_normal((b, dim)).astype(np.float32))
    if precision == "bf16":
        emb, q = oracle.bf16_round(emb), oracle.bf16_round(q)
    r = lrb.BruteForceRetriever(emb, [""] * n, None, metric=metric, precision=precision)
    d, i = r.search(q, k)
    d_ref, i_ref = _oracle(emb, q, k, metric)
    _assert_topk(d_ref, i_ref, d, i, l2=_l2(metric, emb, q))


FP32_TC_SHAPES = [
    # n, b, dim, k          fp32 storage on the tensor cores: split-bf16 planes, K loop tripled
    (1, 1, 64, 1),
    (1000, 50, 32, 5),      # one K block per plane, K padding 32 -> 64
    (5000, 257, 100, 7),    # K padding 100 -> 128: two K blocks per plane, resident query planes
    (20000, 300, 384, 10),  # six K blocks per plane: five hi-plane blocks resident, the rest streamed with the stages
    (3000, 3000, 384, 10),  # the same with many query-tile pairs per CTA pair: segment changes reload the resident blocks
    (3000, 40, 768, 100),
    (60000, 1500, 64, 100),
    (40000, 1, 384, 10),
    (20000, 70, 256, 200),  # above one selector list: slab search over the planes
]


@pytest.mark.parametrize("n,b,dim,k", FP32_TC_SHAPES)
@pytest.mark.parametrize("metric", ["cosine", "euclidean"])
def test_fp32_storage_on_tensor_cores_matches_oracle(lrb, n, b, dim, k, metric, monkeypatch):
    """precision="fp32" through the tcgen05 kernel: rows and queries carried as two bf16 planes
    (x = hi + lo, products hi*hi + hi*lo + lo*hi, fp32 accumulate) against the reference's fp32
    arithmetic on the RAW fp32 inputs, at the north_star tolerance (1e-5 relative)."""
    monkeypatch.setenv("LK_FORCE_KERNEL", "umma")
    rng = np.random.default_rng(n * 3 + b)
    emb = torch.from_numpy(rng.standard_normal((n, dim)).astype(np.float32))
    q = torch.from_numpy(rng.standard_normal((b, dim)).astype(np.float32))
    if n >= 8 and b >= 4:
        q[::4] = emb[(np.arange(0, b, 4) * 13) % n] + 0.05 * q[::4]
    r = lrb.BruteForceRetriever(emb, [""] * n, None, metric=metric, precision="fp32")
    d, i = r.search(q, k)
    d_ref, i_ref = _oracle(emb, q, k, metric)
    _assert_topk(d_ref, i_ref, d, i, l2=_l2(metric, emb, q))
    # rows added later extend the planes; the exact FMA kernel on the same index agrees
    if n >= 1000:
        more = torch.from_numpy(rng.standard_normal((333, dim)).astype(np.float32))
        r.index.add(more)
        both = torch.cat([emb, more])
        d2, i2 = r.index.search(q[:16], min(k, 128))
        d_ref2, i_ref2 = _oracle(both, q[:16], min(k, 128), metric)
        _assert_topk(d_ref2, i_ref2, d2, i2, l2=_l2(metric, both, q[:16]))
        monkeypatch.setenv("LK_FORCE_KERNEL", "simt")
        d3, i3 = r.index.search(q[:16], min(k, 128))
        _assert_topk(d_ref2, i_ref2, d3, i3, l2=_l2(metric, both, q[:16]))


def test_fp32_storage_picks_the_tensor_cores_for_sizeable_work(lrb):
    """AUTO: config 1 at precision="fp32" (10k x 20k x 384) runs on the tcgen05 kernel -- the golden-vector
    sized problems stay on the exact FMA kernel."""
    rng = np.random.default_rng(9)
    emb = torch.from_numpy(rng.standard_normal((20000, 384)).astype(np.float32))
    q = torch.from_numpy(rng.standard_normal((2000, 384)).astype(np.float32))
    r = lrb.BruteForceRetriever(emb, [""] * 20000, None, metric="cosine", precision="fp32")
    r.index.set_timing(True)
    d, i = r.search(q, 10)
    kernel_ms = r.index.last_timing()[0]
    d_ref, i_ref = _oracle(emb, q, 10, "cosine")
    _assert_topk(d_ref, i_ref, d, i)
    assert kernel_ms < 1.5, kernel_ms  # the fp32 FMA kernel needs ~2.8 ms for this batch


@pytest.mark.parametrize("k", [6, 40, 128])
@pytest.mark.parametrize("kernel", ["simt", "umma"])
def test_ties_resolve_to_the_lowest_index(lrb, kernel, k, monkeypatch):
    """Duplicate rows give exactly equal scores; the engine's order is (score desc, index asc)
    whatever the partitioning.  (torch.topk's tie order is unspecified.)"""
    monkeypatch.setenv("LK_FORCE_KERNEL", kernel)
    rng = np.random.default_rng(2)
    base = oracle.bf16_round(torch.from_numpy(rng.standard_normal((300, 64)).astype(np.float32)))
    emb = torch.cat([base, base, base])  # rows j, j+300, j+600 identical
    r = lrb.BruteForceRetriever(emb, [""] * 900, None, metric="euclidean")
    d, i = r.search(base[:20], k)
    for row in range(20):
        assert i[row, :3].tolist() == [row, row + 300, row + 600]
        assert d[row, 0] == d[row, 1] == d[row, 2]
        for t in range(0, k - k % 3, 3):  # every triple of equal scores is in ascending row order
            assert d[row, t] == d[row, t + 1] == d[row, t + 2]
            assert i[row, t] + 300 == i[row, t + 1] and i[row, t] + 600 == i[row, t + 2]
    d_ref, i_ref = _oracle(emb, base[:20], k, "euclidean")
    _assert_topk(d_ref, i_ref, d, i, l2=(emb, base[:20]))


# ---------------------------------------------------------------------------------------
# k > 128: slab search (the reference takes any k: k = min(k, N), bruteforce.py:81-82)
# ---------------------------------------------------------------------------------------
DEEP_SHAPES = [
    # n, b, dim, k
    (3000, 40, 64, 129),
    (20000, 70, 384, 200),
    (300, 3, 32, 300),       # k = N: every unit is split down to its blocks
    (1000, 5, 64, 1000),
    (700, 2, 48, 5000),      # clamps to N
    (50000, 600, 64, 500),   # more than one chunk of queries
    (100000, 9, 128, 4096),
    (5000, 2, 768, 4096),
]


@pytest.mark.parametrize("n,b,dim,k", DEEP_SHAPES)
@pytest.mark.parametrize("metric", ["cosine", "euclidean"])
def test_deep_k_matches_oracle(lrb, n, b, dim, k, metric):
    rng = np.random.default_rng(n + b + k)
    emb = oracle.bf16_round(torch.from_numpy(rng.standard_normal((n, dim)).astype(np.float32)))
    q = oracle.bf16_round(torch.from_numpy(rng.standard_normal((b, dim)).astype(np.float32)))
    r = lrb.BruteForceRetriever(emb, [""] * n, None, metric=metric)
    d, i = r.search(q, k)
    kk = min(k, n)
    assert d.shape == (b, kk) and i.shape == (b, kk)
    d_ref, i_ref = _oracle(emb, q, k, metric)
    _assert_topk(d_ref, i_ref, d, i, l2=_l2(metric, emb, q))
    assert all(len(set(row.tolist())) == kk for row in i)
    # device-resident queries and results, global ids of a shard
    dd, di = r.index.search(q.cuda(), kk, idx_base=7_000_000_000, device_out=True)
    r.index.check()
    np.testing.assert_array_equal(di.cpu().numpy() - 7_000_000_000, i)
    np.testing.assert_array_equal(dd.cpu().numpy(), d)


@pytest.mark.parametrize("precision,kernel", [("fp32", "simt"), ("bf16", "simt"), ("bf16", "umma")])
def test_deep_k_on_every_kernel_with_dirty_workspaces(lrb, precision, kernel, monkeypatch):
    monkeypatch.setenv("LK_FORCE_KERNEL", kernel)
    monkeypatch.setenv("LK_DBG", "8")
    rng = np.random.default_rng(5)
    n, b, dim, k = 9000, 33, 100, 333
    emb = torch.from_numpy(rng.standard_normal((n, dim)).astype(np.float32))
    q = torch.from_numpy(rng.standard_normal((b, dim)).astype(np.float32))
    if precision == "bf16":
        emb, q = oracle.bf16_round(emb), oracle.bf16_round(q)
    r = lrb.BruteForceRetriever(emb, [""] * n, None, metric="euclidean", precision=precision)
    for _ in range(2):
        d, i = r.search(q, k)
        d_ref, i_ref = _oracle(emb, q, k, "euclidean")
        _assert_topk(d_ref, i_ref, d, i, l2=(emb, q))


@pytest.mark.parametrize("metric", ["cosine", "euclidean"])
def test_deep_k_with_the_best_rows_crowded_into_one_slab(lrb, metric):
    """Chunks of one document sit next to each other in the corpus: here 400 adjacent rows are
    all close to the query, so their slab's 128 candidates cannot cover the top 300 and the slab
    is split and searched again, down to single blocks."""
    rng = np.random.default_rng(77)
    n, dim, k = 20000, 64, 300
    emb = torch.from_numpy(rng.standard_normal((n, dim)).astype(np.float32))
    q = torch.from_numpy(rng.standard_normal((3, dim)).astype(np.float32))
    emb[1000:1400] = q[0] + 0.05 * emb[1000:1400]
    emb[19900:20000] = q[1] + 0.05 * emb[19900:20000]   # the ragged last unit
    emb[5000:5200] = q[1] + 0.05 * emb[5000:5200]
    emb, q = oracle.bf16_round(emb), oracle.bf16_round(q)
    r = lrb.BruteForceRetriever(emb, [""] * n, None, metric=metric)
    d, i = r.search(q, k)
    assert set(i[0].tolist()) <= set(range(1000, 1400))
    assert set(i[1].tolist()) == set(range(19900, 20000)) | set(range(5000, 5200))
    d_ref, i_ref = _oracle(emb, q, k, metric)
    _assert_topk(d_ref, i_ref, d, i, l2=_l2(metric, emb, q))


def test_deep_k_ties_resolve_to_the_lowest_index(lrb):
    """(score desc, index asc) also holds across slabs: 1500 identical rows tie at the top and the
    first 300 of them are the answer; triplicated rows come out as ascending triples."""
    rng = np.random.default_rng(3)
    dup = oracle.bf16_round(torch.from_numpy(rng.standard_normal((1, 64)).astype(np.float32)))
    rest = oracle.bf16_round(torch.from_numpy(rng.standard_normal((500, 64)).astype(np.float32)))
    emb = torch.cat([dup.expand(1500, 64), rest])
    r = lrb.BruteForceRetriever(emb, [""] * 2000, None, metric="cosine")
    d, i = r.search(dup, 300)
    assert i[0].tolist() == list(range(300))
    assert (d[0] == d[0, 0]).all()
    base = oracle.bf16_round(torch.from_numpy(rng.standard_normal((300, 64)).astype(np.float32)))
    emb = torch.cat([base, base, base])
    r = lrb.BruteForceRetriever(emb, [""] * 900, None, metric="euclidean")
    d, i = r.search(base[:20], 600)
    for row in range(20):
        assert i[row, :3].tolist() == [row, row + 300, row + 600]
        for t in range(0, 600, 3):
            assert d[row, t] == d[row, t + 1] == d[row, t + 2]
            assert i[row, t] + 300 == i[row, t + 1] and i[row, t] + 600 == i[row, t + 2]
    d_ref, i_ref = _oracle(emb, base[:20], 600, "euclidean")
    _assert_topk(d_ref, i_ref, d, i, l2=(emb, base[:20]))


def test_deep_k_through_the_other_entry_points(lrb):
    """FAISS mirror (k > N pads with -1 like upstream), ShardedRetriever (one shard), k past the limit."""
    rng = np.random.default_rng(9)
    n, dim = 4000, 64
    emb = oracle.bf16_round(torch.from_numpy(rng.standard_normal((n, dim)).astype(np.float32)))
    q = oracle.bf16_round(torch.from_numpy(rng.standard_normal((6, dim)).astype(np.float32)))
    br = lrb.BruteForceRetriever(emb, [""] * n, None, metric="cosine")
    d0, i0 = br.search(q, 250)
    fr = lrb.FAISSEmbeddingRetriever(dim, index_type="flatip")
    fr.build(emb, [""] * n)
    d1, i1 = fr.search(q, 250)
    _assert_topk(d0, i0, d1, i1)
    sr = lrb.ShardedRetriever(emb, 0, "cosine")
    d2, i2 = sr.search(q, 250)
    np.testing.assert_array_equal(i2, i0)
    np.testing.assert_array_equal(d2, d0)
    texts, scores, ids = br.retrieve(q[0], top_k=250)
    assert ids == i0[0].tolist() and len(texts) == 250
    with pytest.raises(ValueError):
        br.index.search(q, lrb._native.LK_MAX_K + 1)


def test_deep_merge_kernel_matches_oracle(lrb):
    rng = np.random.default_rng(14)
    for b, lists, ln, k in [(3, 8, 512, 300), (2, 4, 4096, 4096), (5, 100, 128, 1000), (1, 2, 100, 150), (40, 8, 200, 129)]:
        cd = rng.standard_normal((b, lists, ln)).astype(np.float32)
        ci = rng.permutation(b * lists * ln).reshape(b, lists, ln).astype(np.int64) + 5_000_000_000
        ci[:, -1, -2:] = -1  # padding
        cd[0, 0, :2] = cd[0, 1, 0]  # ties
        kk = min(k, lists * ln - 2)
        d, i = lrb.merge_topk(cd, ci, kk)
        d_ref, i_ref = oracle.merge_topk(cd, ci, kk)
        np.testing.assert_array_equal(i, i_ref)
        np.testing.assert_array_equal(d, d_ref)
        dg, ig = lrb.merge_topk(torch.from_numpy(cd).cuda(), torch.from_numpy(ci).cuda(), kk)
        np.testing.assert_array_equal(ig.cpu().numpy(), i_ref)
    with pytest.raises(lrb._native.NativeError):  # more candidates per query than the sorting merge holds
        lrb.merge_topk(np.zeros((1, 200, 128), np.float32), np.zeros((1, 200, 128), np.int64), 200)


def test_retrieve_batch_with_deep_candidates(lrb):
    """candidate_k = 3 * top_k (main.py:265) above 128."""
    rng = np.random.default_rng(22)
    n, dim, b, top_k = 8000, 64, 20, 60
    emb = oracle.bf16_round(torch.from_numpy(rng.standard_normal((n, dim)).astype(np.float32)))
    q = oracle.bf16_round(emb[rng.integers(0, n, b)] + 0.3 * torch.from_numpy(rng.standard_normal((b, dim)).astype(np.float32)))
    doc_ids = (rng.permutation(n) // 4).tolist()
    r = lrb.BruteForceRetriever(emb, [""] * n, doc_ids, metric="cosine")
    got_ids, got_sc = r.retrieve_batch(q, top_k=top_k, candidate_k=3 * top_k)
    for row in range(b):
        _, scores_k, docids_k = r.retrieve(q[row], top_k=3 * top_k)
        want, want_sc = oracle.maxsim_rerank(scores_k, docids_k, top_k)
        assert got_ids[row] == want
        np.testing.assert_allclose(got_sc[row], want_sc, rtol=1e-6)


def test_metrics_identical_to_oracle_results(lrb):
    """north_star: identical Recall@k, MRR and nDCG."""
    emb, q = inputs.mid_case_inputs()
    r = lrb.BruteForceRetriever(emb, [""] * len(emb), None, metric="cosine")
    _, i = r.search(q, 10)
    _, i_ref = _oracle(emb, q, 10, "cosine")
    rng = np.random.default_rng(0)
    relevant = [[int(i_ref[j, rng.integers(0, 10)]), int(rng.integers(0, len(emb)))] for j in range(len(q))]
    names = ["Recall@10", "MRR@10", "nDCG@10"]
    got = oracle.evaluate_retrieval([row.tolist() for row in i], relevant, names)
    ref = oracle.evaluate_retrieval([row.tolist() for row in i_ref], relevant, names)
    assert got == ref


# ---------------------------------------------------------------------------------------
# FAISSEmbeddingRetriever (flatip) mirror: the reference's own tests
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,nq,dim", inputs.SMALL_CASES)
def test_faiss_matches_bruteforce(lrb, n, nq, dim):
    """test/test_retrieval.py:61-83: ordered doc_ids of FlatIP == brute force, k=5."""
    emb = inputs.reference_test_embeddings(n, dim)
    texts = [f"doc_{j}" for j in range(n)]
    ids = list(range(n))
    bf = lrb.BruteForceRetriever(emb.clone(), texts, ids, metric="cosine")
    fr = lrb.FAISSEmbeddingRetriever(embedding_dim=dim, index_type="flatip", use_gpu=False)
    fr.build(emb.clone(), texts, ids, train=False)
    for qv in emb[:nq]:
        docs_f, _, ids_f = fr.retrieve(qv, top_k=5)
        docs_b, _, ids_b = bf.retrieve(qv, top_k=5)
        assert ids_f == ids_b
        assert docs_f == [texts[j] for j in ids_f] and docs_b == [texts[j] for j in ids_b]


def test_faiss_index_persistence(lrb, tmp_path):
    """test/test_retrieval.py:86-119."""
    emb = inputs.reference_test_embeddings(200, 48)
    texts = [f"chunk_{j}" for j in range(200)]
    ids = list(range(200))
    path = tmp_path / "test.faiss"
    r1 = lrb.FAISSEmbeddingRetriever(embedding_dim=48, index_path=path, index_type="flatip")
    r1.build(emb, texts, ids, train=False)
    r2 = lrb.FAISSEmbeddingRetriever(embedding_dim=48, index_path=path, index_type="flatip")
    docs1, scores1, ids1 = r1.retrieve(emb[0], top_k=10)
    docs2, scores2, ids2 = r2.retrieve(emb[0], top_k=10)
    assert ids1 == ids2 and docs1 == docs2
    np.testing.assert_allclose(scores1, scores2, rtol=1e-6)
    # an incompatible fingerprint rebuilds from scratch (FAISSEmbeddingRetriever.py:225-250)
    r2.build(emb[:50], texts[:50], ids[:50], ae_type="vae")
    assert r2.index.size == 50 and len(r2._texts) == 50


def test_documented_incompatibilities_fail_loudly(lrb):
    """top-k above LK_MAX_K names the limit (the reference takes any k), and a retriever
    built with keep_source=False has no `.emb` to hand out (the corpus only exists as tiles in HBM)."""
    emb = inputs.reference_test_embeddings(6000, 32)
    r = lrb.BruteForceRetriever(emb, [""] * 6000, None, metric="cosine", keep_source=False)
    d, i = r.search(emb[:3], 4096)  # the deepest supported search
    assert d.shape == (3, 4096) and (np.diff(d, axis=1) <= 0).all()
    with pytest.raises(ValueError, match="LK_MAX_K"):
        r.search(emb[:3], 5000)
    with pytest.raises(AttributeError, match="keep_source"):
        r.emb
    r2 = lrb.BruteForceRetriever(emb, [""] * 6000, None, metric="cosine")
    assert torch.allclose(r2.emb, torch.nn.functional.normalize(emb, dim=1))  # bruteforce.py:49-50


def test_truncated_or_corrupt_index_file_starts_clean(lrb, tmp_path):
    """a truncated image (a crash mid-save) or a header that lies about its payload must
    take the reference's 'corrupted file -> start clean' route (FAISSEmbeddingRetriever.py:70-73),
    never reach the device copy; saves go through a temporary file."""
    emb = inputs.reference_test_embeddings(300, 48)
    texts = [f"chunk_{j}" for j in range(300)]
    path = tmp_path / "idx.faiss"
    r1 = lrb.FAISSEmbeddingRetriever(embedding_dim=48, index_path=path, index_type="flatip")
    r1.build(emb, texts, list(range(300)), train=False)
    assert not list(tmp_path.glob("*.tmp*"))
    blob = path.read_bytes()
    path.write_bytes(blob[: len(blob) // 2])            # truncated
    r2 = lrb.FAISSEmbeddingRetriever(embedding_dim=48, index_path=path, index_type="flatip")
    assert r2.index.size == 0 and r2._texts == []
    hlen = int.from_bytes(blob[8:16], "little")
    lying = blob[:16] + blob[16:16 + hlen].replace(b'"n": 300', b'"n": 900') + blob[16 + hlen:]
    path.write_bytes(lying)                              # header rows vs payload sizes disagree
    r3 = lrb.FAISSEmbeddingRetriever(embedding_dim=48, index_path=path, index_type="flatip")
    assert r3.index.size == 0
    # the raw ABI rejects a short buffer by its length
    ix = lrb.ExactIndex(48, 300, metric="cosine")
    tiles, side = r1.index.export_bytes()
    with pytest.raises(ValueError):
        ix.import_bytes(tiles[:-128], side, 300)
    rc = ix._lib.lk_index_import(ix._h, tiles.ctypes.data, int(tiles.nbytes) - 128, side.ctypes.data, int(side.nbytes),
                                 300, None)
    assert rc == -1 and b"need" in ix._lib.lk_last_error()
    ix.import_bytes(tiles, side, 300)
    d, i = ix.search(emb[:5], 3)
    d1, i1 = r1.index.search(emb[:5], 3)
    np.testing.assert_array_equal(i, i1)
    path.write_bytes(blob)
    r4 = lrb.FAISSEmbeddingRetriever(embedding_dim=48, index_path=path, index_type="flatip")
    assert r4.index.size == 300 and r4.retrieve(emb[7], top_k=1)[2] == [7]


def test_adds_and_growth_on_a_side_stream(lrb):
    """add / reserve / export are ordered on the caller's stream -- growing the index right
    after an add on a non-blocking side stream must not lose the rows of that add."""
    rng = np.random.default_rng(3)
    emb = oracle.bf16_round(torch.from_numpy(rng.standard_normal((60_000, 128)).astype(np.float32)))
    whole = lrb.ExactIndex(128, 60_000, metric="euclidean")
    whole.add(emb)
    d0, i0 = whole.search(emb[:64], 5)
    for rep in range(3):
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            ix = lrb.ExactIndex(128, 1024, metric="euclidean")  # grows several times
            e_dev = emb.cuda()
            s.wait_stream(torch.cuda.default_stream())
            for lo in range(0, 60_000, 7_500):
                ix.add(e_dev[lo:lo + 7_500])
            tiles, side = ix.export_bytes()
            d, i = ix.search(emb[:64], 5)
        np.testing.assert_array_equal(i, i0)
        np.testing.assert_array_equal(d, d0)
        t0, s0 = whole.export_bytes()
        np.testing.assert_array_equal(tiles, t0)
        ix.close()


def test_faiss_semantics(lrb):
    emb = inputs.reference_test_embeddings(3, 16)
    fr = lrb.FAISSEmbeddingRetriever(16, index_type="flatip")
    fr.build(emb, ["a", "b", "c"])  # default doc id -1 (FAISSEmbeddingRetriever.py:296)
    assert fr._doc_ids == [-1, -1, -1]
    d, i = fr.search(emb[:2], 5)  # [upstream] pads instead of clamping
    assert i.shape == (2, 5) and (i[:, 3:] == -1).all() and (d[:, 3:] == -np.finfo(np.float32).max).all()
    fr.build(emb, ["d", "e", "f"])  # every build appends (FAISSEmbeddingRetriever.py:252-257,294-296)
    assert fr.index.size == 6 and fr._texts == ["a", "b", "c", "d", "e", "f"]
    with pytest.raises(ValueError, match="Index type not supported"):
        lrb.FAISSEmbeddingRetriever(16, index_type="lsh")
    with pytest.raises(AssertionError):
        fr.build(emb, ["x"])
    r = lrb.build_retriever(emb, ["a", "b", "c"], [7, 8, 9], {"backend": "faiss", "index_type": "flatip"})
    assert r.retrieve(emb[1], top_k=1)[2] == [8]
    r = lrb.build_retriever(emb, ["a", "b", "c"], [7, 8, 9], {"backend": "bruteforce"})
    assert isinstance(r, lrb.BruteForceRetriever) and r.retrieve(emb[2], top_k=1)[2] == [9]


# ---------------------------------------------------------------------------------------
# Mahalanobis (our definition; the reference has none)
# ---------------------------------------------------------------------------------------
def _aniso(n, d, seed):
    rng = np.random.default_rng(seed)
    a = np.diag(np.linspace(0.2, 2.0, d)) @ np.linalg.qr(rng.standard_normal((d, d)))[0]
    return torch.from_numpy((rng.standard_normal((n, d)) @ a).astype(np.float32))


@pytest.mark.parametrize("kernel", ["simt", "umma"])
def test_mahalanobis_bf16_matches_whitened_oracle(lrb, kernel, monkeypatch):
    """The engine whitens in fp64, rounds to bf16 and runs the L2 path; the oracle does the
    same on the CPU (fp64 whitening -> fp32 -> bf16 -> reference euclidean search)."""
    monkeypatch.setenv("LK_FORCE_KERNEL", kernel)
    emb, q = _aniso(3000, 64, 1), _aniso(40, 64, 2)
    p = oracle.mahalanobis_precision(emb)
    r = lrb.BruteForceRetriever(emb, [""] * 3000, None, metric="mahalanobis", precision_matrix=p)
    d, i = r.search(q, 10)
    lw = oracle.mahalanobis_whitener(p)
    ew = oracle.bf16_round(torch.from_numpy((emb.numpy().astype(np.float64) @ lw).astype(np.float32)))
    qw = oracle.bf16_round(torch.from_numpy((q.numpy().astype(np.float64) @ lw).astype(np.float32)))
    d_ref, i_ref = _oracle(ew, qw, 10, "euclidean")
    _assert_topk(d_ref, i_ref, d, i, l2=(ew, qw))


def test_mahalanobis_fp32_matches_fp64_definition(lrb):
    emb, q = _aniso(2000, 48, 3), _aniso(25, 48, 4)
    r = lrb.BruteForceRetriever(emb, [""] * 2000, None, metric="mahalanobis", precision="fp32")
    d, i = r.search(q, 10)
    d_ref, i_ref = oracle.mahalanobis_search(emb, q, 10)  # fp64, precision estimated from the corpus
    lw = oracle.mahalanobis_whitener(oracle.mahalanobis_precision(emb))
    l2 = (torch.from_numpy(emb.numpy().astype(np.float64) @ lw), torch.from_numpy(q.numpy().astype(np.float64) @ lw))
    _assert_topk(d_ref, i_ref, d, i, l2=l2)  # fp32 storage of whitened vectors vs the fp64 definition


# ---------------------------------------------------------------------------------------
# autoencoder encoders against the shipped checkpoints' reference outputs
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["vae", "dae", "cae"])
def test_ae_encoder_matches_reference_outputs(lrb, golden, kind):
    ae = lrb.load_autoencoder(kind, os.path.join(os.path.dirname(__file__), "golden", f"ae_weights_{kind}.npz"))
    x = inputs.ae_input()
    z = ae.encode(x)
    if isinstance(z, tuple):  # retrieval/embedder.py:44-45
        z = z[0]
    assert z.shape == (48, 64) and z.dtype == torch.float32 and not z.is_cuda
    np.testing.assert_allclose(z.numpy(), golden("ae_golden.npz")[f"{kind}_z"], rtol=1e-4, atol=2e-6)
    zc = ae.encode(x.cuda())
    zc = zc[0] if isinstance(zc, tuple) else zc
    assert zc.is_cuda
    np.testing.assert_array_equal(zc.cpu().numpy(), z.numpy())
    if kind == "cae":  # test/test_models.py:26-36
        assert torch.allclose(z.norm(dim=-1), torch.ones(48), atol=1e-6)


def test_ae_encoder_ragged_batch_and_small_dims(lrb):
    rng = np.random.default_rng(4)
    sd = {"encoder.0.weight": rng.standard_normal((8, 16)).astype(np.float32),
          "encoder.0.bias": rng.standard_normal(8).astype(np.float32),
          "encoder.2.weight": rng.standard_normal((4, 8)).astype(np.float32),
          "encoder.2.bias": rng.standard_normal(4).astype(np.float32)}
    for kind, cls in (("dae", lrb.DenoisingAutoencoder), ("cae", lrb.ContrastiveAutoencoder)):
        ae = cls(16, 4, 8)  # the shapes of test/test_models.py
        ae.load_state_dict(sd)
        x = torch.from_numpy(rng.standard_normal((77, 16)).astype(np.float32))
        z = ae.encode(x)
        ref = oracle.ae_encode(x, oracle.load_encoder_weights(sd, kind), kind)
        np.testing.assert_allclose(z.numpy(), ref.numpy(), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("kind", ["vae", "dae", "cae"])
@pytest.mark.parametrize("m", [256, 1000, 40_000])
def test_ae_tensor_core_encoder_matches_oracle(lrb, kind, m):
    """The tcgen05 encoder (split-bf16 operands, 3 MMAs per product) against the fp32 oracle on
    the shipped checkpoints: error below 3e-5 of the row's scale (fp32 itself is ~1e-6), ragged
    last tile, and bit-identical to itself between host and device inputs."""
    gold = os.path.join(os.path.dirname(__file__), "golden")
    ae = lrb.load_autoencoder(kind, os.path.join(gold, f"ae_weights_{kind}.npz")).set_kernel("umma")
    rng = np.random.default_rng(100 + m)
    x = torch.from_numpy(rng.standard_normal((m, 384)).astype(np.float32))
    x /= x.norm(dim=1, keepdim=True)  # SBERT embeddings are unit rows (retrieval/embedder.py:39)
    take = lambda z: z[0] if isinstance(z, tuple) else z
    z = take(ae.encode(x.cuda())).cpu()
    w = oracle.load_encoder_weights(np.load(os.path.join(gold, f"ae_weights_{kind}.npz")), kind)
    ref = oracle.ae_encode(x, w, kind)
    scale = ref.abs().amax(dim=1, keepdim=True)
    err = ((z - ref).abs() / scale).max().item()
    assert err < 3e-5, err
    zs = take(ae.set_kernel("simt").encode(x.cuda())).cpu()  # the fp32 FMA kernel on the same rows
    np.testing.assert_allclose(zs.numpy(), ref.numpy(), rtol=1e-4, atol=2e-6)
    if m <= 1000:
        zh = take(ae.set_kernel("umma").encode(x))  # host input, staged by the library
        np.testing.assert_array_equal(zh.numpy(), z.numpy())
    if kind == "cae":
        assert torch.allclose(z.norm(dim=-1), torch.ones(m), atol=1e-6)


@pytest.mark.parametrize("kind", ["vae", "dae", "cae"])
def test_ae_bf16_precision_matches_the_reference_fed_bf16_values(lrb, kind):
    """The opt-in bf16 encoder (one MMA per product) against the reference's arithmetic fed
    bf16-rounded inputs, weights and hidden activations: fp32 accumulation on both sides, so
    they agree to 2e-5 of the row scale apart from hidden activations that round the other way
    (a different summation order can move a value across a bf16 rounding boundary: 2^-9 of one
    of 512 terms)."""
    gold = os.path.join(os.path.dirname(__file__), "golden")
    ae = lrb.load_autoencoder(kind, os.path.join(gold, f"ae_weights_{kind}.npz")).set_precision("bf16")
    rng = np.random.default_rng(55)
    x = torch.from_numpy(rng.standard_normal((3000, 384)).astype(np.float32))
    x /= x.norm(dim=1, keepdim=True)
    take = lambda z: z[0] if isinstance(z, tuple) else z
    z = take(ae.encode(x.cuda())).cpu()
    w = oracle.load_encoder_weights(np.load(os.path.join(gold, f"ae_weights_{kind}.npz")), kind)
    ref = oracle.ae_encode(x, w, kind, precision="bf16")
    err = (z - ref).abs() / ref.abs().amax(dim=1, keepdim=True)
    assert err.max().item() < 2e-3 and (err > 2e-5).float().mean().item() < 0.02, (err.max().item(), (err > 2e-5).float().mean().item())
    ref32 = oracle.ae_encode(x, w, kind)
    assert ((z - ref32).abs() / ref32.abs().amax(dim=1, keepdim=True)).max().item() < 2e-2  # bf16-level vs fp32
    z32 = take(ae.set_precision("fp32").encode(x.cuda())).cpu()
    assert ((z32 - ref32).abs() / ref32.abs().amax(dim=1, keepdim=True)).max().item() < 3e-5


@pytest.mark.parametrize("m", [1, 127, 129, 256, 385, 5000, 40_000, 300_001])
def test_ae_pair_kernel_ragged_sizes(lrb, m, monkeypatch):
    """The CTA-pair bf16 encoder (lk_ae_pair.cu): a lone row, odd numbers of 128-row tiles (the
    peer CTA of the last pair has no tile), ragged last tiles, many tiles per cluster (every
    barrier's parity wraps), host and device inputs -- against the reference's arithmetic fed
    bf16-rounded operands, and against the single-CTA kernel computing the same thing."""
    gold = os.path.join(os.path.dirname(__file__), "golden")
    w = oracle.load_encoder_weights(np.load(os.path.join(gold, "ae_weights_cae.npz")), "cae")
    rng = np.random.default_rng(m)
    x = torch.from_numpy(rng.standard_normal((m, 384)).astype(np.float32))
    x /= x.norm(dim=1, keepdim=True)
    ae = lrb.load_autoencoder("cae", os.path.join(gold, "ae_weights_cae.npz")).set_precision("bf16")
    n0 = lrb._native.launch_count()
    z = ae.encode(x.cuda()).cpu()
    assert lrb._native.launch_count() - n0 == -(-m // (1 << 18))  # one launch per step of 262,144 rows: no split pass
    ref = oracle.ae_encode(x, w, "cae", precision="bf16")
    err = (z - ref).abs() / ref.abs().amax(dim=1, keepdim=True)
    assert err.max().item() < 2e-3 and (err > 2e-5).float().mean().item() < 0.02, (err.max().item(),)
    assert torch.allclose(z.norm(dim=-1), torch.ones(m), atol=1e-6)
    if m <= 5000:
        np.testing.assert_array_equal(ae.encode(x).numpy(), z.numpy())  # host rows, staged by the library
        monkeypatch.setenv("LK_AE_PAIR", "0")  # the single-CTA kernel on pre-split planes: same operands
        z1 = ae.encode(x.cuda()).cpu()
        e1 = (z - z1).abs() / ref.abs().amax(dim=1, keepdim=True)
        assert e1.max().item() < 2e-3 and (e1 > 2e-5).float().mean().item() < 0.02


@pytest.mark.parametrize("precision,kernel", [("bf16", "auto"), ("fp32", "auto"), ("fp32", "simt")])
def test_ae_host_rows_go_through_the_chunk_pipeline(lrb, precision, kernel):
    """Host rows in / host latents out (retrieval/embedder.py:24-48 returns CPU fp32) are cut into 32 Ki-row chunks whose
    uploads, kernels and downloads overlap on three streams through two staging buffers each way: four chunks with a
    ragged last one (every buffer reused), pageable and page-locked inputs -- bit-equal to the device-to-device call."""
    gold = os.path.join(os.path.dirname(__file__), "golden")
    rng = np.random.default_rng(17)
    m = 3 * 32768 + 1699
    x = torch.from_numpy(rng.standard_normal((m, 384)).astype(np.float32))
    ae = lrb.load_autoencoder("cae", os.path.join(gold, "ae_weights_cae.npz")).set_precision(precision).set_kernel(kernel)
    z_dev = ae.encode(x.cuda()).cpu()
    z_pageable = ae.encode(x)
    z_pinned = ae.encode(x.pin_memory())
    assert not z_pageable.is_cuda and z_pageable.is_pinned()
    np.testing.assert_array_equal(z_pageable.numpy(), z_dev.numpy())
    np.testing.assert_array_equal(z_pinned.numpy(), z_dev.numpy())
    np.testing.assert_array_equal(ae.encode(x[:40_000]).numpy(), z_dev[:40_000].numpy())  # two chunks, a second call


@pytest.mark.parametrize("d_in,d_hidden,d_latent", [(64, 128, 16), (128, 256, 48), (384, 512, 64), (256, 1024, 33)])
def test_ae_pair_kernel_other_dims(lrb, d_in, d_hidden, d_latent):
    rng = np.random.default_rng(d_in + d_latent)
    sd = {"encoder.0.weight": (rng.standard_normal((d_hidden, d_in)) / np.sqrt(d_in)).astype(np.float32),
          "encoder.0.bias": (0.1 * rng.standard_normal(d_hidden)).astype(np.float32),
          "encoder.2.weight": (rng.standard_normal((d_latent, d_hidden)) / np.sqrt(d_hidden)).astype(np.float32),
          "encoder.2.bias": (0.1 * rng.standard_normal(d_latent)).astype(np.float32)}
    ae = lrb.DenoisingAutoencoder(d_in, d_latent, d_hidden)
    ae.load_state_dict(sd)
    ae.set_precision("bf16")
    x = torch.from_numpy(rng.standard_normal((1000, d_in)).astype(np.float32))
    z = ae.encode(x.cuda()).cpu()
    ref = oracle.ae_encode(x, oracle.load_encoder_weights(sd, "dae"), "dae", precision="bf16")
    err = (z - ref).abs() / ref.abs().amax(dim=1, keepdim=True)
    assert err.max().item() < 3e-3 and (err > 3e-5).float().mean().item() < 0.03, (err.max().item(),)


def test_ae_tensor_core_kernel_is_rejected_for_unsupported_dims(lrb):
    rng = np.random.default_rng(4)
    sd = {"encoder.0.weight": rng.standard_normal((8, 16)).astype(np.float32),
          "encoder.0.bias": rng.standard_normal(8).astype(np.float32),
          "encoder.2.weight": rng.standard_normal((4, 8)).astype(np.float32),
          "encoder.2.bias": rng.standard_normal(4).astype(np.float32)}
    ae = lrb.DenoisingAutoencoder(16, 4, 8)
    ae.load_state_dict(sd)
    ae.set_kernel("umma")
    with pytest.raises(lrb.NativeError):
        ae.encode(torch.zeros(300, 16))


def test_embedding_compressor_contract(lrb):
    """retrieval/embedder.py:24-48 with a stand-in sentence encoder (the SBERT forward is out of
    scope): normalised base embeddings -> fused encoder -> float32 CPU tensor; VAE -> mu."""
    class FakeSbert:
        def encode(self, texts, batch_size=64, convert_to_tensor=True, normalize_embeddings=True):
            g = torch.Generator().manual_seed(len(texts))
            x = torch.randn((len(texts), 384), generator=g)
            return x / x.norm(dim=1, keepdim=True) if normalize_embeddings else x

    gold = os.path.join(os.path.dirname(__file__), "golden")
    texts = [f"t{i}" for i in range(300)]
    base = FakeSbert().encode(texts)
    for kind in ("vae", "cae"):
        ae = lrb.load_autoencoder(kind, os.path.join(gold, f"ae_weights_{kind}.npz"))
        comp = lrb.EmbeddingCompressor(autoencoder=ae, device="cuda:0", model=FakeSbert())
        z = comp.encode_text(texts)
        assert z.shape == (300, 64) and z.dtype == torch.float32 and not z.is_cuda
        w = oracle.load_encoder_weights(np.load(os.path.join(gold, f"ae_weights_{kind}.npz")), kind)
        ref = oracle.ae_encode(base, w, kind)
        assert ((z - ref).abs() / ref.abs().amax(dim=1, keepdim=True)).max().item() < 3e-5
        raw = comp.encode_text(texts, compress=False)
        assert raw.shape == (300, 384) and torch.equal(raw, base)


def test_latent_pipeline_config2_shape(lrb):
    """config 2 in miniature: encode corpus + queries with the shipped CAE, cosine top-10."""
    gold = os.path.join(os.path.dirname(__file__), "golden")
    ae = lrb.load_autoencoder("cae", os.path.join(gold, "ae_weights_cae.npz"))
    rng = np.random.default_rng(6)
    docs = torch.from_numpy(rng.standard_normal((6000, 384)).astype(np.float32))
    docs /= docs.norm(dim=1, keepdim=True)
    qs = docs[:64] + 0.05 * torch.from_numpy(rng.standard_normal((64, 384)).astype(np.float32))
    zd, zq = ae.encode(docs.cuda()), ae.encode(qs.cuda())
    r = lrb.BruteForceRetriever(zd, [""] * 6000, None, metric="cosine")
    d, i = r.search(zq, 10)
    w = oracle.load_encoder_weights(np.load(os.path.join(gold, "ae_weights_cae.npz")), "cae")
    zd_ref = oracle.bf16_round(zd.cpu())  # the engine's own latents, rounded as it stores them
    d_ref, i_ref = _oracle(zd_ref, oracle.bf16_round(zq.cpu()), 10, "cosine")
    _assert_topk(d_ref, i_ref, d, i)
    z_ref = oracle.ae_encode(docs, w, "cae")  # 6000 rows -> the tensor-core encoder (split-bf16)
    assert ((zd.cpu() - z_ref).abs() / z_ref.abs().amax(dim=1, keepdim=True)).max().item() < 3e-5
    assert (i[:, 0] == np.arange(64)).mean() > 0.9


# ---------------------------------------------------------------------------------------
# batched retrieve + document-level MaxSim (the reference caller's loop, main.py:264-282)
# ---------------------------------------------------------------------------------------
def test_maxsim_kernel_matches_reference_outputs(lrb):
    """lk_maxsim_rerank on the seeded candidates whose ranked doc ids the reference's own source
    lines produced (tests/golden/maxsim_golden.json)."""
    import json
    from ctypes import c_void_p

    nat = lrb._native
    with open(os.path.join(os.path.dirname(__file__), "golden", "maxsim_golden.json")) as f:
        gold = json.load(f)
    sc, did = inputs.maxsim_case()
    q, ck = sc.shape
    # the kernel maps row ids -> doc ids: use one distinct row per candidate
    rows = torch.arange(q * ck, dtype=torch.int64).view(q, ck).cuda()
    row_doc = torch.from_numpy(did.reshape(-1)).cuda()
    for top_k in (10, 5):
        out_s = torch.empty((q, top_k), dtype=torch.float32, device="cuda")
        out_d = torch.empty((q, top_k), dtype=torch.int64, device="cuda")
        nat.check(nat.load().lk_maxsim_rerank(0, c_void_p(torch.from_numpy(sc).cuda().data_ptr()), c_void_p(rows.data_ptr()),
                                              q, ck, c_void_p(row_doc.data_ptr()), row_doc.numel(), top_k,
                                              c_void_p(out_s.data_ptr()), c_void_p(out_d.data_ptr()), None), "lk_maxsim_rerank")
        got = out_d.cpu().numpy()
        for g in (g for g in gold if g["top_k"] == top_k):
            want = g["ranked_docids"]
            assert got[g["row"], : len(want)].tolist() == want
            assert (got[g["row"], len(want):] == -1).all()


@pytest.mark.parametrize("cls", ["brute", "faiss"])
def test_retrieve_batch_equals_the_per_query_loop(lrb, cls):
    """retrieve_batch == the reference caller's loop: retrieve(q, candidate_k) per query, MaxSim
    per doc id, stable sort, truncate (main.py:264-282), chunked corpus (4 chunks per document)."""
    rng = np.random.default_rng(21)
    n, dim, b, top_k = 6000, 64, 150, 10
    emb = oracle.bf16_round(torch.from_numpy(rng.standard_normal((n, dim)).astype(np.float32)))
    q = oracle.bf16_round(emb[rng.integers(0, n, b)] + 0.3 * torch.from_numpy(rng.standard_normal((b, dim)).astype(np.float32)))
    doc_ids = (rng.permutation(n) // 4).tolist()
    if cls == "brute":
        r = lrb.BruteForceRetriever(emb, [""] * n, doc_ids, metric="cosine")
    else:
        r = lrb.FAISSEmbeddingRetriever(dim, index_type="flatip")
        r.build(emb, [""] * n, doc_ids)
    got_ids, got_sc = r.retrieve_batch(q, top_k=top_k, candidate_k=3 * top_k)
    for row in range(b):
        _, scores_k, docids_k = r.retrieve(q[row], top_k=3 * top_k)
        want, want_sc = oracle.maxsim_rerank(scores_k, docids_k, top_k)
        assert got_ids[row] == want
        np.testing.assert_allclose(got_sc[row], want_sc, rtol=1e-6)


# ---------------------------------------------------------------------------------------
# retrieval metrics on the device (evaluation/retrieval_metrics.py:14-96)
# ---------------------------------------------------------------------------------------
def test_device_metrics_equal_the_reference_bit_for_bit(lrb, golden):
    import json

    with open(os.path.join(os.path.dirname(__file__), "golden", "metrics_golden.json")) as f:
        g = json.load(f)
    # the KATs of test/test_evaluation.py:9-22, through the single-query form
    one = lrb.evaluate_retrieval([1, 2, 3, 4, 5], [3, 4, 6], ["recall@3", "mrr", "ndcg@3"])
    assert one["recall@3"] == 1 / 3 == g["kat"]["recall_at_3"]
    assert one["mrr"] == 1 / 3 == g["kat"]["mrr"]
    assert one["ndcg@3"] == g["kat"]["ndcg_at_3"]
    # the seeded batch whose summary the reference produced (tests/golden/make_golden.py::metrics)
    retrieved, relevant = inputs.metrics_case()
    names = list(g["evaluate_retrieval"])
    res, per_query = lrb.evaluate_retrieval(retrieved, relevant, names, return_per_query=True)
    for name, ref in g["evaluate_retrieval"].items():
        assert res[name]["mean"] == ref["mean"] and res[name]["std"] == ref["std"], name
    # ragged lists, duplicates, empty relevant sets, string ids: against the oracle, per query
    rng = np.random.default_rng(8)
    ret = [[f"d{v}" for v in rng.integers(0, 30, int(rng.integers(1, 40)))] for _ in range(300)]
    rel = [[f"d{v}" for v in rng.integers(0, 30, int(rng.integers(0, 5)))] for _ in range(300)]
    names = ["Recall@10", "MRR@10", "nDCG@10", "recall@3", "mrr", "ndcg@50"]
    res, per_query = lrb.evaluate_retrieval(ret, rel, names, return_per_query=True)
    want = oracle.evaluate_retrieval(ret, rel, names)
    for n in names:
        assert abs(res[n]["mean"] - want[n]["mean"]) < 1e-15 and abs(res[n]["std"] - want[n]["std"]) < 1e-15
    for r, l, pq in zip(ret, rel, per_query):
        assert pq["Recall@10"] == oracle.recall_at_k(r, l, 10)
        assert pq["MRR@10"] == oracle.mrr(r[:10], l)
        assert abs(pq["nDCG@10"] - oracle.ndcg_at_k(r, l, 10)) < 1e-15
    with pytest.raises(ValueError):
        lrb.evaluate_retrieval(ret, rel, ["precision@5"])
    # fewer retrieved ids than the cut-off, more relevant ones than retrieved -- the ideal DCG
    # runs over min(len(relevant), k) with the caller's k (retrieval_metrics.py:29), not the retrieved length
    # (per-query values of the reference itself: tests/golden/make_golden.py::metrics, "ragged_per_query")
    rr, rl = inputs.metrics_ragged_case()
    got, pq = lrb.evaluate_retrieval(rr, rl, ["ndcg@10", "ndcg@3", "recall@10", "mrr"], return_per_query=True)
    for t, one in enumerate(pq):
        for name in ("ndcg@10", "ndcg@3", "recall@10", "mrr"):
            assert one[name] == g["ragged_per_query"][name][t], (t, name)
    assert pq[0]["ndcg@10"] < (1 / np.log2(3) + 1 / np.log2(6)) / sum(1 / np.log2(j + 2) for j in range(5))


def test_rank_positive_matches_reference_outputs(lrb):
    import json

    with open(os.path.join(os.path.dirname(__file__), "golden", "rank_golden.json")) as f:
        gold = json.load(f)["ranks"]
    q, d = inputs.rank_case()
    assert lrb.rank_positive(q, d).tolist() == gold
    assert lrb.rank_positive(q.cuda(), d.cuda()).is_cuda
    rng = np.random.default_rng(3)
    q2 = torch.from_numpy(rng.standard_normal((3000, 384)).astype(np.float32))
    d2 = q2 + 2.0 * torch.from_numpy(rng.standard_normal((3000, 384)).astype(np.float32))
    got, want = lrb.rank_positive(q2, d2), oracle.rank_positive(q2, d2)
    assert (got - want).abs().max().item() <= 1 and (got != want).float().mean().item() < 0.01  # fp32 near-ties


# ---------------------------------------------------------------------------------------
# merge kernel + sharding
# ---------------------------------------------------------------------------------------
def test_merge_kernel_matches_oracle(lrb):
    rng = np.random.default_rng(12)
    # the last cases go through the selection merge (few queries, >= 2048 candidates each)
    for b, lists, ln, k in [(1, 8, 10, 10), (33, 3, 100, 100), (257, 8, 10, 7), (5, 2, 128, 128),
                            (1, 296, 10, 10), (3, 296, 256, 100), (64, 40, 64, 33), (2, 300, 256, 128)]:
        cd = rng.standard_normal((b, lists, ln)).astype(np.float32)
        ci = rng.permutation(b * lists * ln).reshape(b, lists, ln).astype(np.int64)
        ci[:, -1, -2:] = -1  # padding
        cd[0, 0, :2] = cd[0, 1, 0]  # ties
        d, i = lrb.merge_topk(cd, ci, k)
        d_ref, i_ref = oracle.merge_topk(cd, ci, k)
        np.testing.assert_array_equal(i, i_ref)
        np.testing.assert_array_equal(d, d_ref)
        dg, ig = lrb.merge_topk(torch.from_numpy(cd).cuda(), torch.from_numpy(ci).cuda(), k)
        np.testing.assert_array_equal(ig.cpu().numpy(), i_ref)


def test_merge_selection_with_massive_ties(lrb):
    """More tied candidates at the k-th score than the selection merge gathers: lowest ids win."""
    rng = np.random.default_rng(13)
    cd = rng.standard_normal((2, 30, 100)).astype(np.float32)
    ci = rng.permutation(2 * 30 * 100).reshape(2, 30, 100).astype(np.int64)
    cd[0, :, 10:40] = 2.5   # 900 candidates share the score around rank 13..912
    cd[1, :, :] = -3.0      # every candidate ties
    d, i = lrb.merge_topk(cd, ci, 100)
    d_ref, i_ref = oracle.merge_topk(cd, ci, 100)
    np.testing.assert_array_equal(i, i_ref)
    np.testing.assert_array_equal(d, d_ref)


@pytest.mark.parametrize("world,b,k", [(2, 1, 10), (8, 70, 10), (3, 300, 100), (16, 5, 128)])
def test_peer_exchange_protocol_on_one_gpu(lrb, world, b, k):
    """The peer-to-peer candidate exchange with every rank driven from this process on one GPU:
    all ranks publish, then all ranks collect (the fused kernel does both in one launch per rank
    and needs one GPU per rank: tests/test_gpu_multi.py).  Repeated so that both buffer slots
    and growing epochs are used."""
    rng = np.random.default_rng(world * 100 + b)
    comms = [lrb.PeerExchange(0, r, world, max_b=512, max_k=128) for r in range(world)]
    for c in comms:
        c.attach_local(comms)
    for rep in range(3):
        cd = rng.standard_normal((b, world, k)).astype(np.float32)
        cd = -np.sort(-cd, axis=2)
        ci = rng.permutation(b * world * k).reshape(b, world, k).astype(np.int64)
        ci[:, -1, -3:] = -1  # a short shard pads with -1 / -inf
        cd[:, -1, -3:] = -np.inf
        cd[0, 0, 0] = cd[0, 1 % world, 0]  # a tie across ranks
        if rep == 2:  # massive ties within and across the lists
            cd = np.where(np.isfinite(cd), np.round(cd, 1), cd).astype(np.float32)
        # every rank's list sorted by (score desc, id asc) like a search result: the merge ranks by binary searches
        order = np.lexsort((np.where(ci < 0, np.iinfo(np.int64).max, ci), -cd), axis=2)
        cd, ci = np.take_along_axis(cd, order, 2), np.take_along_axis(ci, order, 2)
        for r, c in enumerate(comms):
            c.begin()
            c.publish(torch.from_numpy(cd[:, r]).cuda(), torch.from_numpy(ci[:, r]).cuda())
        d_ref, i_ref = oracle.merge_topk(cd, ci, k)
        for c in comms:
            d, i = c.collect(b, k)
            c.check()
            np.testing.assert_array_equal(i.cpu().numpy(), i_ref)
            np.testing.assert_array_equal(d.cpu().numpy(), d_ref)
    for c in comms:
        c.close()


@pytest.mark.parametrize("world", [1, 2, 8])
def test_row_sharding_is_invisible(lrb, world):
    """Searching `world` row shards (global ids via idx_base) and merging equals the unsharded
    search bit for bit -- the multi-GPU data path run on one GPU."""
    rng = np.random.default_rng(31)
    emb = oracle.bf16_round(torch.from_numpy(rng.standard_normal((5003, 384)).astype(np.float32)))
    q = oracle.bf16_round(torch.from_numpy(rng.standard_normal((70, 384)).astype(np.float32)))
    whole = lrb.BruteForceRetriever(emb, [""] * 5003, None, metric="euclidean")
    d0, i0 = whole.search(q, 10)
    cd, ci = [], []
    for lo, hi in lrb.shard_bounds(5003, world):
        ix = lrb.ExactIndex(384, hi - lo, metric="euclidean")
        ix.add(emb[lo:hi])
        d, i = ix.search(q, 10, idx_base=lo, device_out=True)
        cd.append(d)
        ci.append(i)
    d, i = lrb.merge_topk(torch.stack(cd, 1), torch.stack(ci, 1), 10)
    np.testing.assert_array_equal(i.cpu().numpy(), i0)
    np.testing.assert_array_equal(d.cpu().numpy(), d0)
    sr = lrb.ShardedRetriever(emb, 0, "euclidean")  # world 1, native merge
    ds, is_ = sr.search(q, 10)
    np.testing.assert_array_equal(is_, i0)


# ---------------------------------------------------------------------------------------
# full-size properties (no oracle at this scale)
# ---------------------------------------------------------------------------------------
def test_large_corpus_properties(lrb):
    """2M x 384 bf16 generated on the device: planted neighbours are found, scores sorted,
    indices unique and in range, both kernels agree, batch-1 equals its row of the batch."""
    n, dim, b = 2_000_000, 384, 256
    g = torch.Generator(device="cuda").manual_seed(1234)
    ix = lrb.ExactIndex(dim, n, metric="cosine")
    planted = {}
    for lo in range(0, n, 250_000):
        chunk = torch.randn((250_000, dim), generator=g, device="cuda", dtype=torch.float32).to(torch.bfloat16)
        for j in range(lo, lo + 250_000, 31_250):
            planted[j] = chunk[j - lo].float().cpu()
        ix.add(chunk)
    assert ix.size == n
    rows = sorted(planted)[:b // 4]
    q = torch.randn((b, dim), generator=torch.Generator().manual_seed(4321))
    for t, row in enumerate(rows):
        q[t] = planted[row] + 0.1 * q[t]
    q = oracle.bf16_round(q)
    d, i = ix.search(q, 10)
    assert (i[: len(rows), 0] == np.asarray(rows)).all()
    assert (np.diff(d, axis=1) <= 0).all() and (i >= 0).all() and (i < n).all()
    assert all(len(set(row.tolist())) == 10 for row in i)
    d1, i1 = ix.search(q[3], 10)
    np.testing.assert_array_equal(i1[0], i[3])
    np.testing.assert_array_equal(d1[0], d[3])
    ds, is_ = ix.search(q[:8], 10, kernel="simt")
    _assert_topk(ds, is_, d[:8], i[:8])
