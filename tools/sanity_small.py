"""Every kernel of liblatentknn once, at small sizes (for compute-sanitizer memcheck):
    compute-sanitizer --tool memcheck python tools/sanity_small.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import latent_rag_b200 as lrb  # noqa: E402

rng = np.random.default_rng(0)
t = lambda *s: torch.from_numpy(rng.standard_normal(s).astype(np.float32))

for dim, n, b, k in [(384, 3000, 5, 10), (384, 3000, 300, 10), (768, 2000, 140, 100), (64, 5000, 1, 128),
                     (100, 700, 33, 40), (64, 400000, 40, 100)]:
    for metric in ("cosine", "euclidean"):
        r = lrb.BruteForceRetriever(t(n, dim), [""] * n, None, metric=metric)
        d, i = r.search(t(b, dim), k)
        assert (i >= 0).all() and (np.diff(d, axis=1) <= 0).all(), (dim, n, b, k, metric)
        r.index.close()
r = lrb.BruteForceRetriever(t(2000, 48), [""] * 2000, (np.arange(2000) // 3).tolist(), metric="cosine", precision="fp32")
r.search(t(9, 48), 7)
r.retrieve_batch(t(20, 48), top_k=5, candidate_k=15)
r = lrb.BruteForceRetriever(t(1500, 32), [""] * 1500, None, metric="mahalanobis")
r.search(t(4, 32), 5)
cd = rng.standard_normal((3, 40, 64)).astype(np.float32)
ci = rng.permutation(3 * 40 * 64).reshape(3, 40, 64).astype(np.int64)
lrb.merge_topk(torch.from_numpy(cd).cuda(), torch.from_numpy(ci).cuda(), 50)
lrb.merge_topk(cd[:, :2], ci[:, :2], 10)
gold = os.path.join(ROOT, "tests", "golden")
ae = lrb.load_autoencoder("cae", os.path.join(gold, "ae_weights_cae.npz"), device=0)
ae.set_kernel("umma").encode(t(700, 384).cuda())
ae.set_kernel("simt").encode(t(70, 384).cuda())
comms = [lrb.PeerExchange(0, rk, 3, max_b=64, max_k=16) for rk in range(3)]
for c in comms:
    c.attach_local(comms)
for c in comms:
    c.begin()
    c.publish(t(10, 16).cuda(), torch.arange(160).view(10, 16).cuda() + 1000 * c.rank)
for c in comms:
    c.collect(10, 16)
    c.check()
# top-k above 128: slab search, crowded slab (refinement down to single blocks), deep merge
emb = t(3000, 64)
emb[256:700] = emb[5] + 0.01 * emb[256:700]
r = lrb.BruteForceRetriever(emb, [""] * 3000, None, metric="cosine")
d, i = r.search(emb[5:7], 300)
assert (i >= 0).all() and (np.diff(d, axis=1) <= 0).all()
lrb.merge_topk(torch.from_numpy(cd).cuda(), torch.from_numpy(ci).cuda(), 200)
r.retrieve_batch(emb[:4], top_k=50, candidate_k=150)
# sentence encoder: both attention kernels, both operand precisions
sys.path.insert(0, ROOT)
from tests.golden import inputs  # noqa: E402

cfg = dict(inputs.SBERT_SMALL, max_pos=128)
enc = lrb.SentenceEncoder(inputs.sbert_weights(cfg), heads=cfg["heads"])
for s_len in (20, 100):
    ids, mask = inputs.sbert_tokens(cfg, 5, s_len)
    enc.set_precision("fp32").encode_tokens(ids, mask)
    enc.set_precision("bf16").encode_tokens(ids.cuda(), mask.cuda())
enc.check()
torch.cuda.synchronize()
print("sanity ok")
