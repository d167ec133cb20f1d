"""world_size-2 (and 3) gloo runs of the row-sharded retriever's host logic on CPU: shard
bounds, global-id offsets, the all-gather layout and the merge, with the oracle injected
as the per-shard searcher (the CUDA searcher is covered by the -m gpu tests)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, metric, k, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from latent_rag_b200.sharded import ShardedRetriever, shard_bounds

        rng = np.random.default_rng(21)
        n, d, b = 301, 24, 9
        emb = torch.from_numpy(rng.standard_normal((n, d)).astype(np.float32))
        q = torch.from_numpy(rng.standard_normal((b, d)).astype(np.float32))
        lo, hi = shard_bounds(n, world)[rank]
        local = emb[lo:hi].contiguous()
        idx = oracle.bruteforce_build(local, metric) if hi > lo else local

        def local_search(queries, kk):
            dd, ii = oracle.bruteforce_search(idx, queries, kk, metric)
            return torch.from_numpy(dd), torch.from_numpy(ii + lo)

        def merge(cd, ci, kk):
            return oracle.merge_topk(cd.numpy(), ci.numpy(), kk)

        r = ShardedRetriever(local, lo, metric, local_search=local_search, merge=merge,
                             texts=[f"t{j}" for j in range(n)], doc_ids=list(range(1000, 1000 + n)))
        assert r.n_total == n and r.world == world
        d_out, i_out = r.search(q, k)
        texts, scores, docids = r.retrieve(q[0], top_k=3)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), d=d_out, i=i_out, docids=np.asarray(docids))
        assert texts == [f"t{j}" for j in i_out[0, :3]]
        assert r.get_stats()["search_calls"] == 2
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,metric,k", [(2, "cosine", 10), (2, "euclidean", 5), (3, "cosine", 120),
                                            (2, "cosine", 200)])  # 200: more than a shard holds, above one fused pass
def test_sharded_search_equals_single_shard(tmp_path, world, metric, k):
    import oracle

    port = _free_port()
    mp.spawn(_worker, args=(world, port, metric, k, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(21)
    emb = torch.from_numpy(rng.standard_normal((301, 24)).astype(np.float32))
    q = torch.from_numpy(rng.standard_normal((9, 24)).astype(np.float32))
    d_ref, i_ref = oracle.bruteforce_search(oracle.bruteforce_build(emb, metric), q, k, metric)
    for rank in range(world):
        got = np.load(os.path.join(str(tmp_path), f"rank{rank}.npz"))
        ok, why = oracle.topk_equivalent(d_ref, i_ref, got["d"], got["i"])
        assert ok, f"rank {rank}: {why}"
        assert got["docids"].tolist() == (got["i"][0, :3] + 1000).tolist()
