"""Row-sharded search over 2 GPUs (skipped on a one-GPU box), candidates exchanged either by the
fused peer-to-peer kernel over NVLink or by an NCCL all-gather: every rank must return the
global top-k of the oracle, repeatedly (the exchange buffers alternate between calls)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _data(metric):
    rng = np.random.default_rng(77)
    emb = torch.from_numpy(rng.standard_normal((30011, 384)).astype(np.float32)).bfloat16().float()
    q = torch.from_numpy(rng.standard_normal((130, 384)).astype(np.float32)).bfloat16().float()
    return emb, q


def _worker(rank, world, port, metric, k, exchange, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    try:
        import latent_rag_b200 as lrb

        emb, q = _data(metric)
        lo, hi = lrb.shard_bounds(len(emb), world)[rank]
        r = lrb.ShardedRetriever(emb[lo:hi].cuda(), lo, metric, device=rank, exchange=exchange, max_batch=64)
        assert r.n_total == len(emb)
        assert (r._xchg is not None) == (exchange == "p2p")
        d, i = r.search(q.cuda(), k)  # 130 queries: three exchange rounds of at most 64
        for rep in range(3):          # the two buffer slots get reused
            d1, i1 = r.search(q[:5 + rep].cuda(), k)
            np.testing.assert_array_equal(i1, i[:5 + rep])
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), d=d, i=i)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
@pytest.mark.parametrize("metric,k", [("cosine", 10), ("euclidean", 32), ("mahalanobis", 10),
                                      ("euclidean", 100),  # config 4's selector: append buffers + 100 candidates per rank
                                      ("cosine", 300)])  # above 128: slab search per shard, all-gather + sorting merge
def test_sharded_search_two_gpus(tmp_path, metric, k, exchange):
    import oracle

    port = _free_port()
    mp.spawn(_worker, args=(2, port, metric, k, exchange, str(tmp_path)), nprocs=2, join=True)
    emb, q = _data(metric)
    if metric == "mahalanobis":
        p = oracle.mahalanobis_precision(emb)
        lw = oracle.mahalanobis_whitener(p)
        ew = oracle.bf16_round(torch.from_numpy((emb.numpy().astype(np.float64) @ lw).astype(np.float32)))
        qw = oracle.bf16_round(torch.from_numpy((q.numpy().astype(np.float64) @ lw).astype(np.float32)))
        d_ref, i_ref = oracle.bruteforce_search(ew, qw, k, "euclidean")
        scale = oracle.euclidean_scale(ew, qw)
    else:
        d_ref, i_ref = oracle.bruteforce_search(oracle.bruteforce_build(emb, metric), q, k, metric)
        scale = oracle.euclidean_scale(emb, q) if metric == "euclidean" else None
    got = [np.load(os.path.join(str(tmp_path), f"rank{r}.npz")) for r in range(2)]
    np.testing.assert_array_equal(got[0]["i"], got[1]["i"])
    ok, why = oracle.topk_equivalent(d_ref, i_ref, got[0]["d"], got[0]["i"], scale=scale)
    assert ok, why
