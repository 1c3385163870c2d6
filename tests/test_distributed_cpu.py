"""CPU, world_size 2 over gloo: the N>1 host logic -- batch sharding and the single all-reduce of the fitted target
statistics in fit().  The kernels are replaced by a stub `_fit_local` (no GPU here); the collective, the zero
contribution of non-source ranks and the shard arithmetic are the real code."""
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
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from stainlib_b200.normalization.normalizer import ExtractiveStainNormalizer, ReinhardStainNormalizer
    from stainlib_b200.distributed import shard_range, share_fit_statistics

    truth = torch.tensor([0.56, 0.75, 0.33, 0.08, 0.95, 0.27, 1.46, 1.26], dtype=torch.float64)

    class Stub(ExtractiveStainNormalizer):
        def _fit_local(self, target):
            assert dist.get_rank() == 1, "only the source rank may touch its GPU in fit()"
            return truth.clone()

    n = Stub("macenko")
    n.fit(None if rank != 1 else "target-tile", src=1)
    ok = np.allclose(n.stain_matrix_target.reshape(-1), truth[:6].numpy()) and np.allclose(n.maxC_target.reshape(-1), truth[6:].numpy())
    ok = ok and n.stain_matrix_target.shape == (2, 3) and n.maxC_target.shape == (1, 2)
    # the helper alone: non-source values must not leak into the sum
    v = share_fit_statistics(torch.full((6,), float(rank + 5), dtype=torch.float64), src=0)
    ok = ok and bool((v == 5.0).all())
    # shard arithmetic: disjoint cover of the batch
    B = 1001
    lo, hi = shard_range(B, rank, world)
    t = torch.zeros(B, dtype=torch.int64)
    t[lo:hi] = 1
    dist.all_reduce(t)
    ok = ok and bool((t == 1).all())
    q.put((rank, bool(ok), lo, hi))
    dist.destroy_process_group()


def test_fit_allreduce_and_sharding_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), res
    assert (res[0][2], res[0][3], res[1][2], res[1][3]) == (0, 501, 501, 1001)


def test_shard_range_edges():
    from stainlib_b200.distributed import shard_range
    for B in (1, 7, 8, 100000):
        for world in (1, 2, 4, 8):
            spans = [shard_range(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert all(lo <= hi for lo, hi in spans)
