"""CPU, world_size 2 over gloo: the multi-GPU host logic -- agent sharding and the per-generation fitness
all-gather / ranking that every rank must agree on."""
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.multiprocessing as mp  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from openkitchen_b200 import dist as okd

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = okd.shard_bounds(n_total, rank, world)
    ids = torch.arange(lo, hi, dtype=torch.float32)
    local = (ids * 37.0) % 11.0  # fitness with many ties
    f, order = okd.global_ranking(local, n_total)
    vals, idx = okd.top_k(local, 5, n_total)
    q.put((rank, lo, hi, f.numpy(), order.numpy(), vals.numpy(), idx.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [10, 11])
def test_fitness_allgather_and_ranking_agree_across_ranks(n_total):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted((q.get(timeout=120) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, f0, o0, v0, i0), (r1, lo1, hi1, f1, o1, v1, i1) = got
    assert (lo0, hi1) == (0, n_total) and hi0 == lo1 and abs((hi0 - lo0) - (hi1 - lo1)) <= 1
    want = (np.arange(n_total, dtype=np.float32) * 37.0) % 11.0
    assert np.array_equal(f0, want) and np.array_equal(f1, want)
    assert np.array_equal(o0, o1) and np.array_equal(i0, i1) and np.array_equal(v0, v1)
    assert np.array_equal(o0, np.argsort(-want, kind="stable"))


def test_shard_bounds_cover_everything():
    from openkitchen_b200.dist import shard_bounds

    for n in (1, 7, 64, 8_388_608):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
