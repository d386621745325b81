"""CPU: the population-sharded CMA-ES (openkitchen_b200/cmaes.py): single process == 2 ranks over gloo.  Parity with the
reference solver itself is tests/test_cmaes_ref_cpu.py."""
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.multiprocessing as mp  # noqa: E402

from openkitchen_b200.cmaes import CmaEs  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fitness(x):  # a smooth test function (higher is better)
    return -((x - 0.3) ** 2).sum(-1) + 0.1 * np.sin(5 * x).sum(-1)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from openkitchen_b200.cmaes import CmaEs as Solver

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, lam = 12, 10
    s = Solver(n, lam, device="cpu")
    rng = np.random.default_rng(5)
    for _ in range(6):
        # candidates that do not depend on the solver's eigenvectors: eigh of a near-identity C is ill-conditioned,
        # so two runs that differ by one rounding draw different (equally valid) samples from the same z
        x_full = torch.from_numpy((0.3 * rng.standard_normal((lam, n)) + 0.2).astype(np.float32))
        s.sample()
        x = x_full[s.lo:s.hi]
        s.tell(x, torch.from_numpy(_fitness(x.numpy()).astype(np.float32)))
    q.put((rank, s.mean.numpy(), s.C.numpy(), s.sigma))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_solver_equals_single_process():
    n, lam = 12, 10
    ref = CmaEs(n, lam, device="cpu")
    rng = np.random.default_rng(5)
    for _ in range(6):
        x = torch.from_numpy((0.3 * rng.standard_normal((lam, n)) + 0.2).astype(np.float32))
        ref.sample()
        ref.tell(x, torch.from_numpy(_fitness(x.numpy()).astype(np.float32)))
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted((q.get(timeout=180) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, mean, C, sigma in got:
        assert np.allclose(mean, ref.mean.numpy(), rtol=1e-4, atol=1e-5)
        assert np.allclose(C, ref.C.numpy(), rtol=1e-3, atol=1e-5)
        assert sigma == pytest.approx(ref.sigma, rel=1e-4)
