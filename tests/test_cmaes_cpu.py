"""CPU: the batched / sharded CMA-ES (openkitchen_b200/cmaes.py) against the candidate-by-candidate numpy
restatement of the reference solver (oracle/cmaes_oracle.py), and single-process == 2-rank (gloo)."""
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.multiprocessing as mp  # noqa: E402

from openkitchen_b200.cmaes import CmaEs, PopulationController  # noqa: E402
from oracle.cmaes_oracle import CmaEsOracle, controller_forward  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fitness(x):  # a smooth test function (higher is better)
    return -((x - 0.3) ** 2).sum(-1) + 0.1 * np.sin(5 * x).sum(-1)


def test_batched_solver_tracks_the_reference_algorithm():
    n, lam = 24, 20  # kPopulationSize 20 (main_eigen.cpp:12)
    a, b = CmaEs(n, lam, device="cpu"), CmaEsOracle(n, lam)
    assert np.allclose(a.weights.numpy(), b.weights, rtol=1e-6)
    for k in ("c_sigma", "d_sigma", "c_c", "c_1", "c_mu", "chi_n"):
        assert getattr(a, k) == pytest.approx(float(getattr(b, k)), rel=1e-6), k
    rng = np.random.default_rng(0)
    for gen in range(30):
        z = rng.standard_normal((lam, n)).astype(np.float32)
        xa = a.sample(torch.from_numpy(z)).numpy()
        xb = b.sample(z)
        # eigenvectors are defined up to sign/rotation in degenerate subspaces: compare the distribution-defining
        # products instead of B itself
        assert np.allclose((a.B * a.D**2) @ a.B.t(), (b.B * b.D**2) @ b.B.T, rtol=1e-3, atol=1e-4)
        fit = _fitness(xb).astype(np.float32)
        a.tell(torch.from_numpy(xb), torch.from_numpy(fit))
        b.tell(xb, fit)
        assert np.allclose(a.mean.numpy(), b.mean, rtol=1e-4, atol=1e-5), gen
        assert np.allclose(a.C.numpy(), b.C, rtol=1e-3, atol=1e-5), gen
        assert np.allclose(a.p_sigma.numpy(), b.p_sigma, rtol=1e-3, atol=1e-4), gen
        assert a.sigma == pytest.approx(float(b.sigma), rel=1e-4), gen
    assert _fitness(a.mean.numpy()) > _fitness(np.zeros(n)) + 0.5  # it optimises


def test_population_controller_matches_per_candidate_forward():
    rays = 32
    pc = PopulationController(rays)
    assert pc.num_params == 673  # 16R+16+136+9 at R = 32
    rng = np.random.default_rng(1)
    flat = rng.standard_normal((7, pc.num_params)).astype(np.float32) * 0.5
    obs = rng.random((7, rays)).astype(np.float32)
    got = pc.forward(torch.from_numpy(flat), torch.from_numpy(obs)).numpy()
    want = np.stack([controller_forward(flat[i], obs[i], rays) for i in range(7)])
    assert np.allclose(got, want, rtol=1e-5, atol=1e-6)


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
