"""CMA-ES on the device, population-sharded.

Restates the reference's solver (CovarianceMatrixAdaptationEvolution/CmaEsSolverTorch.cpp:5-131) with the
per-candidate host loops turned into batched tensor ops:

* ``sample``  CmaEsSolverTorch.cpp:48-79  one ``randn [lambda, N]`` and ONE GEMM ``(z * D) @ B^T`` instead of a
  ``matmul`` per candidate;
* ``tell``    CmaEsSolverTorch.cpp:82-128 the weighted mean and the rank-mu update ``sum_i w_i y_i y_i^T`` as
  ``Y^T diag(w) Y`` (one SYRK-shaped GEMM) instead of mu outer products.  Like the reference, which promotes every
  solution to float64 before it touches the sums (CmaEsSolverTorch.cpp:96,116), both sums are computed in BINARY64
  (a DGEMM: B200 keeps full-rate FP64) and narrowed to the float32 state once -- the reference narrows after every
  candidate, so the two agree to float32 rounding (tests/test_cmaes_ref_cpu.py states the tolerance), and the
  eigendecomposition is the reference's float32 ``linalg_eigh`` (CmaEsSolverTorch.cpp:61).

With ``torch.distributed`` initialised every rank holds the candidates it sampled (its slice of the population) and
only these cross NVLink per generation: the fitness all-gather (f32[lambda]) and the all-reduce of the partial weighted
mean (N) and partial rank-mu sum (N x N).  The state update itself is replicated.  These are plain library GEMMs
(torch.matmul); the controller below is the reference's R->16->8->1 tanh MLP evaluated for a whole population with bmm.
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist

from . import dist as okd


def _f32(v) -> float:
    """a value narrowed to binary32, as when the reference stores a binary64 expression into a C++ `float` member"""
    return float(torch.as_tensor(v, dtype=torch.float64).to(torch.float32))


class CmaEs:
    def __init__(self, num_params: int, population_size: int, device="cuda", sigma: float = 0.5,
                 generator: torch.Generator | None = None, seed: int | None = None, chunk: int = 131072):
        """``generator`` / ``seed``: the source of the N(0, I) draws.  Every rank must draw DIFFERENT numbers for its
        slice of the population (the reference draws lambda independent candidates, CmaEsSolverTorch.cpp:73-79): with
        ``seed`` the solver builds its own generator seeded by (seed, rank); a caller-supplied ``generator`` must
        already be rank-distinct."""
        self.n, self.lam, self.mu = num_params, population_size, population_size // 2
        self.device = torch.device(device)
        self.sigma = float(torch.tensor(sigma, dtype=torch.float32))
        self.gen = generator
        self.chunk = int(chunk)
        f32 = dict(dtype=torch.float32, device=self.device)
        self.mean = torch.zeros(num_params, **f32)
        self.C = torch.eye(num_params, **f32)
        self.p_sigma = torch.zeros(num_params, **f32)
        self.p_c = torch.zeros(num_params, **f32)
        i = torch.arange(self.mu, dtype=torch.float64)
        w = (math.log(self.mu + 0.5) - torch.log(i + 1.0)).to(torch.float32)  # CmaEsSolverTorch.cpp:20-25
        w = w / w.sum()
        self.weights = w.to(self.device)
        # learning rates, CmaEsSolverTorch.cpp:30-44 (binary64 arithmetic narrowed to binary32 members)
        f = _f32
        self.mu_eff = f(1.0 / (w.pow(2).sum().item()))
        n, me = num_params, self.mu_eff
        # C++ usual arithmetic conversions in the reference's expressions: `num_params_ + mu_eff_` (int + float) and
        # `mu_eff_ / num_params_` are FLOAT operations, the rest is binary64
        n_plus_me, me_over_n = f(f(n) + me), f(torch.tensor(me, dtype=torch.float32) / torch.tensor(float(n), dtype=torch.float32))
        self.c_sigma = f((me + 2.0) / (n_plus_me + 5.0))
        self.d_sigma = f(1.0 + 2.0 * max(0.0, math.sqrt((me - 1.0) / (n + 1.0)) - 1.0) + self.c_sigma)
        self.c_c = f((4.0 + me_over_n) / (n + 4.0 + 2.0 * me / n))
        self.c_1 = f(2.0 / ((n + 1.3) * (n + 1.3) + me))
        self.c_mu = f(min(1.0 - self.c_1, 2.0 * (me - 2.0 + 1.0 / me) / ((n + 2.0) * (n + 2.0) + me)))
        self.chi_n = f(math.sqrt(n) * (1.0 - 1.0 / (4.0 * n) + 1.0 / (21.0 * n * n)))
        self.B = torch.eye(num_params, **f32)
        self.D = torch.ones(num_params, **f32)
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        self.lo, self.hi = okd.shard_bounds(self.lam, self.rank, self.world)
        if seed is not None:
            if generator is not None:
                raise ValueError("pass a generator or a seed, not both")
            self.gen = torch.Generator(device=self.device).manual_seed(int(seed) * 1_000_003 + self.rank)

    # ---- ask -----------------------------------------------------------------------------------
    def sample(self, z: torch.Tensor | None = None) -> torch.Tensor:
        """This rank's slice of the population, f32[hi-lo, N].  ``z`` (standard normal draws for the slice) may be
        supplied for reproducibility tests; otherwise drawn from ``generator``."""
        self.C = (self.C + self.C.t()) / 2.0
        evals, evecs = torch.linalg.eigh(self.C)
        self.D = evals.clamp_min(1e-12).sqrt()
        self.B = evecs.contiguous()
        if z is None:
            z = torch.randn(self.hi - self.lo, self.n, device=self.device, dtype=torch.float32, generator=self.gen)
        y = (z * self.D) @ self.B.t()  # row i: B @ (D * z_i)
        return self.mean + self.sigma * y

    # ---- tell ----------------------------------------------------------------------------------
    def tell(self, local_solutions: torch.Tensor, local_fitness: torch.Tensor):
        """local_solutions f32[hi-lo, N] as returned by ``sample``; local_fitness f32[hi-lo] (higher is better)."""
        fitness, order = okd.global_ranking(local_fitness.to(self.device, torch.float32), self.lam, descending=True)
        rank_of = torch.empty_like(order)
        rank_of[order] = torch.arange(self.lam, device=self.device)
        my_rank = rank_of[self.lo:self.hi]
        w_full = torch.zeros(self.lam, device=self.device, dtype=torch.float32)
        w_full[: self.mu] = self.weights
        w = w_full[my_rank]  # weight of each local candidate (0 for non-parents)
        old_mean = self.mean.clone()
        x = local_solutions.to(self.device, torch.float32)
        # the parents among this rank's candidates, promoted to binary64 chunk by chunk (CmaEsSolverTorch.cpp:96,116)
        parents = torch.nonzero(w > 0, as_tuple=False).flatten()
        part_mean = torch.zeros(self.n, device=self.device, dtype=torch.float64)
        part_rank_mu = torch.zeros(self.n, self.n, device=self.device, dtype=torch.float64)
        old64, sigma64 = old_mean.double(), float(self.sigma)
        for c0 in range(0, int(parents.numel()), self.chunk):
            sel = parents[c0: c0 + self.chunk]
            xs, ws = x[sel].double(), w[sel].double()
            part_mean += (ws[:, None] * xs).sum(0)
            y = (xs - old64) / sigma64
            part_rank_mu += (y * ws[:, None]).t() @ y  # sum_i w_i y_i y_i^T, a DGEMM
        if self.world > 1:
            buf = torch.cat([part_mean, part_rank_mu.flatten()])
            dist.all_reduce(buf)
            part_mean, part_rank_mu = buf[: self.n], buf[self.n:].view(self.n, self.n)
        self.mean = part_mean.to(torch.float32)
        part_rank_mu = part_rank_mu.to(torch.float32)
        y_w = (self.mean - old_mean) / self.sigma
        # evolution paths, equations 31 and 24 of arXiv:1604.00772 (CmaEsSolverTorch.cpp:100-106)
        self.p_sigma = (1.0 - self.c_sigma) * self.p_sigma + math.sqrt(self.c_sigma * (2.0 - self.c_sigma) * self.mu_eff) * (
            self.B @ ((self.B.t() @ y_w) / self.D))
        self.p_c = (1.0 - self.c_c) * self.p_c + math.sqrt(self.c_c * (2.0 - self.c_c) * self.mu_eff) * y_w
        # covariance, equation 30 (CmaEsSolverTorch.cpp:108-121)
        self.C = (1.0 - self.c_1 - self.c_mu) * self.C + self.c_1 * torch.outer(self.p_c, self.p_c) + self.c_mu * part_rank_mu
        # step size, equation 37 (CmaEsSolverTorch.cpp:123-127)
        ps_norm = float(self.p_sigma.norm())
        f = _f32
        # float / float divisions, then binary64, narrowed into the C++ float `sigma_` (CmaEsSolverTorch.cpp:127)
        ratio = f(torch.tensor(self.c_sigma, dtype=torch.float32) / torch.tensor(self.d_sigma, dtype=torch.float32))
        rel = f(torch.tensor(ps_norm, dtype=torch.float32) / torch.tensor(self.chi_n, dtype=torch.float32))
        self.sigma = f(self.sigma * math.exp(ratio * (rel - 1.0)))
        return fitness, order

    def best_solution(self) -> torch.Tensor:
        return self.mean.clone()


class PopulationController:
    """The reference's Controller (Controller.cpp:3-23: Linear R->H, H->H/2, H/2->1, tanh after each) for a whole
    population at once: candidate i's flat parameter vector (torch.nn.Module.parameters() order: W1, b1, W2, b2, W3, b3)
    is viewed as its own weights and applied with bmm."""

    def __init__(self, inputs: int, hidden: int = 16, outputs: int = 1):
        self.shapes = [(hidden, inputs), (hidden,), (hidden // 2, hidden), (hidden // 2,), (outputs, hidden // 2), (outputs,)]
        self.num_params = sum(int(torch.tensor(s).prod()) for s in self.shapes)

    def act(self, env, flat: torch.Tensor, throttle: float = 100.0, steer_scale: float = 5.0):
        """CmaEsAgent::updateAction for the whole population in ONE hand-written kernel (ok_cmaes_controller): reads
        ``env.obs`` and candidate i's parameters ``flat[i]``, writes the env's action buffers (throttle 100,
        steering 5 * output, main_torch.cpp:68-69).  ``env`` is a BatchEnv; follow with ``env.step()``."""
        if flat.dtype != torch.float32 or not flat.is_contiguous() or flat.shape != (env.n_agents, self.num_params):
            raise ValueError("flat must be a contiguous f32[n_agents, num_params] CUDA tensor")
        env.env.cmaes_controller(flat.data_ptr(), self.num_params, self.shapes[0][0], throttle, steer_scale, env._stream())

    def forward(self, flat: torch.Tensor, obs: torch.Tensor) -> torch.Tensor:
        """flat f32[P, num_params], obs f32[P, inputs] -> f32[P, outputs]"""
        x, off = obs[:, :, None], 0
        for k in range(0, 6, 2):
            o, i = self.shapes[k]
            w = flat[:, off: off + o * i].view(-1, o, i)
            off += o * i
            b = flat[:, off: off + o]
            off += o
            x = torch.tanh(torch.bmm(w, x) + b[:, :, None])
        return x[:, :, 0]
