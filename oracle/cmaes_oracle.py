"""Restatement of the reference's CMA-ES solver and controller (TEST INFRASTRUCTURE ONLY).

Follows CovarianceMatrixAdaptationEvolution/CmaEsSolverTorch.cpp statement by statement, candidate by candidate (the
reference's own loop structure), with torch CPU tensors so that every operation is the ATen kernel the reference
itself calls -- including the details that decide the bits:

* state tensors are float32 (`opts32()`), the constants are C++ `float` members computed in binary64 and narrowed;
* `sample` symmetrises C, takes a float32 `linalg_eigh`, and draws ONE `randn(N)` per candidate from the default
  CPU generator, in candidate order (CmaEsSolverTorch.cpp:73-79);
* `tell` promotes every solution to float64 (`.to(device_, torch::kFloat64)`, :96,:116): the in-place updates
  `param_mean_ += w * x` and `rank_mu_update += w * y y^T` therefore compute in binary64 and round to the float32
  destination after every candidate; `y_i` and the outer products are binary64.

PINNED: tests/test_cmaes_ref_cpu.py runs this class against the reference's own objects (oracle/_ref/libcmaes_ref.so =
CmaEsSolverTorch.cpp + Controller.cpp compiled unchanged against pip-torch's libtorch, oracle/cmaes_ref_harness.cpp) on
identical seeds -- every state tensor identical bit for bit, generation after generation -- and against
tests/golden/cmaes_ref.npz minted from that library by tools/make_golden_cmaes.py."""
import ctypes as C
import math
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(HERE, "_ref", "libcmaes_ref.so")


def _f32(v: float) -> float:
    """a binary64 expression stored into a C++ `float` member"""
    return float(np.float32(v))


class CmaEsOracle:
    def __init__(self, num_params, population_size, sigma=0.5):
        n = self.n = int(num_params)
        self.lam, self.mu = int(population_size), int(population_size) // 2
        self.sigma = _f32(sigma)
        f32 = dict(dtype=torch.float32)
        self.mean = torch.zeros(n, **f32)
        self.C = torch.eye(n, **f32)
        self.p_sigma = torch.zeros(n, **f32)
        self.p_c = torch.zeros(n, **f32)
        # CmaEsSolverTorch.cpp:20-25: each weight is a binary64 expression stored into a float32 element
        w = torch.empty(self.mu, **f32)
        for i in range(self.mu):
            w[i] = math.log(self.mu + 0.5) - math.log(i + 1.0)
        w /= w.sum()
        self.weights = w
        self.mu_eff = _f32(1.0 / w.pow(2).sum().item())  # :30 (1.0 / float -> binary64 -> float member)
        me = self.mu_eff
        # C++ usual arithmetic conversions: `num_params_ + mu_eff_` is int + float = FLOAT, `mu_eff_ / num_params_` is a
        # FLOAT division; everything that touches a double literal first is binary64
        n_plus_me = float(np.float32(n) + np.float32(me))
        me_over_n = float(np.float32(me) / np.float32(n))
        self.c_sigma = _f32((me + 2.0) / (n_plus_me + 5.0))
        self.d_sigma = _f32(1.0 + 2.0 * max(0.0, math.sqrt((me - 1.0) / (n + 1.0)) - 1.0) + self.c_sigma)
        self.c_c = _f32((4.0 + me_over_n) / (n + 4.0 + 2.0 * me / n))
        self.c_1 = _f32(2.0 / ((n + 1.3) * (n + 1.3) + me))
        self.c_mu = _f32(min(1.0 - self.c_1, 2.0 * (me - 2.0 + 1.0 / me) / ((n + 2.0) * (n + 2.0) + me)))
        self.chi_n = _f32(math.sqrt(n) * (1.0 - 1.0 / (4.0 * n) + 1.0 / (21.0 * n * n)))
        self.B = torch.eye(n, **f32)
        self.D = torch.ones(n, **f32)

    def sample(self, z=None):
        """CmaEsSolverTorch.cpp:48-79.  z: optional f32[lambda, N] draws; None = torch.randn per candidate from the
        default generator, the reference's own order."""
        self.C = (self.C + self.C.t()) / 2.0
        evals, evecs = torch.linalg.eigh(self.C)
        self.D = evals.clamp_min(1e-12).sqrt()
        self.B = evecs.contiguous()
        out = []
        for i in range(self.lam):
            zi = torch.randn(self.n, dtype=torch.float32) if z is None else torch.as_tensor(z[i], dtype=torch.float32)
            y = torch.matmul(self.B, self.D * zi)
            out.append((self.mean + self.sigma * y).to(torch.float32))
        return torch.stack(out)

    def tell(self, solutions, fitness):
        """CmaEsSolverTorch.cpp:82-128.  solutions f32[lambda, N], fitness f32[lambda]"""
        solutions = torch.as_tensor(solutions, dtype=torch.float32)
        fit = [float(np.float32(f)) for f in np.asarray(fitness)]
        # std::sort is not stable; with distinct fitness values any sort agrees.  Ties keep the lower index first here.
        order = sorted(range(self.lam), key=lambda i: -fit[i])
        old_mean = self.mean.clone()
        self.mean.zero_()
        for k in range(self.mu):
            weight = float(self.weights[k].item())
            sol = solutions[order[k]].to(torch.float64).contiguous()
            self.mean += weight * sol  # binary64 arithmetic, rounded into the float32 tensor every step
        y_w = (self.mean - old_mean) / self.sigma
        self.p_sigma = (1.0 - self.c_sigma) * self.p_sigma + math.sqrt(self.c_sigma * (2.0 - self.c_sigma) * self.mu_eff) * torch.matmul(
            self.B, torch.matmul(self.B.t(), y_w) / self.D)
        self.p_c = (1.0 - self.c_c) * self.p_c + math.sqrt(self.c_c * (2.0 - self.c_c) * self.mu_eff) * y_w
        rank_mu = torch.zeros(self.n, self.n, dtype=torch.float32)
        for k in range(self.mu):
            weight = float(self.weights[k].item())
            sol = solutions[order[k]].to(torch.float64).contiguous()
            y_i = (sol - old_mean) / self.sigma
            rank_mu += weight * (y_i.view(-1, 1) * y_i.view(1, -1))
        self.C = (1.0 - self.c_1 - self.c_mu) * self.C + self.c_1 * (self.p_c.view(-1, 1) * self.p_c.view(1, -1)) + self.c_mu * rank_mu
        ps_norm = float(np.float32(self.p_sigma.norm().item()))
        # `c_sigma_ / d_sigma_` and `ps_norm / chiN_` are float / float = FLOAT divisions; the rest is binary64 (:127)
        ratio = float(np.float32(self.c_sigma) / np.float32(self.d_sigma))
        rel = float(np.float32(ps_norm) / np.float32(self.chi_n))
        self.sigma = _f32(self.sigma * math.exp(ratio * (rel - 1.0)))

    def state(self):
        return {"mean": self.mean.numpy().copy(), "C": self.C.numpy().copy(), "p_sigma": self.p_sigma.numpy().copy(),
                "p_c": self.p_c.numpy().copy(), "B": self.B.numpy().copy(), "D": self.D.numpy().copy(),
                "weights": self.weights.numpy().copy(),
                "scalars": np.array([self.sigma, self.mu_eff, self.c_sigma, self.d_sigma, self.c_c, self.c_1, self.c_mu, self.chi_n],
                                    dtype=np.float32)}


def controller_shapes(inputs, hidden=16, outputs=1):
    """Controller.cpp:3-8 in parameters() order: fc1.weight, fc1.bias, fc2.weight, fc2.bias, fc3.weight, fc3.bias"""
    return [(hidden, inputs), (hidden,), (hidden // 2, hidden), (hidden // 2,), (outputs, hidden // 2), (outputs,)]


def controller_forward(flat, obs, inputs, hidden=16, outputs=1):
    """Controller::forward for ONE candidate (Controller.cpp:16-23) with torch's own Linear / tanh kernels"""
    flat = torch.as_tensor(np.asarray(flat), dtype=torch.float32)
    x, off = torch.as_tensor(np.asarray(obs), dtype=torch.float32), 0
    shapes = controller_shapes(inputs, hidden, outputs)
    for k in range(0, 6, 2):
        o, i = shapes[k]
        w = flat[off: off + o * i].view(o, i)
        off += o * i
        b = flat[off: off + o]
        off += o
        x = torch.tanh(torch.nn.functional.linear(x, w, b))
    return x.numpy()


# ---- the reference's own objects -------------------------------------------------------------------------------------
def have_ref() -> bool:
    return os.path.exists(REF_LIB)


class CmaEsReference:
    """The reference's CmaEsSolver (compiled from /root/reference, oracle/_ref/libcmaes_ref.so) behind the same
    interface as CmaEsOracle.  Its randn comes from torch's default CPU generator: torch.manual_seed() in this process
    seeds it (the library links the very libtorch this interpreter has loaded)."""

    def __init__(self, num_params, population_size, sigma=0.5):
        self.lib = C.CDLL(REF_LIB)
        self.lib.cmr_create.restype = C.c_void_p
        self.lib.cmr_create.argtypes = [C.c_int, C.c_int, C.c_float]
        self.lib.cmr_destroy.argtypes = [C.c_void_p]
        self.lib.cmr_sample.argtypes = [C.c_void_p, C.c_void_p]
        self.lib.cmr_tell.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        self.lib.cmr_get.argtypes = [C.c_void_p] * 10
        self.n, self.lam, self.mu = int(num_params), int(population_size), int(population_size) // 2
        self.h = self.lib.cmr_create(self.n, self.lam, float(sigma))

    def __del__(self):
        try:
            if self.h:
                self.lib.cmr_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def sample(self):
        out = np.empty((self.lam, self.n), dtype=np.float32)
        self.lib.cmr_sample(self.h, out.ctypes.data_as(C.c_void_p))
        return torch.from_numpy(out)

    def tell(self, solutions, fitness):
        s = np.ascontiguousarray(np.asarray(solutions), dtype=np.float32)
        f = np.ascontiguousarray(np.asarray(fitness), dtype=np.float32)
        self.lib.cmr_tell(self.h, s.ctypes.data_as(C.c_void_p), f.ctypes.data_as(C.c_void_p))

    def state(self):
        n = self.n
        out = {"mean": np.empty(n, np.float32), "C": np.empty((n, n), np.float32), "p_sigma": np.empty(n, np.float32),
               "p_c": np.empty(n, np.float32), "B": np.empty((n, n), np.float32), "D": np.empty(n, np.float32),
               "weights": np.empty(self.mu, np.float32), "scalars": np.empty(8, np.float32)}
        dt = np.zeros(5, np.int32)
        self.lib.cmr_get(self.h, *(out[k].ctypes.data_as(C.c_void_p) for k in ("mean", "C", "p_sigma", "p_c", "B", "D", "weights", "scalars")),
                         dt.ctypes.data_as(C.c_void_p))
        out["dtypes"] = dt
        return out


def reference_controller_forward(flat, obs, inputs, hidden=16, outputs=1):
    """the reference's Controller objects (one set_params + forward per candidate, as CmaEsAgent::updateAction)"""
    lib = C.CDLL(REF_LIB)
    lib.cmr_controller_forward.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    flat = np.ascontiguousarray(flat, dtype=np.float32)
    obs = np.ascontiguousarray(obs, dtype=np.float32)
    batch = flat.shape[0]
    out = np.empty((batch, outputs), dtype=np.float32)
    lib.cmr_controller_forward(inputs, hidden, outputs, flat.ctypes.data_as(C.c_void_p), obs.ctypes.data_as(C.c_void_p), batch,
                               out.ctypes.data_as(C.c_void_p))
    return out
