"""numpy restatement of the reference's CMA-ES solver and controller (TEST INFRASTRUCTURE ONLY).

Follows CovarianceMatrixAdaptationEvolution/CmaEsSolverTorch.cpp line by line, candidate by candidate (the
reference's own loop structure), in float32 like the reference's tensors; Controller.cpp:3-23 for the MLP.
Parity is to floating-point tolerance (the reference uses libtorch kernels whose summation order is unspecified);
no reference run is available here (libtorch C++ is not installed), so this oracle is pinned only by the
published algorithm (arXiv:1604.00772, cited by the reference at CmaEsSolverTorch.cpp:81): "parity unpinned"."""
import math

import numpy as np


class CmaEsOracle:
    def __init__(self, num_params, population_size, sigma=0.5):
        n = self.n = num_params
        self.lam, self.mu = population_size, population_size // 2
        self.sigma = np.float32(sigma)
        self.mean = np.zeros(n, np.float32)
        self.C = np.eye(n, dtype=np.float32)
        self.p_sigma = np.zeros(n, np.float32)
        self.p_c = np.zeros(n, np.float32)
        w = np.array([math.log(self.mu + 0.5) - math.log(i + 1.0) for i in range(self.mu)], dtype=np.float32)
        self.weights = w / w.sum(dtype=np.float32)
        me = self.mu_eff = np.float32(1.0 / np.sum(self.weights ** 2, dtype=np.float32))
        me = float(me)
        self.c_sigma = np.float32((me + 2.0) / (n + me + 5.0))
        self.d_sigma = np.float32(1.0 + 2.0 * max(0.0, math.sqrt((me - 1.0) / (n + 1.0)) - 1.0) + float(self.c_sigma))
        self.c_c = np.float32((4.0 + me / n) / (n + 4.0 + 2.0 * me / n))
        self.c_1 = np.float32(2.0 / ((n + 1.3) * (n + 1.3) + me))
        self.c_mu = np.float32(min(1.0 - float(self.c_1), 2.0 * (me - 2.0 + 1.0 / me) / ((n + 2.0) * (n + 2.0) + me)))
        self.chi_n = np.float32(math.sqrt(n) * (1.0 - 1.0 / (4.0 * n) + 1.0 / (21.0 * n * n)))
        self.B = np.eye(n, dtype=np.float32)
        self.D = np.ones(n, np.float32)

    def sample(self, z):
        """z: f32[lambda, N] standard normal draws (the reference draws them with torch::randn, one row per candidate)"""
        self.C = ((self.C + self.C.T) / np.float32(2.0)).astype(np.float32)
        evals, evecs = np.linalg.eigh(self.C.astype(np.float64))
        self.D = np.sqrt(np.maximum(evals, 1e-12)).astype(np.float32)
        self.B = evecs.astype(np.float32)
        out = []
        for i in range(self.lam):
            y = self.B @ (self.D * z[i])
            out.append((self.mean + self.sigma * y).astype(np.float32))
        return np.stack(out)

    def tell(self, solutions, fitness):
        order = sorted(range(self.lam), key=lambda i: -float(fitness[i]))  # std::sort by fitness, descending
        old_mean = self.mean.copy()
        self.mean = np.zeros(self.n, np.float32)
        for k in range(self.mu):
            self.mean += self.weights[k] * solutions[order[k]]
        y_w = (self.mean - old_mean) / self.sigma
        self.p_sigma = ((1.0 - self.c_sigma) * self.p_sigma + np.float32(math.sqrt(float(self.c_sigma) * (2.0 - float(self.c_sigma)) * float(self.mu_eff))) *
                        (self.B @ ((self.B.T @ y_w) / self.D))).astype(np.float32)
        self.p_c = ((1.0 - self.c_c) * self.p_c + np.float32(math.sqrt(float(self.c_c) * (2.0 - float(self.c_c)) * float(self.mu_eff))) * y_w).astype(np.float32)
        rank_mu = np.zeros((self.n, self.n), np.float32)
        for k in range(self.mu):
            y_i = (solutions[order[k]] - old_mean) / self.sigma
            rank_mu += self.weights[k] * np.outer(y_i, y_i)
        self.C = ((1.0 - self.c_1 - self.c_mu) * self.C + self.c_1 * np.outer(self.p_c, self.p_c) + self.c_mu * rank_mu).astype(np.float32)
        ps_norm = float(np.linalg.norm(self.p_sigma))
        self.sigma = np.float32(self.sigma * math.exp((float(self.c_sigma) / float(self.d_sigma)) * (ps_norm / float(self.chi_n) - 1.0)))


def controller_forward(flat, obs, inputs, hidden=16, outputs=1):
    """Controller::forward for ONE candidate (Controller.cpp:16-23), parameters in parameters() order"""
    shapes = [(hidden, inputs), (hidden,), (hidden // 2, hidden), (hidden // 2,), (outputs, hidden // 2), (outputs,)]
    x, off = obs.astype(np.float32), 0
    for k in range(0, 6, 2):
        o, i = shapes[k]
        w = flat[off: off + o * i].reshape(o, i)
        off += o * i
        b = flat[off: off + o]
        off += o
        x = np.tanh(w @ x + b).astype(np.float32)
    return x
