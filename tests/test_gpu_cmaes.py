"""GPU: one CMA-ES generation end to end on the batched environment (sampling GEMM, population controller on the
DLPack views, rollout, fitness from the kernel's CMA-ES reward mode, batched tell)."""
import numpy as np
import pytest

import openkitchen_b200 as ok

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def test_one_generation_runs_and_fitness_matches_oracle():
    from openkitchen_b200.cmaes import CmaEs, PopulationController
    from oracle.api import Oracle

    rays, pop = 15, 64
    ctrl = PopulationController(rays)
    solver = CmaEs(ctrl.num_params, pop, device="cuda", generator=torch.Generator(device="cuda").manual_seed(0))
    env = ok.BatchEnv(["Monza"], pop, rays=rays, reward_mode=ok.REWARD_CMAES_PROGRESS)
    ora = Oracle("port", movement_mode=0, reward_mode=2)
    ora.add_track(ok.track_columns("Monza"))
    ora.alloc_agents(pop, ok.ray_fan(rays))
    x = solver.sample()
    env.reset()
    thr = torch.full((pop,), 100.0, device="cuda")
    zero = torch.zeros(pop, device="cuda")
    env.step(zero, zero)
    ora.step(np.zeros(pop, np.float32), np.zeros(pop, np.float32))
    for _ in range(80):
        steer = 5.0 * ctrl.forward(x, env.obs)[:, 0]
        env.step(thr, steer)
        ora.step(thr.cpu().numpy(), steer.cpu().numpy())
    torch.cuda.synchronize()
    assert np.array_equal(env.fitness.cpu().numpy(), ora.buffer("fitness"))  # main_eigen.cpp:147-163 semantics
    assert np.array_equal(env.crashed.cpu().numpy(), ora.buffer("crashed"))
    mean0, sigma0 = solver.mean.clone(), solver.sigma
    fitness, order = solver.tell(x, env.fitness)
    assert float(fitness[order[0]]) == float(env.fitness.max()) and order.numel() == pop
    assert not torch.equal(solver.mean, mean0) and solver.sigma != sigma0 and torch.isfinite(solver.C).all()
