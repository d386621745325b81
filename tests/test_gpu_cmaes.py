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


def test_controller_kernel_matches_the_reference_controller():
    """ok_cmaes_controller (one warp per candidate, weights streamed once) against Controller::forward evaluated
    candidate by candidate with torch's own Linear / tanh (oracle/cmaes_oracle.py, pinned bit-exact to the reference's
    Controller objects by tests/test_cmaes_ref_cpu.py).  Tolerance: 1e-6 absolute on the tanh output (the kernel sums
    in another order than sgemv and uses CUDA's tanhf), i.e. 5e-6 on the steering command."""
    from openkitchen_b200.cmaes import PopulationController
    from oracle.cmaes_oracle import controller_forward

    for rays in (32, 15, 128, 5):
        pop = 97
        ctrl = PopulationController(rays)
        env = ok.BatchEnv(["Monza", "Spa"], pop, rays=rays)
        env.step_random(7)
        torch.manual_seed(rays)
        flat = (0.4 * torch.randn(pop, ctrl.num_params, device="cuda")).contiguous()
        ctrl.act(env, flat)
        torch.cuda.synchronize()
        obs, fl = env.obs.cpu().numpy(), flat.cpu().numpy()
        want = np.stack([controller_forward(fl[i], obs[i], rays) for i in range(pop)])[:, 0]
        got = env.act_steer.cpu().numpy()
        assert np.allclose(got, 5.0 * want, rtol=0, atol=5e-6), (rays, np.abs(got - 5.0 * want).max())
        assert (env.act_throttle.cpu().numpy() == 100.0).all()
        # and the batched torch form agrees too
        assert np.allclose((5.0 * ctrl.forward(flat, env.obs)[:, 0]).cpu().numpy(), got, rtol=0, atol=5e-6)
        env.close()
