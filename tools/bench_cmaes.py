#!/usr/bin/env python
"""N1 at BASELINE configs[4]'s per-GPU shape: lambda = 1,048,576 candidates on one B200, N = 673 parameters (the
reference's 16-8-1 controller at 32 rays).  Times every piece of a CMA-ES generation with CUDA events:

  sample : float32 linalg_eigh of C + the lambda x N x N sampling GEMM (cuBLAS SGEMM, TF32 off)
  rollout: per tick, the hand-written controller kernel (ok_cmaes_controller) + the step kernel
  tell   : global ranking + the binary64 partial sums (weighted mean, rank-mu DGEMM over the parents) + state update

and prints one JSON line with each share of a generation of `--ticks` ticks, the controller kernel's HBM fraction,
and the GEMMs' TFLOP/s.  This is the measurement that decides whether the two GEMMs deserve hand-written kernels
("or justify leaving them on cuBLAS with a measured share of generation time")."""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import openkitchen_b200 as ok  # noqa: E402
from openkitchen_b200.cmaes import CmaEs, PopulationController  # noqa: E402


def timed(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--population", type=int, default=1 << 20)
    ap.add_argument("--rays", type=int, default=32)
    ap.add_argument("--ticks", type=int, default=600, help="episode length a generation is charged with (examples/cmaes_racer.py cap)")
    args = ap.parse_args()
    torch.backends.cuda.matmul.allow_tf32 = False
    lam, rays = args.population, args.rays
    ctrl = PopulationController(rays)
    n = ctrl.num_params
    solver = CmaEs(n, lam, device="cuda", seed=3)
    bench.N_RAYS = rays
    env = ok.BatchEnv(ok.track_names(), lam, rays=rays, reward_mode=ok.REWARD_CMAES_PROGRESS, auto_reset=1)
    env.step_random(3)
    # --- sample, split ---
    solver.C = solver.C + 0.01 * torch.randn(n, n, device="cuda")
    solver.C = solver.C @ solver.C.t() / n + torch.eye(n, device="cuda")
    eigh_ms, _ = timed(lambda: torch.linalg.eigh((solver.C + solver.C.t()) / 2.0), reps=2)
    z = torch.randn(lam, n, device="cuda")
    gemm_ms, _ = timed(lambda: (z * solver.D) @ solver.B.t())
    sample_ms, x = timed(lambda: solver.sample(z), reps=2)
    x = x.contiguous()
    # --- rollout ---
    ctl_ms, _ = timed(lambda: ctrl.act(env, x), reps=20, warm=3)
    bmm_ms, _ = timed(lambda: ctrl.forward(x, env.obs), reps=5, warm=1)
    tick_ms, _ = timed(lambda: env.step(), reps=20, warm=3)
    # --- tell ---
    fit = env.fitness + torch.rand(lam, device="cuda")
    tell_ms, _ = timed(lambda: solver.tell(x, fit), reps=2)
    parents = lam // 2
    xs = x[:parents].double()
    w = torch.rand(parents, device="cuda", dtype=torch.float64)
    dgemm_ms, _ = timed(lambda: (xs * w[:, None]).t() @ xs, reps=2)
    del xs
    gen_ms = sample_ms + tell_ms + args.ticks * (ctl_ms + tick_ms)
    peak = bench.peak_hbm()[0]
    ctl_bytes = lam * (4 * n + 4 * rays + 8)
    line = {
        "what": "one CMA-ES generation on one B200 (config 5's per-GPU shape)", "population": lam, "num_params": n, "rays": rays,
        "ticks_per_generation": args.ticks,
        "ms": {"sample_total": sample_ms, "sample_eigh_f32": eigh_ms, "sample_gemm_sgemm": gemm_ms, "tell_total": tell_ms,
               "tell_rank_mu_dgemm_alone": dgemm_ms, "controller_kernel_per_tick": ctl_ms, "controller_torch_bmm_per_tick": bmm_ms,
               "step_kernel_per_tick": tick_ms, "generation": gen_ms},
        "share_of_generation": {"sample": sample_ms / gen_ms, "tell": tell_ms / gen_ms,
                                "controller": args.ticks * ctl_ms / gen_ms, "step": args.ticks * tick_ms / gen_ms},
        "tflops": {"sample_sgemm_fp32": 2.0 * lam * n * n / (gemm_ms * 1e-3) / 1e12,
                   "rank_mu_dgemm_fp64": 2.0 * parents * n * n / (dgemm_ms * 1e-3) / 1e12},
        "controller_kernel": {"bytes_per_launch": ctl_bytes, "GBps": ctl_bytes / (ctl_ms * 1e-3) / 1e9, "hbm_peak_GBps": peak,
                              "frac": ctl_bytes / (ctl_ms * 1e-3) / 1e9 / peak, "speedup_over_bmm": bmm_ms / ctl_ms},
        "agent_steps_per_sec_generation": lam * args.ticks / (gen_ms * 1e-3),
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    t0 = time.time()
    main()
