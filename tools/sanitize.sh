#!/bin/bash
# compute-sanitizer over the hot path: memcheck and racecheck on smoke() (3 tracks, 96 agents,
# auto-reset, both ray passes + grid-walk fallback) and on a multi-track beam rollout with sensor offset; then
# memcheck / racecheck / synccheck on the small kernels (actor, controller, track query).  Logs -> gpurun_out/sanitize_*.log
# usage (on the GPU box): tools/sanitize.sh [tag]
cd "$(dirname "$0")/.."
tag=${1:-r2}
mkdir -p gpurun_out
cat > /tmp/ok_sanitize_case.py <<'PY'
import sys
import numpy as np
sys.path.insert(0, ".")
import openkitchen_b200 as ok
import __graft_entry__ as g
which = sys.argv[1]
if which == "smoke":
    g.smoke()
elif which == "rollout":
    # staged beam kernel (n >= 2048 agents) on 4 tracks incl. the largest, sensor offset, auto-reset, then the unstaged shape
    for n in (2304, 96):
        env = ok.Env(device=0, reward_mode=ok.REWARD_CMAES_PROGRESS, auto_reset=1, sensor_offset=3.0)
        names = ["Spa", "Monza", "Sepang", "Austin"]
        for nm in names:
            env.add_named_track(nm)
        tid = (np.arange(n) * len(names) // n).astype(np.int32)
        env.alloc_agents(n, ok.ray_fan(32), tid)
        env.launch_steps_random(0, 6)
        env.sync()
        print("rollout ok", n, float(env.read("crashed").mean()))
        env.close()
else:
    import torch
    from openkitchen_b200.cmaes import PopulationController
    from openkitchen_b200.rollout import FusedActorRollout
    env = ok.BatchEnv(["Monza"], 300, rays=[-70, -30, 0, 30, 70], reward_mode=ok.REWARD_CONSTANT, auto_reset=1)
    l1, l2 = torch.nn.Linear(5, 128).cuda(), torch.nn.Linear(128, 3).cuda()
    r = FusedActorRollout(env, l1, l2, torch.tensor([[60.0, 0.0], [30.0, 5.0], [30.0, -5.0]]), steps=4)
    r.run_eager()
    env2 = ok.BatchEnv(["Monza"], 200, rays=32)
    ctrl = PopulationController(32)
    ctrl.act(env2, torch.randn(200, ctrl.num_params, device="cuda"))
    env2.step()
    env2.env.track_query(np.linspace(100, 900, 64, dtype=np.float32), np.linspace(100, 900, 64, dtype=np.float32))
    torch.cuda.synchronize()
    print("small kernels ok")
PY
for tool in memcheck racecheck; do
  for case in smoke rollout small; do
    log=gpurun_out/sanitize_${tag}_${tool}_${case}.log
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python /tmp/ok_sanitize_case.py $case > $log 2>&1
    echo "$tool $case: exit $? -- $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1)"
  done
done
timeout 600 compute-sanitizer --tool synccheck --print-limit 20 python /tmp/ok_sanitize_case.py rollout > gpurun_out/sanitize_${tag}_synccheck_rollout.log 2>&1
echo "synccheck rollout: exit $? -- $(grep -E 'ERROR SUMMARY' gpurun_out/sanitize_${tag}_synccheck_rollout.log | tail -1)"
