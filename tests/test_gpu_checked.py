"""GPU: index self-checks.  With the product library this only asserts that the checks are compiled out; run through
tools/checked_build.sh (OK_B200_LIB = a -DOK_CHECKED=1 build) the same workload range-checks every table row, chunk,
segment, ray and agent index in the beam kernel and must count zero violations (the pool's compute-sanitizer is closed)."""
import numpy as np
import pytest

import openkitchen_b200 as ok

pytestmark = pytest.mark.gpu


def test_no_index_leaves_its_array():
    names = ["Spa", "Monza", "Sepang", "Austin", "Oschersleben"]
    for n, offset in ((2560, 0.0), (2304, 4.0), (160, -3.0)):  # staged kernel, with sensor offset, unstaged kernel
        env = ok.Env(device=0, reward_mode=ok.REWARD_CMAES_PROGRESS, auto_reset=1, sensor_offset=offset)
        for nm in names:
            env.add_named_track(nm)
        env.alloc_agents(n, ok.ray_fan(32), (np.arange(n) * len(names) // n).astype(np.int32))
        env.launch_steps_random(0, 60)
        # poses far outside the lane and non-finite ones take the uncovered / grid-walk paths
        x = env.read("pos_x")
        x[::7] += 500.0
        x[3::97] = np.nan
        env.write("pos_x", x)
        env.write("crashed", np.zeros(n, dtype=np.uint8))
        env.launch_steps_random(60, 5)
        count, compiled_in = env.debug_violations()
        assert count == 0, f"{count} out-of-range indices"
        env.close()
    if not compiled_in:
        assert ok.LIB_PATH.endswith("libopenkitchen_b200.so")
