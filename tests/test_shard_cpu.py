"""CPU: the global agent id (OkoConfig / OkConfig::agent_id_base) makes a shard reproduce its slice of the unsharded
population -- checked here on the oracle (the GPU twin is tests/test_gpu_parity_gaps.py)."""
import numpy as np

import bench
import openkitchen_b200 as ok
from oracle.api import Oracle
from tests.util import ALL_BUFS, same_bits


def test_oracle_shard_equals_slice_of_unsharded_population():
    n, lo, m, ticks = 23 * 6, 51, 40, 30
    full = Oracle("port", reward_mode=2, auto_reset=1)
    bench.build_workload(ok, full, n, is_oracle=True)
    shard = Oracle("port", reward_mode=2, auto_reset=1, agent_id_base=lo)
    bench.build_workload(ok, shard, m, is_oracle=True, id_base=lo, n_total=n)
    for s in range(ticks):
        for o in (full, shard):
            o.fill_random_actions(s, bench.SEED)
            o.step()
    for name in ALL_BUFS:
        assert same_bits(full.buffer(name)[lo:lo + m], shard.buffer(name)).all(), name
    # and the stream really is keyed by the global id: without the base the shard diverges
    wrong = Oracle("port", reward_mode=2, auto_reset=1)
    bench.build_workload(ok, wrong, m, is_oracle=True, id_base=lo, n_total=n)
    wrong.fill_random_actions(0, bench.SEED)
    shard.fill_random_actions(0, bench.SEED)
    assert not np.array_equal(wrong.buffer("act_throttle"), shard.buffer("act_throttle"))


def test_config_struct_layouts_match():
    """ctypes mirrors of OkConfig / OkoConfig have the C sizes (a mismatch would shift agent_id_base)"""
    import ctypes as C

    from openkitchen_b200 import _capi
    from oracle import api

    assert C.sizeof(_capi.OkConfig) == 72 and _capi.OkConfig.agent_id_base.offset == 64
    assert C.sizeof(api.OkoConfig) == 56 and api.OkoConfig.agent_id_base.offset == 48
    cfg = _capi.OkConfig()
    _capi.load().ok_config_default(C.byref(cfg))
    assert cfg.agent_id_base == 0 and cfg.beam_bins == 256
