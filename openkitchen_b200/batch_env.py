"""``BatchEnv``: the batch API the Python RL / evolution trainers use ("Python" row).

All state lives on the GPU; every attribute is a ZERO-COPY ``torch`` view (via DLPack) of the buffers the
kernels read and write, on torch's current CUDA stream.  A policy writes ``env.actions`` (or passes action
tensors to ``step``), calls ``env.step()`` and reads ``env.obs`` / ``env.reward`` / ``env.done`` -- nothing
crosses PCIe.  Replaces the reference's per-agent host loop of ``torch::tensor(getCurrentState())`` +
``.item()`` (RLRacers/PPO/PPOAgent.hpp:68-102) and ``Pybind/bindings.cpp``'s single-agent ``BoundEnv``.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _capi
from .dlpack import DeviceBuffer
from .env import Env, ray_fan, track_names


class BatchEnv:
    def __init__(self, tracks, n_agents: int, rays=32, device=0, track_id=None, **cfg):
        """tracks: names from the packed racetrack-database set, CSV paths, or (x, y, w_right, w_left)
        column tuples.  rays: a count (evenly spaced -70..70 deg fan) or explicit angles in degrees."""
        if isinstance(device, torch.device):
            device = device.index or 0
        self.device = torch.device("cuda", int(device))
        self.env = Env(device=int(device), **cfg)
        known = set(track_names())
        for t in ([tracks] if isinstance(tracks, str) else tracks):
            if isinstance(t, str):
                self.env.add_named_track(t) if t in known else self.env.load_track_csv(t)
            else:
                self.env.add_track(t)
        nt = self.env.num_tracks()
        fan = ray_fan(rays) if np.isscalar(rays) else np.asarray(rays, dtype=np.float32)
        if track_id is None:
            track_id = (np.arange(n_agents, dtype=np.int64) * nt // n_agents).astype(np.int32)
        self.env.alloc_agents(n_agents, fan, track_id)
        self.n_agents, self.n_rays = n_agents, len(fan)
        self.track_id = torch.as_tensor(np.asarray(track_id, dtype=np.int32))
        self.points_per_track = [self.env.track_info(t).n_points for t in range(nt)]
        self._views = {}
        for name in _capi.BUFFERS:
            self._views[name] = torch.from_dlpack(DeviceBuffer(self.env, name))
        v = self._views
        # the reference's names for things
        self.obs = v["obs"]                  # f32[N,R]  sensor_hits_[i].norm() / kSensorRange
        self.hits = v["hit_rel"]             # f32[N,R,2] Agent::sensor_hits_
        self.hit_points = v["hit_abs"]       # f32[N,R,2] Ray_::hit_x / hit_y
        self.hit_seg = v["hit_seg"]          # i32[N,R]
        self.hit_t = v["hit_t"]              # f32[N,R]
        self.reward = v["reward"]
        self.fitness = v["fitness"]
        self.done = v["done"]
        self.crashed = v["crashed"]
        self.timed_out = v["timed_out"]
        self.nearest_idx = v["nearest_idx"]
        self.pos_x, self.pos_y, self.rot, self.speed = v["pos_x"], v["pos_y"], v["rot"], v["speed"]
        self.act_throttle, self.act_steer = v["act_throttle"], v["act_steer"]
        self._step_count = 0

    # ---- views -----------------------------------------------------------------------------
    def __getitem__(self, name: str) -> torch.Tensor:
        return self._views[name]

    @property
    def pose(self) -> torch.Tensor:
        """f32[N,4] = (x, y, rot_deg, speed), a fresh stack (the state itself is structure-of-arrays)"""
        return torch.stack([self.pos_x, self.pos_y, self.rot, self.speed], dim=1)

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    # ---- the tick --------------------------------------------------------------------------
    def step(self, throttle: torch.Tensor | None = None, steer: torch.Tensor | None = None, k: int = 1):
        """One tick (k ticks with the same action).  throttle/steer: f32[N] CUDA tensors; None = use
        ``act_throttle`` / ``act_steer`` as they are (what ``Agent::updateAction`` left there)."""
        if (throttle is None) != (steer is None):
            raise ValueError("pass both action tensors or neither")
        dt = ds = None
        if throttle is not None:
            throttle = throttle.to(self.device, torch.float32).contiguous()
            steer = steer.to(self.device, torch.float32).contiguous()
            if throttle.numel() != self.n_agents or steer.numel() != self.n_agents:
                raise ValueError("one action per agent")
            dt, ds = throttle.data_ptr(), steer.data_ptr()
        for _ in range(k):
            self.env.launch_step(dt, ds, self._stream())
        self._step_count += k
        return self.obs, self.reward, self.done

    def step_random(self, k: int = 1, seed: int = 0x0C17C4E2):
        """k ticks driven by the counter-based Philox action stream generated inside the kernel"""
        self.env.launch_steps_random(self._step_count, k, seed, self._stream())
        self._step_count += k
        return self.obs, self.reward, self.done

    def cast_rays(self):
        self.env.cast_rays(self._stream())
        return self.obs

    # ---- resets ----------------------------------------------------------------------------
    def reset(self, agents=None, pt_idx=None, lane_alpha=None, heading_off=None):
        """Environment::resetAgent with explicit draws.  agents: None (all), a bool mask or an index tensor.
        pt_idx None = RaceTrack::kStartingIdx (3).  All arguments may be CUDA tensors (no host sync)."""
        if agents is None:
            idx = torch.arange(self.n_agents, device=self.device, dtype=torch.int64)
        else:
            agents = torch.as_tensor(agents, device=self.device)
            idx = torch.nonzero(agents, as_tuple=False).flatten() if agents.dtype == torch.bool else agents.to(torch.int64)
        n = int(idx.numel())
        if n == 0:
            return
        if pt_idx is None:
            pt = torch.full((n,), 3, device=self.device, dtype=torch.int32)
        else:
            pt = torch.as_tensor(pt_idx, device=self.device).to(torch.int32).contiguous()
        la = None if lane_alpha is None else torch.as_tensor(lane_alpha, device=self.device).to(torch.float32).contiguous()
        ho = None if heading_off is None else torch.as_tensor(heading_off, device=self.device).to(torch.float32).contiguous()
        idx = idx.contiguous()
        self.env.reset_device(idx.data_ptr(), pt.data_ptr(), None if la is None else la.data_ptr(),
                              None if ho is None else ho.data_ptr(), n, self._stream())
        self._keep = (idx, pt, la, ho)  # alive until the kernel has run

    def reset_random(self, agents=None, generator: torch.Generator | None = None, randomize_lane=False,
                     randomize_heading=False):
        """Environment::resetAgent(agent, true, lane, heading) (Environment.cpp:79-122) with torch's RNG standing
        in for raylib's GetRandomValue: idx ~ U{0..P-1}, alpha ~ U{10..90}/100, offset ~ +-(45 + U{0..45})."""
        if agents is None:
            idx = torch.arange(self.n_agents, device=self.device)
        else:
            agents = torch.as_tensor(agents, device=self.device)
            idx = torch.nonzero(agents, as_tuple=False).flatten() if agents.dtype == torch.bool else agents.to(torch.int64)
        n = int(idx.numel())
        if n == 0:
            return
        pts = torch.as_tensor(self.points_per_track, device=self.device)[self.track_id.to(self.device)[idx].long()]
        u = torch.rand(n, device=self.device, generator=generator)
        pt = torch.minimum((u * pts).to(torch.int32), (pts - 1).to(torch.int32))
        la = ho = None
        if randomize_lane:
            la = torch.randint(10, 91, (n,), device=self.device, generator=generator).float() / 100.0
        if randomize_heading:
            mag = 45.0 + torch.randint(0, 46, (n,), device=self.device, generator=generator).float()
            sign = torch.where(torch.arange(n, device=self.device) % 2 == 0, -1.0, 1.0)
            ho = mag * sign
        self.reset(idx, pt, la, ho)

    def close(self):
        """Drops this object's views and closes the env.  Tensors the caller still holds keep the device memory alive:
        the OkEnv is destroyed when the last of them is released (openkitchen_b200.dlpack)."""
        self._views.clear()
        for name in ("obs", "hits", "hit_points", "hit_seg", "hit_t", "reward", "fitness", "done", "crashed", "timed_out",
                     "nearest_idx", "pos_x", "pos_y", "rot", "speed", "act_throttle", "act_steer"):
            setattr(self, name, None)
        self.env.close()
