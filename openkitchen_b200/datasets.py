"""Dataset writer compatible with the reference's expert-data collectors.

`FieldNavigators/collect_data/collect_data_random.cpp:62-88` (Laser2d mode) writes one text file per sample,
``laser2d_<TRACK>_<n>.txt``: R lines ``x y`` -- the agent's ``sensor_hits_`` through ``std::ostream << float`` (i.e.
``%g``, 6 significant digits) -- and a last line ``throttle steering`` without a trailing newline.  The imitation /
transformer / autoencoder trainers read exactly that.  Here the samples come from the device buffers of a BatchEnv
(``hits`` = f32[N,R,2], actions = f32[N]) after one host copy."""
from __future__ import annotations

import os

import numpy as np


def _g(v: float) -> str:
    return "%g" % float(np.float32(v))  # operator<<(float): default floatfield, precision 6


def format_laser2d(hits: np.ndarray, throttle: float, steering: float) -> str:
    lines = [f"{_g(x)} {_g(y)}" for x, y in np.asarray(hits, dtype=np.float32).reshape(-1, 2)]
    return "\n".join(lines) + "\n" + f"{_g(throttle)} {_g(steering)}"


def write_laser2d(directory: str, track_name: str, hits, throttle, steering, first_index: int = 0, agents=None) -> int:
    """Writes one file per selected agent; returns the next free index (the reference's ``ctr_``)."""
    os.makedirs(directory, exist_ok=True)
    hits = np.asarray(hits, dtype=np.float32)
    throttle = np.asarray(throttle, dtype=np.float32).reshape(-1)
    steering = np.asarray(steering, dtype=np.float32).reshape(-1)
    sel = range(hits.shape[0]) if agents is None else agents
    n = first_index
    for a in sel:
        with open(os.path.join(directory, f"laser2d_{track_name}_{n}.txt"), "w") as f:
            f.write(format_laser2d(hits[a], throttle[a], steering[a]))
        n += 1
    return n


def read_laser2d(path: str):
    """(hits f32[R,2], (throttle, steering)) -- the inverse, as the reference's python loaders parse it"""
    rows = [ln.split() for ln in open(path).read().split("\n") if ln.strip()]
    arr = np.array(rows, dtype=np.float32)
    return arr[:-1], (float(arr[-1, 0]), float(arr[-1, 1]))
