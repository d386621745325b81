"""CPU: the beam table of OK_RAYCAST_BEAM (openkitchen_b200/csrc/ok_beam.cpp) is a conservative candidate list.

For random rays that start inside the lane, the segment the reference's loop selects (CollisionChecker.cu:49-66,
restated here in numpy float64) must be in the ray's list whenever its hit parameter lies within the list's
completeness distance -- that is the only property the kernel relies on; everything else goes to the grid walk."""
import numpy as np
import pytest

import openkitchen_b200 as ok


def _first_hits(seg, ox, oy, ang):
    """vectorised over segments: (t, index) of CollisionChecker.cu:49-66 for one ray (float64)"""
    dx, dy = np.cos(ang), np.sin(ang)
    sx, sy = seg[:, 2] - seg[:, 0], seg[:, 3] - seg[:, 1]
    den = dx * sy - dy * sx
    ok_den = np.abs(den) >= 1e-8
    den = np.where(ok_den, den, 1.0)
    ex, ey = seg[:, 0] - ox, seg[:, 1] - oy
    t = (ex * sy - ey * sx) / den
    s = (ex * dy - ey * dx) / den
    valid = ok_den & (t >= 0) & (t <= 200.0) & (s >= 0) & (s <= 1)
    if not valid.any():
        return 200.0, -1, np.zeros(0), np.zeros(0, dtype=np.int64)
    tt = np.where(valid, t, np.inf)
    tmin = tt.min()
    hit = np.flatnonzero(valid)
    return float(tmin), int(np.flatnonzero(tt == tmin).max()), t[hit], hit


@pytest.mark.parametrize("name,cell,bins", [("Monza", 8.0, 64), ("Spa", 8.0, 64), ("Zandvoort", 6.0, 32)])
def test_beam_lists_contain_the_reference_winner(name, cell, bins):
    check_beam_lists(ok.Env(device=-1, beam_cell=cell, beam_bins=bins), name)


def test_default_resolution_settles_most_rays_in_the_first_pass():
    """at the default 2 px x 256 bins the four inline candidates decide most rays without the rest of the list"""
    first_pass, decided, mean_len = check_beam_lists(ok.Env(device=-1, beam_cell=2.0, beam_bins=256), "Monza", n_rays=800)
    assert first_pass > 0.7 and decided > 0.99 and mean_len < 12


def check_beam_lists(env, name, n_rays=1500):
    """shared with tests/test_gpu_beam_builder.py (the same property for tables built by the device builder)"""
    t = env.add_named_track(name)
    seg = env.track_array(t, "segments").astype(np.float64)
    x, y, head = env.track_array(t, "x"), env.track_array(t, "y"), env.track_array(t, "heading")
    li, ri = env.track_array(t, "li"), env.track_array(t, "ri")
    rng = np.random.default_rng(7)
    decided, covered, first_pass = 0, 0, 0
    lengths = []
    for _ in range(n_rays):
        i = int(rng.integers(0, len(x)))
        a = rng.uniform(0.05, 0.95)
        ox = np.float32(a * li[i, 0] + (1 - a) * ri[i, 0])
        oy = np.float32(a * li[i, 1] + (1 - a) * ri[i, 1])
        ang = np.float32(np.deg2rad(head[i] + rng.uniform(-180, 180)))
        got = env.beam_lookup_ex(t, float(ox), float(oy), float(ang))
        if got is None:
            continue
        covered += 1
        items, d, d_inline, n_inline = got
        assert env.beam_lookup(t, float(ox), float(oy), float(ang))[0].tolist() == items.tolist()
        assert 0 <= n_inline <= min(4, len(items)) and d_inline <= d + 1e-6
        lengths.append(len(items))
        tmin, best, all_t, all_idx = _first_hits(seg, float(ox), float(oy), float(ang))
        if tmin <= d - 0.25:
            decided += 1
            if best >= 0:
                assert best in items, f"{name}: winner {best} (t={tmin:.3f}) missing from the list (d={d:.3f})"
        # what the kernel's first pass relies on: up to d_inline the entry's inline candidates are the ONLY contenders,
        # i.e. EVERY segment the ray crosses before d_inline - slack is one of them (not just the winner)
        inline = set(items[:n_inline].tolist())
        early = all_idx[all_t <= d_inline - 0.25]
        assert set(early.tolist()) <= inline, f"{name}: segments {sorted(set(early.tolist()) - inline)} are crossed before d_inline={d_inline:.3f}"
        if tmin <= d_inline - 0.25:
            first_pass += 1
        # the list is sorted by a lower bound of the distance and never names a segment twice
        assert len(set(items.tolist())) == len(items)
    assert covered > 0.99 * n_rays, "points inside the lane must be covered by the table"
    assert decided > 0.97 * covered, "the lists should decide almost every ray"
    assert np.mean(lengths) < 60
    assert first_pass > 0
    return first_pass / covered, decided / covered, float(np.mean(lengths))
    assert env.beam_table_bytes(t) > 0


def test_beam_lookup_outside_the_lane_is_uncovered():
    env = ok.Env(device=-1)
    t = env.add_named_track("Monza")
    assert env.beam_lookup(t, 5.0, 5.0, 0.3) is None          # far from the track
    assert env.beam_lookup(t, float("nan"), 5.0, 0.3) is None
    x, y = env.track_array(t, "x"), env.track_array(t, "y")
    assert env.beam_lookup(t, float(x[10]), float(y[10]), 1000.0) is None  # |angle| beyond the table's range
    assert env.beam_lookup(t, float(x[10]), float(y[10]), -3.0) is not None


def test_beam_config_is_validated():
    with pytest.raises(ok.OkError):
        ok.Env(device=-1, beam_bins=48)
    with pytest.raises(ok.OkError):
        ok.Env(device=-1, beam_cell=0.5)


def test_disk_cache_is_shared_between_processes(tmp_path):
    """the ranks of one box build each table once between them: two processes, one cache directory"""
    import os
    import subprocess
    import sys

    code = (
        "import openkitchen_b200 as ok\n"
        "env = ok.Env(device=-1, beam_cell=8.0, beam_bins=32)\n"
        "out = []\n"
        "for nm in ['Monza', 'Spielberg', 'Zandvoort']:\n"
        "    t = env.add_named_track(nm)\n"
        "    items, d = env.beam_lookup(t, float(env.track_array(t, 'x')[5]), float(env.track_array(t, 'y')[5]), 0.5)\n"
        "    out.append((env.beam_table_bytes(t), items.tolist(), d))\n"
        "print(out)\n"
    )
    envv = dict(os.environ, OK_BEAM_CACHE_DIR=str(tmp_path), OK_BEAM_THREADS="2")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    procs = [subprocess.Popen([sys.executable, "-c", code], cwd=root, env=envv, stdout=subprocess.PIPE, text=True) for _ in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs)
    assert outs[0] == outs[1] and "[" in outs[0]
    files = sorted(f for f in os.listdir(tmp_path) if f.endswith(".bin"))
    assert len(files) == 3 and not [f for f in os.listdir(tmp_path) if f.endswith(".lock")]
    # a third process only reads; a cache written with OK_BEAM_CACHE=0 semantics must give the same answer
    again = subprocess.run([sys.executable, "-c", code], cwd=root, env=envv, capture_output=True, text=True, timeout=300)
    fresh = subprocess.run([sys.executable, "-c", code], cwd=root, env=dict(envv, OK_BEAM_CACHE="0"), capture_output=True,
                           text=True, timeout=300)
    assert again.returncode == 0 and fresh.returncode == 0 and again.stdout == outs[0] == fresh.stdout


def _fnv1a(data: bytes) -> int:
    h = 1469598103934665603
    for b in data:
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def test_planted_cache_file_is_rejected(tmp_path, monkeypatch):
    """round-1 advisor finding: a table from the on-disk cache steers device reads.  A file whose checksum is right but
    whose indices leave their arrays (here: a rest-chunk range past the end) must be rejected by the structural check
    and rebuilt; a cache directory other users can write is not used at all."""
    import os
    import struct

    monkeypatch.setenv("OK_BEAM_CACHE_DIR", str(tmp_path))
    monkeypatch.setenv("OK_BEAM_CACHE", "1")

    def lookup():
        ok.release_caches()
        env = ok.Env(device=-1, beam_cell=8.0, beam_bins=32)
        t = env.add_named_track("Zandvoort")
        x, y = env.track_array(t, "x"), env.track_array(t, "y")
        return [env.beam_lookup(t, float(x[i]), float(y[i]), 0.3 * i)[0].tolist() for i in range(0, 200, 10)]

    good = lookup()
    files = [f for f in os.listdir(tmp_path) if f.endswith(".bin")]
    assert len(files) == 1
    path = os.path.join(tmp_path, files[0])
    raw = bytearray(open(path, "rb").read())
    magic, nbytes, checksum = raw[:8], struct.unpack_from("<Q", raw, 8)[0], struct.unpack_from("<Q", raw, 16)[0]
    blob = raw[24:24 + nbytes]
    assert _fnv1a(bytes(blob)) == checksum
    off_entries = struct.unpack_from("<I", blob, 40)[0]
    struct.pack_into("<I", blob, off_entries + 8, 0xFFFFFFF0)  # entry 0: first rest chunk far past the item array
    struct.pack_into("<I", blob, off_entries + 12, struct.unpack_from("<I", blob, off_entries + 12)[0] | (3 << 24))  # ... with 3 chunks
    open(path, "wb").write(bytes(magic) + struct.pack("<QQ", nbytes, _fnv1a(bytes(blob))) + bytes(blob))
    assert lookup() == good  # rejected (bounds), rebuilt
    os.chmod(tmp_path, 0o777)  # a directory anybody can write is not trusted: the table is built, nothing is read or written
    os.remove(path)
    assert lookup() == good
    assert not [f for f in os.listdir(tmp_path) if f.endswith(".bin")]
