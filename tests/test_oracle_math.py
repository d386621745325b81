"""CPU: the scalar pieces of the oracle -- sincosf restatement vs the libm the reference links, Philox KAT,
nearest-index tie rule, angle normalisation."""
import ctypes as C
import math

import numpy as np
import pytest

import openkitchen_b200 as ok
from oracle.api import Oracle, have_ref


def _bits(x):
    return np.float32(x).view(np.uint32)


def test_sincosf_restatement_matches_libm_on_2e6_inputs():
    """oko_sincosf is the glibc 2.39 algorithm (FMA ifunc variant).  The exhaustive 2^32 comparison is
    oracle/sincosf_exhaustive.c (result recorded in DESIGN.md); here: 2M random bit patterns plus the
    branch boundaries, against the libm of this box."""
    o = Oracle("port")
    libm = C.CDLL("libm.so.6")
    libm.sincosf.argtypes = [C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    rng = np.random.default_rng(0)
    bits = rng.integers(0, 2**32, size=200_000, dtype=np.uint64).astype(np.uint32)
    edge = np.array([0.0, -0.0, 2.0**-12, np.pi / 4, 0.78539818, 0.785398, 119.99999, 120.0, 120.00001, 1e9, -1e9, 3.4e38,
                     1e-40, np.inf, np.nan, 1.5707964, 3.1415927, 6.2831855], dtype=np.float32)
    degs = (rng.uniform(-4000, 4000, size=100_000).astype(np.float32) * np.float32(math.pi / 180.0)).astype(np.float32)
    vals = np.concatenate([bits.view(np.float32), edge, np.nextafter(edge, np.float32(np.inf)), degs])
    s1, c1, s2, c2 = C.c_float(), C.c_float(), C.c_float(), C.c_float()
    bad = 0
    for v in vals:
        o._sincosf(C.c_float(v), C.byref(s1), C.byref(c1))
        libm.sincosf(C.c_float(v), C.byref(s2), C.byref(c2))
        for p, q in ((s1.value, s2.value), (c1.value, c2.value)):
            if not ((math.isnan(p) and math.isnan(q)) or _bits(p) == _bits(q)):
                bad += 1
    assert bad == 0, f"{bad} mismatches against libm sincosf"


def test_philox_known_answers():
    """Random123 known-answer vectors for Philox4x32-10"""
    o = Oracle("port")
    assert list(o.philox([0, 0, 0, 0], [0, 0])) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert list(o.philox([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2)) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert list(o.philox([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0])) == [
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_nearest_index_tie_rule_and_helpers():
    o = Oracle("port")
    # a straight track with a duplicated point: strict '<' keeps the LOWEST index (RaceTrack.cpp:24-28)
    x = np.array([0, 10, 20, 20, 30, 40], dtype=np.float32)
    y = np.zeros(6, dtype=np.float32)
    w = np.full(6, 5.0, dtype=np.float32)
    t = o.add_track((x, y, w, w))
    px, py = o.track_array(t, "x"), o.track_array(t, "y")
    assert px[2] == px[3] and py[2] == py[3]
    assert o.nearest_index(t, float(px[2]), float(py[2]) + 1.0) == 2
    assert o.nearest_index(t, float(px[5]) + 100.0, float(py[5])) == 5
    d = o.dist_lane_center(t, float(px[1]), float(py[1]) + 3.0)
    wl, wr = o.track_array(t, "w_left"), o.track_array(t, "w_right")
    assert d == pytest.approx(3.0 / float(wl[1] + wr[1]), rel=1e-6)
    assert o.normalize_angle_deg(-30.0) == 330.0 and o.normalize_angle_deg(725.0) == 5.0 and o.normalize_angle_deg(360.0) == 0.0


def test_width_clamp_and_fit_to_window():
    """RaceTrack.cpp:138-160,198-229: widths min(max(4,w)*3,17)*scale; geometry centred in 1600x1400"""
    o = Oracle("port")
    x = np.array([0, 100, 200, 200, 100, 0], dtype=np.float32)
    y = np.array([0, 0, 50, 150, 200, 100], dtype=np.float32)
    wr = np.array([1.0, 4.0, 5.0, 5.7, 9.0, 100.0], dtype=np.float32)
    t = o.add_track((x, y, wr, wr))
    scale = np.float32(0.9) * min(np.float32(1600) / np.float32(200), np.float32(1400) / np.float32(200))
    exp = np.minimum(np.maximum(np.float32(4.0), wr) * np.float32(3.0), np.float32(17.0)) * scale
    assert np.array_equal(o.track_array(t, "w_right"), exp.astype(np.float32))
    px, py = o.track_array(t, "x"), o.track_array(t, "y")
    assert (px.max() + px.min()) / 2 == pytest.approx(800.0, abs=1e-3) and (py.max() + py.min()) / 2 == pytest.approx(700.0, abs=1e-3)
    assert o.track_segments(t) == 4 * 5 + 4


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
def test_reference_rejects_nothing_silently():
    r = Oracle("reference")
    t = r.add_track(ok.track_columns("Monza"))
    assert r.track_points(t) == 1159 and r.track_segments(t) == 4636
