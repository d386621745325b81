"""CPU: the laser2d dataset writer produces byte-for-byte what the reference's C++ collector would
(std::ofstream << float formatting), checked against a tiny program compiled with g++."""
import os
import subprocess

import numpy as np

from openkitchen_b200 import datasets

CPP = r'''
#include <fstream>
#include <cstdio>
int main(int argc, char** argv) {
    // same statements as FieldNavigators/collect_data/collect_data_random.cpp:79-86
    FILE* in = fopen(argv[1], "rb"); int n; fread(&n, 4, 1, in);
    std::ofstream out{argv[2]};
    for (int i = 0; i < n; ++i) { float x, y; fread(&x, 4, 1, in); fread(&y, 4, 1, in); out << x << " " << y << std::endl; }
    float t, s; fread(&t, 4, 1, in); fread(&s, 4, 1, in);
    out << t << " " << s;
    return 0;
}
'''


def test_laser2d_matches_cxx_stream_formatting(tmp_path):
    src = tmp_path / "w.cpp"
    src.write_text(CPP)
    exe = tmp_path / "w"
    subprocess.run(["/usr/bin/g++", "-O1", "-o", str(exe), str(src)], check=True)
    rng = np.random.default_rng(0)
    hits = np.concatenate([rng.normal(0, 60, size=(13, 2)), [[0.0, -0.0], [200.0, 1e-7], [123456.789, -0.000012345], [1e10, 3.0]]]).astype(np.float32)
    act = np.array([57.25, -3.999999], dtype=np.float32)
    raw = tmp_path / "in.bin"
    with open(raw, "wb") as f:
        f.write(np.int32(len(hits)).tobytes() + hits.tobytes() + act.tobytes())
    subprocess.run([str(exe), str(raw), str(tmp_path / "ref.txt")], check=True)
    want = open(tmp_path / "ref.txt").read()
    got = datasets.format_laser2d(hits, act[0], act[1])
    assert got == want
    nxt = datasets.write_laser2d(str(tmp_path / "d"), "MONZA", hits[None], act[:1], act[1:], first_index=7)
    assert nxt == 8 and open(tmp_path / "d" / "laser2d_MONZA_7.txt").read() == want
    back, (t, s) = datasets.read_laser2d(str(tmp_path / "d" / "laser2d_MONZA_7.txt"))
    assert back.shape == hits.shape and np.allclose(back, hits, rtol=1e-5, atol=1e-6) and abs(t - 57.25) < 1e-4
