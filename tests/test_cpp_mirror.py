"""The C++ host mirror (include/dplug_b200/compositor.hpp) drives the same C ABI as the Python one."""
import os
import subprocess
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    exe = str(tmp_path / "host_mirror")
    libdir = os.path.join(ROOT, "dplug_b200")
    subprocess.run(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "host_mirror.cpp"),
                    "-o", exe, os.path.join(libdir, "libb2pbr.so"), "-Wl,-rpath," + libdir], check=True)
    return exe


def test_cpp_mirror_builds_and_fails_loudly_without_gpu(b2, tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.gpu
def test_cpp_mirror_matches_python_mirror(b2, tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe, "gpu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    got = int(r.stdout.split()[-1])
    W, H = 160, 96
    y, x = np.mgrid[0:H, 0:W]
    d = np.stack([(x * 3 + y) & 255, (x + y * 2) & 255, (x ^ y) & 255, np.where((x // 16 + y // 16) % 7 == 0, 255, 0)], axis=-1).astype(np.uint8)
    m = np.empty((H, W, 4), dtype=np.uint8); m[...] = (128, 64, 128, 255)
    z = (15000 + (x * 37 + y * 91) % 3000).astype(np.uint16)
    c = b2.PBRCompositor(0)
    c.resizeBuffers(W, H, 64, 64)
    out = np.zeros((H, W, 4), dtype=np.uint8)
    c.compositeTile(out, b2.tileAreas([(0, 0, W, H)], 64, 64), d, m, z)
    v = out.astype(np.uint64)
    px = (v[..., 0] + 256 * v[..., 1] + 65536 * v[..., 2] + 16777216 * v[..., 3]).ravel()
    s = 0
    for p in px.tolist():
        s = (s * 1000003 + p) & 0xFFFFFFFFFFFFFFFF
    assert s == got
