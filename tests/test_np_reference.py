"""Oracle pinning: the C++ restatement (oracle/) against the independent numpy float32 restatement (tools/np_reference.py,
written from the D sources with whole-image array operations).  They must agree BIT FOR BIT — mip levels, samplers, and
the composited frame — on small frames that include every edge case of the clamp logic."""
import os
import sys
import numpy as np
import pytest
from dplug_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import np_reference as npr  # noqa: E402


@pytest.mark.parametrize("W,H,kind,seed", [(70, 45, "noise", 1), (96, 80, "smooth", 2), (33, 37, "smooth", 3), (64, 64, "noise", 4), (131, 67, "smooth", 5)])
def test_frame_bit_identical(oracle, W, H, kind, seed):
    d, m, z = synth.frame(W, H, synth.SEED + seed, kind)
    sky = synth.skybox(64, 48, synth.SEED + seed)
    ref = oracle.Graphics(W, H, numThreads=4)
    ref.setSkybox(sky); ref.setInputs(d, m, z)
    want = ref.fullFrame()
    D, Z = npr.diffuse_pyramid(d), npr.depth_pyramid(z)
    for l in range(1, len(D)):
        assert np.array_equal(D[l], ref.diffuseMap.download(l)), ("diffuse level", l)
    for l in range(1, len(Z)):
        assert np.array_equal(Z[l], ref.depthMap.download(l)), ("depth level", l)
    got = npr.composite(d, m, z, sky)
    diff = np.abs(got.astype(np.int16) - want.astype(np.int16))
    assert np.array_equal(got, want), "numpy restatement and oracle differ: max %d LSB on %d channels, first at %s" % (
        diff.max(), int((diff > 0).sum()), np.argwhere(diff > 0)[:3].tolist())


def test_frame_without_skybox_and_small_canvas(oracle):
    for W, H in ((40, 33), (17, 9)):      # the second canvas has incomplete pyramids (levels clamp, mipmap.d:296-300)
        d, m, z = synth.frame(W, H, synth.SEED + 7, "noise")
        ref = oracle.Graphics(W, H, numThreads=2)
        ref.setInputs(d, m, z)
        assert np.array_equal(npr.composite(d, m, z, None), ref.fullFrame()), (W, H)


def test_samplers_bit_identical(oracle):
    rng = np.random.default_rng(5)
    W, H = 90, 61
    img = rng.integers(0, 256, size=(H, W, 4), dtype=np.uint8)
    dep = rng.integers(0, 65536, size=(H, W), dtype=np.uint16)
    a = oracle.Mipmap(oracle.RGBA8, 5, W, H); a.upload(0, img); a.generateMipmaps(oracle.Q_BOX)
    b = oracle.Mipmap(oracle.L16, 4, W, H); b.upload(0, dep); b.generateMipmaps(oracle.Q_BOX)
    A = [a.download(l) for l in range(a.numLevels())]
    B = [b.download(l) for l in range(b.numLevels())]
    sky = npr.skybox_pyramid(img)          # generateMipmaps(box): box for levels 1, 2, cubic from level 3 (mipmap.d:517-528)
    for l in range(len(A)):
        assert np.array_equal(sky[l], A[l]), l
    x = rng.uniform(-3, W + 3, 400).astype(np.float32)
    y = rng.uniform(-3, H + 3, 400).astype(np.float32)
    for level in range(0, 6):
        assert np.array_equal(npr.linear_sample(A, level, x, y), a.linearSample(level, x, y)), level
        assert np.array_equal(npr.cubic_sample(A, level, x, y), a.cubicSample(level, x, y)), level
        assert np.array_equal(npr.linear_sample(B, level, x, y), b.linearSample(level, x, y)), level
        assert np.array_equal(npr.cubic_sample(B, level, x, y), b.cubicSample(level, x, y)), level
    lv = rng.uniform(0, 5.5, 400).astype(np.float32)
    lv[:40] = np.floor(lv[:40])            # integer levels take the early return
    assert np.array_equal(npr.linear_mipmap_sample(A, lv, x, y), a.linearMipmapSample(lv, x, y))


def test_math_helpers_bit_identical(oracle):
    import ctypes as C
    l = oracle.lib()
    rng = np.random.default_rng(9)
    xs = np.concatenate([rng.uniform(1e-3, 1.0, 3000), rng.uniform(0.3, 4.0, 500)]).astype(np.float32)
    ys = rng.uniform(0.01, 548.0, xs.size).astype(np.float32)
    want = np.array([l.b2ref_pow_ps(C.c_float(float(a)), C.c_float(float(b))) for a, b in zip(xs, ys)], dtype=np.float32)
    assert np.array_equal(npr.pow_ps(xs, ys).view(np.uint32), want.view(np.uint32))
    v = rng.uniform(1e-6, 1e6, 2000).astype(np.float32)
    want = np.array([l.b2ref_fastlog2(C.c_float(float(a))) for a in v], dtype=np.float32)
    assert np.array_equal(npr.fastlog2(v), want)
    want = np.array([l.b2ref_rsqrtss(C.c_float(float(a)), 0) for a in v], dtype=np.float32)
    assert np.array_equal(npr.rsqrt_ss(v), want)


def test_committed_goldens_are_the_numpy_restatement():
    """tests/golden/*.npz (which the oracle and the CUDA path are tested against) are what tools/make_golden.py produces
    from the numpy restatement — not the C++ oracle's own output."""
    import subprocess
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_golden.py"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]


def test_pow_ps_against_third_party_sse_mathfun_port(tmp_path):
    """_mm_pow_ps = exp_ps(y * log_ps(x)) of sse_mathfun: the restatement against the AVX2 port of the same algorithm
    that ships with PyTorch (ATen/native/cpu/avx_mathfun.h), 10^6 operands over the specular pass's domain."""
    import ctypes as C
    import subprocess
    try:
        import torch
    except Exception:
        pytest.skip("torch headers unavailable")
    inc = os.path.join(os.path.dirname(torch.__file__), "include")
    if not os.path.exists(os.path.join(inc, "ATen", "native", "cpu", "avx_mathfun.h")) or "avx2" not in open("/proc/cpuinfo").read():
        pytest.skip("avx_mathfun.h or AVX2 unavailable")
    so = str(tmp_path / "avx_pow.so")
    subprocess.run(["g++", "-O2", "-mavx2", "-ffp-contract=off", "-DCPU_CAPABILITY_AVX2", "-I" + inc, "-shared", "-fPIC", "-o", so,
                    os.path.join(ROOT, "tools", "pin", "avx_pow.cpp")], check=True)
    l = C.CDLL(so)
    rng = np.random.default_rng(1)
    n = 1_000_000
    x = rng.uniform(1e-3, 1.0, n).astype(np.float32)
    y = rng.uniform(0.01, 548.0, n).astype(np.float32)
    out = np.zeros(n, np.float32)
    l.avx_pow(x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), n)
    assert np.array_equal(out.view(np.uint32), npr.pow_ps(x, y).view(np.uint32))
