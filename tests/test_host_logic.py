"""CPU-only checks of the product library: it loads, exports every symbol include/b2pbr.h declares, fails
loudly without a GPU (no CPU fallback), and its host-side tile scheduler (rectangle algebra) matches the
reference's known answers and the oracle."""
import ctypes as C
import os
import re
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "b2pbr.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2(?:pbr|mip|layer|band)?_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(b2):
    from dplug_b200 import _lib
    l = C.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(l, n)]
    assert not missing, missing
    # and the Python binding table covers exactly the header
    assert sorted(_lib.SIGNATURES) == names
    assert _lib.lib().b2_abi_version() == 2


def test_no_cpu_fallback(b2):
    """Without a CUDA device creation must fail with B2_ERR_NO_DEVICE — never silently compute on the CPU."""
    from dplug_b200 import _lib
    l = _lib.lib()
    if l.b2_device_count() > 0:
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    assert l.b2pbr_create(C.byref(h), 0) == 3
    assert b"no CUDA device" in l.b2_last_error()
    assert l.b2mip_create(C.byref(h), 0, 0, 4, 64, 64) == 3
    with pytest.raises(b2.B2Error):
        b2.PBRCompositor(0)
    with pytest.raises(b2.B2Error):
        b2.Mipmap(b2.RGBA8, 4, 64, 64)


def test_product_does_not_link_oracle():
    """The shipped library and package must not reference oracle/ (parity claims depend on it)."""
    pkg = os.path.join(ROOT, "dplug_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "b2ref" not in text and "oracle/" not in text.replace("oracle/rsqrtss_table.inc", ""), (dirpath, f)


def test_exact_mip_kernels_contain_no_fused_multiply_add(b2):
    """The alpha-premultiplying box must be a chain of single IEEE multiplies and adds (mipmap.d:869-894).  ptxas
    (CUDA 12.9) contracts a packed FMUL2 feeding a packed FADD2 into FFMA2 even with explicit .rn and -fmad=false
    (dplug_b200/csrc/mip_math.cuh explains the rule that avoids it); the built objects must hold no FFMA at all."""
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    for obj in ("mip_kernels.o", "mip_fused.o", "mip_wave.o"):
        path = os.path.join(ROOT, "dplug_b200", "build", obj)
        if not os.path.exists(path):
            pytest.skip("objects not kept (library built elsewhere)")
        sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
        # the only fused form allowed is mulToAdd2: FFMA2 a, b, <register holding the (-0, -0) kernel argument>.  A
        # contracted multiply-add shows up as a scalar FFMA, or as FFMA2 whose addends are the many different
        # registers of running sums; the legitimate ones all use the one or two registers (pairs) holding -0
        scalar = [ln.strip() for ln in sass.splitlines() if re.search(r"\bFFMA\b", ln)]
        assert not scalar, (obj, scalar[:3])
        fn, per_fn = None, {}
        for ln in sass.splitlines():
            m = re.search(r"Function : (\S+)", ln)
            if m:
                fn = m.group(1)
            elif re.search(r"\bFFMA2\b", ln):
                add = ln.split(";")[0].split(",")[-1].strip()
                per_fn.setdefault(fn, set()).add(re.sub(r"\.reuse|\.F32x2\.HI_LO|\.F32", "", add))
        assert per_fn, obj
        for fn, addends in per_fn.items():
            assert len(addends) <= 4 and all(a.startswith("R") for a in addends), (obj, fn, sorted(addends))
        assert "FMUL2" in sass and "FADD2" in sass                        # the packed pipe is actually used


def test_blackwell_native_paths_in_sass(b2):
    """SASS evidence of the hardware paths DESIGN.md claims: the large cubic levels are fed by the TMA engine (1-D bulk
    asynchronous copies completing on mbarriers: UBLKCP + SYNCS), and the lighting kernel's mip / skybox samplers run on
    the texture units (TEX for the bilinear fetches, TLD4 for the gather variant); the default lighting kernel computes two
    rows per lane with packed FP32 instructions (FFMA2 / FADD2 / FMUL2)."""
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    want = {"mip_kernels.o": ("UBLKCP", "SYNCS"), "pbr_tex.o": ("TEX", "TLD4", "FFMA2", "FADD2", "FMUL2")}
    for obj, mnemonics in want.items():
        path = os.path.join(ROOT, "dplug_b200", "build", obj)
        if not os.path.exists(path):
            pytest.skip("objects not kept (library built elsewhere)")
        sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
        for m in mnemonics:
            assert re.search(r"\b%s\b" % m, sass), (obj, m)


def test_remove_overlapping_areas_known_answer(b2):
    # gui/dplug/gui/boxlist.d:120-135
    assert b2.removeOverlappingAreas([(0, 0, 4, 4), (2, 2, 6, 6), (1, 1, 2, 2)]) == [(2, 2, 6, 6), (0, 0, 4, 2), (0, 2, 2, 4)]


def test_tile_areas(b2, oracle):
    assert len(b2.tileAreas([(0, 0, 3840, 2160)], 64, 64)) == 2040
    t = b2.tileAreas([(0, 0, 620, 330)], 64, 64)
    assert len(t) == 60 and t[0] == (0, 0, 64, 64) and t[9] == (576, 0, 620, 64) and t[59] == (576, 320, 620, 330)
    rng = np.random.default_rng(11)
    for _ in range(50):
        rects = []
        for _ in range(int(rng.integers(1, 5))):
            x0, y0 = int(rng.integers(0, 500)), int(rng.integers(0, 300))
            rects.append((x0, y0, x0 + int(rng.integers(1, 300)), y0 + int(rng.integers(1, 200))))
        mw, mh = int(rng.integers(1, 100)), int(rng.integers(1, 100))
        out = (oracle.box2i * 100000)()
        n = oracle.lib().b2ref_tile_areas(oracle.boxes(rects), len(rects), mw, mh, out, 100000)
        assert b2.tileAreas(rects, mw, mh) == [out[i].astuple() for i in range(n)]


def test_remove_overlapping_random_vs_oracle(b2, oracle):
    rng = np.random.default_rng(12)
    for _ in range(200):
        rects = []
        for _ in range(int(rng.integers(1, 9))):
            x0, y0 = int(rng.integers(0, 40)), int(rng.integers(0, 40))
            rects.append((x0, y0, x0 + int(rng.integers(0, 30)), y0 + int(rng.integers(0, 30))))
        out = (oracle.box2i * 4096)()
        n = oracle.lib().b2ref_remove_overlapping_areas(oracle.boxes(rects), len(rects), out, 4096)
        exp = [out[i].astuple() for i in range(n)]
        got = b2.removeOverlappingAreas(rects)
        assert got == exp
        # property: disjoint, same coverage
        cover = np.zeros((80, 80), dtype=np.int32)
        for (a, b, c, d) in got:
            cover[b:d, a:c] += 1
        assert cover.max() <= 1
        want = np.zeros((80, 80), dtype=bool)
        for (a, b, c, d) in rects:
            want[b:d, a:c] = True
        assert np.array_equal(cover.astype(bool), want)


def test_impact_on_next_level_vs_oracle(b2, oracle):
    rng = np.random.default_rng(13)
    for _ in range(500):
        w, h = int(rng.integers(1, 200)), int(rng.integers(1, 200))
        x0, y0 = int(rng.integers(0, w)), int(rng.integers(0, h))
        r = (x0, y0, int(rng.integers(x0, w + 1)), int(rng.integers(y0, h + 1)))
        for q in (0, 1, 2):
            assert b2.impactOnNextLevel(q, r, w, h) == oracle.lib().b2ref_impact_on_next_level(q, oracle.box2i(*r), w, h).astuple()


def test_grow_rect(b2):
    # gui/dplug/gui/graphics.d:981-1002 with the default 20-px update margin (graphics.d:719)
    assert b2.growRect((100, 100, 150, 130), 20, 620, 330) == (80, 80, 170, 150)
    assert b2.growRect((5, 5, 615, 325), 20, 620, 330) == (0, 0, 620, 330)
    assert b2.growRect((-100, -100, -50, -50), 20, 620, 330) == (0, 0, 0, 0)


def test_default_params_match_reference_defaults(b2, oracle):
    p, q = b2.default_params(), oracle.default_params()
    assert bytes(p) == bytes(q)
    assert abs(p.light1_color[0] - 0.325) < 1e-6 and p.pass_active_mask == 0xff
    assert abs(p.light2_dir[1] - 0.99503719) < 1e-6


def test_have_no_overlap_known_answer(b2, oracle):
    # gui/dplug/gui/boxlist.d:188-192
    assert b2.haveNoOverlap([(0, 0, 1, 1), (1, 1, 2, 2)])
    assert not b2.haveNoOverlap([(0, 0, 1, 1), (0, 0, 2, 1)])
    assert b2.haveNoOverlap([])


def _coverage(rects, size=64):
    m = np.zeros((size, size), dtype=np.int32)
    for (x0, y0, x1, y1) in rects:
        m[y0:y1, x0:x1] += 1
    return m


def test_dirty_rect_list_matches_reference_algorithm(b2, oracle):
    """DirtyRectList.addRect (boxlist.d:236-292): same list, in the same order, as the restated reference algorithm;
    the invariant (no overlaps) and the coverage (union of everything added) hold after every call."""
    rng = np.random.default_rng(17)
    for trial in range(30):
        g, o = b2.DirtyRectList(), oracle.DirtyRectList()
        assert g.isEmpty()
        union = np.zeros((64, 64), dtype=bool)
        for k in range(int(rng.integers(1, 25))):
            x0, y0 = int(rng.integers(0, 60)), int(rng.integers(0, 60))
            r = (x0, y0, int(rng.integers(x0, 65)), int(rng.integers(y0, 65)))   # may be empty: ignored (boxlist.d:241)
            g.addRect(r); o.addRect(r)
            union[r[1]:r[3], r[0]:r[2]] = True
        assert g.isEmpty() == (len(o.rects) == 0)
        got = g.pullAllRectangles()
        assert got == o.rects
        assert g.isEmpty() and g.pullAllRectangles() == []
        cov = _coverage(got)
        assert cov.max() <= 1 and b2.haveNoOverlap(got)
        assert np.array_equal(cov > 0, union)


def test_dirty_rect_list_cases(b2):
    g = b2.DirtyRectList()
    g.addRect((10, 10, 20, 20))
    g.addRect((12, 12, 15, 15))            # inside an element: discarded
    assert g.pullAllRectangles() == [(10, 10, 20, 20)]
    g.addRect((12, 12, 15, 15)); g.addRect((10, 10, 20, 20))   # contains an element: the element goes
    assert g.pullAllRectangles() == [(10, 10, 20, 20)]
    g.addRect((0, 0, 10, 10)); g.addRect((5, 5, 15, 15))       # partial overlap: the old one is cut (above + left parts)
    assert g.pullAllRectangles() == [(0, 0, 10, 5), (0, 5, 5, 10), (5, 5, 15, 15)]
    with pytest.raises(b2.B2Error):
        g.addRect((5, 5, 1, 1))            # unsorted (assert in the reference, boxlist.d:239)
