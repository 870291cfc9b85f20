"""The work list of the whole-pyramid wavefront kernel (dplug_b200/csrc/mip_wave.cu, host part), checked without a GPU:
the list is replayed in order by a model of the kernel's queue.  Every item must find the row chunks it waits for
already complete when it is reached (list order = a topological order: that is what makes the persistent grid
deadlock-free), every destination texel of every level's update rectangle (graphics/dplug/graphics/mipmap.d:623-645
chain) must be written exactly once, the source rows an item reads inside the previous level's rectangle must belong to
chunks it waits for, and all counters end where the last CTA's reset expects them."""
import ctypes as C

import numpy as np
import pytest

from dplug_b200 import _lib

BOX, CUBIC, PREMUL = 0, 1, 2
KIND_BLOCK = {0: 64, 1: 64, 2: 128, 3: 128, 4: 64}       # texels per virtual block along x (mip_bodies.cuh)
KIND_UNIT = {0: 2, 1: 2, 2: 4, 3: 4, 4: 2}               # texels per thread: the grid origin is rounded down to it
KIND_CUBIC = {0: False, 1: False, 2: True, 3: False, 4: True}


def _plan(color, w, h, first, q, rect):
    l = _lib.lib()
    qa = (C.c_int * len(q))(*q)
    nctl = C.c_int(0)
    n = l.b2_mip_wave_plan(color, w, h, first, len(q), qa, _lib.box2i(*rect), None, 0, C.byref(nctl))
    assert n >= 0, _lib.last_error() if hasattr(_lib, "last_error") else n
    buf = (C.c_int * (8 * max(n, 1)))()
    assert l.b2_mip_wave_plan(color, w, h, first, len(q), qa, _lib.box2i(*rect), buf, n, C.byref(nctl)) == n
    return np.frombuffer(buf, dtype=np.int32).reshape(-1, 8)[:n].copy(), nctl.value


def _chain(w, h, first, q, rect):
    l = _lib.lib()
    out = []
    r = _lib.box2i(*rect)
    for k, qq in enumerate(q):
        lvl = first - 1 + k
        r = l.b2_impact_on_next_level(qq, r, w >> lvl, h >> lvl)
        out.append((r.min_x, r.min_y, r.max_x, r.max_y))
    return out


@pytest.mark.parametrize("color,w,h,first,q,rect", [
    (0, 8192, 8192, 1, [PREMUL] + [CUBIC] * 7, (0, 0, 8192, 8192)),          # BASELINE config 3
    (0, 3840, 2160, 1, [PREMUL] + [CUBIC] * 4, (0, 0, 3840, 2160)),          # config 2, diffuse
    (1, 3840, 2160, 1, [BOX, BOX, CUBIC, CUBIC], (0, 0, 3840, 2160)),        # config 2, depth
    (0, 1237, 1011, 1, [BOX, BOX, CUBIC, CUBIC, CUBIC, CUBIC, CUBIC], (0, 0, 1237, 1011)),
    (0, 1700, 1300, 1, [PREMUL] + [CUBIC] * 4, (133, 77, 1651, 1299)),       # a dirty rectangle
    (1, 1700, 1300, 1, [BOX, BOX, CUBIC, CUBIC], (1, 3, 1700, 1211)),
    (0, 2600, 2100, 2, [CUBIC, BOX, CUBIC, CUBIC], (3, 5, 1290, 1040)),      # starting above level 1
    (1, 4099, 2050, 1, [CUBIC, CUBIC, BOX, CUBIC, BOX], (0, 0, 4099, 2050)),
])
def test_plan_replays_in_order(color, w, h, first, q, rect):
    items, nctl = _plan(color, w, h, first, q, rect)
    rects = _chain(w, h, first, q, rect)
    assert len(items) > 0
    ctl = np.zeros(nctl, dtype=np.int64)
    written = [np.zeros(((h >> (first + k)), (w >> (first + k))), dtype=np.uint8) for k in range(len(q))]
    chunk_of = {}                                          # counter index -> (step, block row)
    for it in items:
        step, kind = int(it[0]) & 0xffff, int(it[0]) >> 16
        by, bx0, bx1, dep_lo, dep_hi, target, done = (int(v) for v in it[1:])
        assert (kind in (0, 1, 2)) == (color == 0)
        rx0, ry0, rx1, ry1 = rects[step]
        assert bx1 > bx0 and 2 <= done < nctl
        chunk_of.setdefault(done, (step, by))
        assert chunk_of[done] == (step, by)
        # the kernel's wait: every source chunk complete when the item is reached in list order
        assert dep_hi - dep_lo + 1 <= 4
        for c in range(dep_lo, dep_hi + 1):
            assert ctl[c] == target and target > 0, (step, by, c, int(ctl[c]), target)
        # rows this item reads inside the previous level's rectangle belong to the chunks it waits for
        y0, y1 = ry0 + 8 * by, min(ry0 + 8 * by + 8, ry1)
        assert y0 < y1
        if step > 0:
            src_h = h >> (first - 1 + step)
            lo, hi = (2 * y0 - 1, 2 * (y1 - 1) + 2) if KIND_CUBIC[kind] else (2 * y0, 2 * (y1 - 1) + 1)
            lo, hi = max(lo, 0), min(hi, src_h - 1)
            px0, py0, px1, py1 = rects[step - 1]
            if px1 > px0 and py1 > py0:
                waited = {chunk_of[c][1] for c in range(dep_lo, dep_hi + 1)}
                for c in range(dep_lo, dep_hi + 1):
                    assert chunk_of[c][0] == step - 1
                for row in range(max(lo, py0), min(hi, py1 - 1) + 1):
                    assert (row - py0) // 8 in waited, (step, by, row)
        else:
            assert dep_hi < dep_lo
        # texels the bodies of these virtual blocks write
        unit, bw = KIND_UNIT[kind], KIND_BLOCK[kind]
        origin = (rx0 // unit) * unit
        xa, xb = max(rx0, origin + bx0 * bw), min(rx1, origin + bx1 * bw)
        assert xa < xb
        written[step][y0:y1, xa:xb] += 1
        ctl[done] += 1
    for k, (rx0, ry0, rx1, ry1) in enumerate(rects):
        want = np.zeros_like(written[k])
        if rx1 > rx0 and ry1 > ry0:
            want[ry0:ry1, rx0:rx1] = 1
        assert np.array_equal(written[k], want), f"step {k}: texels written {int(written[k].sum())}, wanted {int(want.sum())}"
    # wavefront property: the next level follows closely (its first item comes long before the previous level's last)
    steps = items[:, 0] & 0xffff
    if len(q) > 1 and (rects[0][3] - rects[0][1]) >= 256:
        first_of_next = int(np.argmax(steps == 1))
        last_of_first = len(steps) - 1 - int(np.argmax(steps[::-1] == 0))
        assert first_of_next < last_of_first // 8
