"""Timing of the mip chains with the wavefront kernel on / off (one GPU): BASELINE config 3 (8192^2, both data sets) and
the two chains of a 3840x2160 frame.  PROBE_LIB=<path> loads another build of libb2pbr.so (compile-time variants)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

from dplug_b200 import _lib
if os.environ.get("PROBE_LIB"):
    _lib.LIB_PATH = os.path.abspath(os.environ["PROBE_LIB"])
import dplug_b200 as b2
from dplug_b200 import synth

BOX, CUBIC, PREMUL = 0, 1, 2


def timed(fn, steps=40, warm=4):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps * 1e3


stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
rng = np.random.default_rng(1)
out = []
which = os.environ.get("PROBE_WHICH", "c3s,c3r,c2d,c2z").split(",")
for name, color, (W, H), recipe, fill in [
    ("c3s", 0, (8192, 8192), [PREMUL] + [CUBIC] * 7, "survey"),
    ("c3r", 0, (8192, 8192), [PREMUL] + [CUBIC] * 7, "random"),
    ("c2d", 0, (3840, 2160), [PREMUL] + [CUBIC] * 4, "survey"),
    ("c2z", 1, (3840, 2160), [BOX, BOX, CUBIC, CUBIC], "depth"),
]:
    if name not in which:
        continue
    R = 2 if W == 8192 else 6
    mips = []
    for k in range(R):
        m = b2.Mipmap(color, len(recipe), W, H)
        m.set_stream(stream.cuda_stream)
        rows = 1024 if H >= 1024 else H
        if fill == "survey":
            blk = synth.diffuse(W, rows, seed=synth.SEED + k)
        elif fill == "depth":
            blk = synth.depth(W, rows, seed=synth.SEED + k)
        else:
            blk = rng.integers(0, 256, size=(rows, W, 4), dtype=np.uint8)
        for y0 in range(0, H, rows):
            m.uploadRows(0, blk[: min(rows, H - y0)], y0)
        mips.append(m)
    i = [0]

    def step():
        mips[i[0] % R].generateLevels(1, recipe, (0, 0, W, H))
        i[0] += 1

    res = {}
    for wave in (1, 0):
        _lib.lib().b2_set_mip_wave(wave)
        res[wave] = timed(step)
    out.append("%s wave %.1f us  levels %.1f us" % (name, res[1], res[0]))
    del mips
print(os.environ.get("PROBE_LIB", "default"), " | ".join(out), flush=True)
