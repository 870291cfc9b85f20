#!/usr/bin/env python
"""Generates the committed golden fixtures under tests/golden/ from the CPU oracle.

The reference (pure D) cannot be run in this environment, so these vectors pin the ORACLE's behaviour
against accidental drift; they are not outputs of the reference itself ("parity status").

  python tools/make_golden.py        # rewrites tests/golden/*.npz
"""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dplug_b200 import synth  # noqa: E402
from tests import oracle  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def frame_case(W, H, kind, seed):
    d, m, z = synth.frame(W, H, seed, kind)
    sky = synth.skybox(64, 48, seed)
    g = oracle.Graphics(W, H, numThreads=1)
    g.setSkybox(sky)
    g.setInputs(d, m, z)
    out = g.fullFrame()
    lv = {"diffuse_l%d" % l: g.diffuseMap.download(l) for l in range(1, g.diffuseMap.numLevels())}
    lv.update({"depth_l%d" % l: g.depthMap.download(l) for l in range(1, g.depthMap.numLevels())})
    return dict(diffuse=d, material=m, depth=z, skybox=sky, composited=out, **lv)


def main():
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "frame_96x80_smooth.npz"), **frame_case(96, 80, "smooth", synth.SEED))
    np.savez_compressed(os.path.join(OUT, "frame_70x45_noise.npz"), **frame_case(70, 45, "noise", synth.SEED + 7))
    # mip kernels on an odd-sized random image
    rng = np.random.default_rng(2024)
    rgba = rng.integers(0, 256, size=(37, 53, 4), dtype=np.uint8)
    l16 = rng.integers(0, 65536, size=(37, 53), dtype=np.uint16)
    res = {"rgba": rgba, "l16": l16}
    for name, color, src in (("rgba", oracle.RGBA8, rgba), ("l16", oracle.L16, l16)):
        for q, qn in ((0, "box"), (1, "cubic"), (2, "premul")):
            if color == oracle.L16 and q == 2:
                continue
            m = oracle.Mipmap(color, 1, 53, 37)
            m.upload(0, src)
            m.generateNextLevel(q, (0, 0, 53, 37), 1)
            res["%s_%s" % (name, qn)] = m.download(1)
    np.savez_compressed(os.path.join(OUT, "mips_53x37.npz"), **res)
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
