#!/usr/bin/env python
"""Writes the level-0 inputs of the golden cases as raw files for tools/d/dump_fixtures.d (see its header)."""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from dplug_b200 import synth  # noqa: E402

CASES = {"frame_96x80_smooth": (96, 80, "smooth", synth.SEED), "frame_70x45_noise": (70, 45, "noise", synth.SEED + 7)}


def main(out):
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "cases.txt"), "w") as f:
        f.write("\n".join(CASES) + "\n")
    for name, (W, H, kind, seed) in CASES.items():
        d, m, z = synth.frame(W, H, seed, kind)
        sky = synth.skybox(64, 48, seed)
        cdir = os.path.join(out, name)
        os.makedirs(cdir, exist_ok=True)
        d.tofile(os.path.join(cdir, "diffuse.raw")); m.tofile(os.path.join(cdir, "material.raw"))
        z.astype("<u2").tofile(os.path.join(cdir, "depth.raw")); sky.tofile(os.path.join(cdir, "skybox.raw"))
        open(os.path.join(cdir, "dims.txt"), "w").write("%d %d %d %d\n" % (W, H, sky.shape[1], sky.shape[0]))
    print("wrote", sorted(CASES), "under", out)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/tmp/b2_fixtures")
