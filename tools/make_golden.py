#!/usr/bin/env python
"""Generates the committed golden fixtures under tests/golden/ from the NUMPY restatement (tools/np_reference.py, written
from the D sources independently of oracle/), so that the C++ oracle is checked against vectors it did not produce.

The reference (pure D) cannot be run in this environment: these vectors are not outputs of the reference itself (see
DESIGN.md, "parity status"); tools/d/dump_fixtures.d is the program a maintainer with LDC runs to replace them with
reference-held ones (same .npz keys, tools/d/README.md).

  python tools/make_golden.py            # rewrites tests/golden/*.npz
  python tools/make_golden.py --check    # verifies the committed files are what this script produces
"""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from dplug_b200 import synth  # noqa: E402
import np_reference as npr  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def frame_case(W, H, kind, seed):
    d, m, z = synth.frame(W, H, seed, kind)
    sky = synth.skybox(64, 48, seed)
    out = npr.composite(d, m, z, sky)
    D, Z = npr.diffuse_pyramid(d), npr.depth_pyramid(z)
    lv = {"diffuse_l%d" % l: D[l] for l in range(1, len(D))}
    lv.update({"depth_l%d" % l: Z[l] for l in range(1, len(Z))})
    return dict(diffuse=d, material=m, depth=z, skybox=sky, composited=out, **lv)


def mip_case():
    rng = np.random.default_rng(2024)
    rgba = rng.integers(0, 256, size=(37, 53, 4), dtype=np.uint8)
    l16 = rng.integers(0, 65536, size=(37, 53), dtype=np.uint16)
    return {"rgba": rgba, "l16": l16, "rgba_box": npr.box_rgba(rgba), "rgba_cubic": npr.cubic_rgba(rgba), "rgba_premul": npr.box_premul(rgba),
            "l16_box": npr.box_l16(l16), "l16_cubic": npr.cubic_l16(l16)}


CASES = {"frame_96x80_smooth.npz": lambda: frame_case(96, 80, "smooth", synth.SEED),
         "frame_70x45_noise.npz": lambda: frame_case(70, 45, "noise", synth.SEED + 7),
         "mips_53x37.npz": mip_case}


def main():
    os.makedirs(OUT, exist_ok=True)
    check = "--check" in sys.argv
    for name, make in CASES.items():
        data = make()
        path = os.path.join(OUT, name)
        if check:
            have = np.load(path)
            assert sorted(have.files) == sorted(data), (name, sorted(have.files), sorted(data))
            for k in data:
                assert np.array_equal(have[k], data[k]), (name, k)
        else:
            np.savez_compressed(path, **data)
    print("checked" if check else "wrote", sorted(CASES))


if __name__ == "__main__":
    main()
