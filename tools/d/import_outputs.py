#!/usr/bin/env python
"""Turns the raw files written by tools/d/dump_fixtures.d (the REAL PBRCompositor / Mipmap) into tests/golden/*.npz —
the same keys tools/make_golden.py writes — and reports how they compare with the numpy restatement."""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import np_reference as npr  # noqa: E402


def main(src):
    for name in open(os.path.join(src, "cases.txt")).read().split():
        cdir = os.path.join(src, name)
        W, H, SW, SH = [int(v) for v in open(os.path.join(cdir, "dims.txt")).read().split()]
        rd = lambda f, shape, dt=np.uint8: np.fromfile(os.path.join(cdir, f), dtype=dt).reshape(shape)
        data = {"diffuse": rd("diffuse.raw", (H, W, 4)), "material": rd("material.raw", (H, W, 4)), "depth": rd("depth.raw", (H, W), "<u2"),
                "skybox": rd("skybox.raw", (SH, SW, 4)), "composited": rd("composited.raw", (H, W, 4))}
        for l in range(1, npr.num_levels(5, W, H)):
            data["diffuse_l%d" % l] = rd("diffuse_l%d.raw" % l, (H >> l, W >> l, 4))
        for l in range(1, npr.num_levels(4, W, H)):
            data["depth_l%d" % l] = rd("depth_l%d.raw" % l, (H >> l, W >> l), "<u2")
        mine = npr.composite(data["diffuse"], data["material"], data["depth"], data["skybox"])
        diff = np.abs(mine.astype(np.int16) - data["composited"].astype(np.int16))
        print("%s: reference vs numpy restatement: max |diff| %d LSB on %d channels" % (name, diff.max(), int((diff > 0).sum())))
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), **data)
        # SURVEY 8(f) rows 2 and 4 against the CPU restatement (reported, not stored: no golden reads them yet)
        sys.path.insert(0, ROOT)
        from tests import oracle
        qb_path = os.path.join(cdir, "screenshot.qb")
        if os.path.exists(qb_path):
            rendered = data["composited"].copy(); rendered[..., 3] = 255
            same = open(qb_path, "rb").read() == oracle.encodeScreenshotAsQB(rendered, 0, data["depth"])
            print("%s: encodeScreenshotAsQB reference == restatement: %s" % (name, same))
        for k, (num, den) in enumerate(((1, 2), (3, 2))):
            ow, oh = W * num // den, H * num // den
            for what, kind, src, shape, dt in (("diffuse", 0, data["diffuse"], (oh, ow, 4), np.uint8), ("material", 1, data["material"], (oh, ow, 4), np.uint8),
                                               ("depth", 2, data["depth"], (oh, ow), "<u2")):
                f = os.path.join(cdir, "resized%d_%s.raw" % (k, what))
                if os.path.exists(f):
                    ref = np.fromfile(f, dtype=dt).reshape(shape)
                    d = np.abs(ref.astype(np.int64) - oracle.resizeImage(kind, src, ow, oh).astype(np.int64))
                    print("%s: resizeImage %s x%d/%d reference vs in-tree-port restatement: max |diff| %d on %d samples "
                          "(0 expected only when dplug:graphics was built with DPLUG_USE_STB_IMAGE_RESIZE_V2 = false)" % (name, what, num, den, d.max(), int((d > 0).sum())))
    print("tests/golden rewritten from the reference's output: run pytest tests/test_golden.py tests/test_np_reference.py")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/tmp/b2_fixtures")
