"""The C++ restatement of the in-tree stb_image_resize port (oracle/ref_resizer.cpp) against a second restatement written
independently from the D source in numpy float32 (tools/np_resizer.py): byte for byte, no GPU."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import np_resizer as npz  # noqa: E402
from tests.test_resizer import KINDS, _src  # noqa: E402


@pytest.mark.parametrize("kind", sorted(KINDS))
def test_oracle_equals_numpy_restatement(oracle, kind):
    rng = np.random.default_rng(40 + kind)
    for (iw, ih), (ow, oh) in [((62, 33), (31, 17)), ((31, 17), (62, 33)), ((50, 21), (69, 12)), ((19, 46), (18, 100)),
                               ((64, 40), (48, 30)), ((40, 9), (7, 5)), ((40, 30), (40, 15)), ((9, 5), (40, 31)), ((33, 33), (33, 33))]:
        src = _src(rng, kind, iw, ih)
        got, want = npz.resize_image(kind, src, ow, oh), oracle.resizeImage(kind, src, ow, oh)
        assert np.array_equal(got, want), (KINDS[kind], (iw, ih), (ow, oh), int((got != want).sum()))
