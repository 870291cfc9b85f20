import os, sys
sys.path.insert(0, ".")
import numpy as np, torch
import dplug_b200 as b2
from dplug_b200 import synth
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
for (W, H) in [(4096, 4096), (8192, 1024), (2048, 2048), (1920, 1080)]:
    ms = []
    R = 8
    mips = []
    for k in range(R):
        m = b2.Mipmap(0, 1, W, H); m.set_stream(stream.cuda_stream)
        blk = synth.diffuse(W, min(H, 512), seed=synth.SEED + k)
        for y0 in range(0, H, blk.shape[0]): m.uploadRows(0, blk[: min(blk.shape[0], H - y0)], y0)
        mips.append(m)
    for i in range(8): mips[i % R].generateLevels(1, [1], (0, 0, W, H))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(80): mips[i % R].generateLevels(1, [1], (0, 0, W, H))
    e1.record(stream); torch.cuda.synchronize()
    print("%dx%d -> cubic: %.2f us" % (W, H, e0.elapsed_time(e1) / 80 * 1e3), flush=True)
    del mips
