"""Own-row mip chain of ONE band of the 16384^2 canvas (rank 3 of 8) in isolation on one GPU: what the chain costs
without the exchange, the halo pass and the lighting around it (b2band's debug phase time for it is 0.098 ms)."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
import dplug_b200 as b2
from dplug_b200 import sharding as sh, synth
W = H = 16384
N, r = 8, 3
y0, y1 = sh.band_rows(H, r, N)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
cs = []
for k in range(3):                       # three bands' worth of maps so that the inputs do not stay in L2
    c = b2.PBRCompositor(0, band=(y0, y1)); c.resizeBuffers(W, H); c.set_stream(stream.cuda_stream)
    blk = (synth.diffuse(W, 512, synth.SEED + k), synth.material(W, 512, synth.SEED + k), synth.depth(W, 512, synth.SEED + k))
    lo, hi = y0, y1                      # own rows only (the halo rows come from the neighbours in a real frame)
    for yy in range(lo, hi, 512):
        n = min(512, hi - yy)
        c.diffuseMap.uploadRows(0, blk[0][:n], yy); c.materialMap.uploadRows(0, blk[1][:n], yy); c.depthMap.uploadRows(0, blk[2][:n], yy)
    cs.append(c)
own = [(0, y0, W, y1)]
for i in range(6): cs[i % 3].regenerateMipmaps(own)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
K = 60
for i in range(K): cs[i % 3].regenerateMipmaps(own)
e1.record(stream); torch.cuda.synchronize()
print("own-row mip chain of one band (rows %d..%d): %.1f us" % (y0, y1, e0.elapsed_time(e1) / K * 1e3))
