"""Per-phase cost of a row band on ONE GPU (profiling aid): the 16384^2 canvas as N band objects on device 0 (PEERS
transport, device-to-device halo copies); every frame runs the N bands one after the other.  With B2_BAND_DEBUG=1 the
library prints each band's phase times; the total per frame / N is a band's busy time without any cross-GPU overlap."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import dplug_b200 as b2
from dplug_b200 import sharding as sh, synth
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
W = H = 16384
sky = synth.skybox()
blk = (synth.diffuse(W, 512), synth.material(W, 512), synth.depth(W, 512))
bands = []
for r in range(N):
    y0, y1 = sh.band_rows(H, r, N)
    c = b2.PBRCompositor(0, band=(y0, y1)); c.resizeBuffers(W, H); c.setSkybox(sky)
    for yy in range(y0, y1, 512):
        n = min(512, y1 - yy)
        c.diffuseMap.uploadRows(0, blk[0][:n], yy); c.materialMap.uploadRows(0, blk[1][:n], yy); c.depthMap.uploadRows(0, blk[2][:n], yy)
    bands.append(c)
g = sh.BandGroup.peers(bands)
for _ in range(3):
    g.compositeFrame(); g.sync()
t0 = time.perf_counter()
K = 10
for _ in range(K):
    g.compositeFrame(); g.sync()          # synchronised per frame so that the debug events are read
dt = (time.perf_counter() - t0) / K
print("bands=%d: %.3f ms per frame of all bands on one GPU = %.3f ms per band" % (N, dt * 1e3, dt * 1e3 / N))
g.close()
