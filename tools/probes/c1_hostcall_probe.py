"""Host-image call at the reference's default UI size (620x330): wall time per call vs the device-side trace."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import dplug_b200 as b2
from dplug_b200 import synth
W, H = 620, 330
torch.cuda.set_device(0)
d, m, z = synth.frame(W, H)
hosts = [torch.from_numpy(a).pin_memory().numpy() for a in (d, m, z)]
wfb = torch.empty((H, W, 4), dtype=torch.uint8).pin_memory().numpy()
c = b2.PBRCompositor(0); c.resizeBuffers(W, H, 64, 64); c.setSkybox(synth.skybox())
tiles = b2.BoxArray(b2.tileAreas([(0, 0, W, H)], 64, 64)); whole = b2.BoxArray([(0, 0, W, H)])
for _ in range(5):
    c.compositeTile(wfb, tiles, *hosts, updateRects=whole)
n = 300
t0 = time.perf_counter()
for _ in range(n):
    c.compositeTile(wfb, tiles, *hosts, updateRects=whole)
print("host call: %.1f us per frame" % ((time.perf_counter() - t0) / n * 1e6))
c.profileEnable(True)
c.compositeTile(wfb, tiles, *hosts, updateRects=whole)
tr = json.loads(c.traceJSON())
ev = tr["traceEvents"] if isinstance(tr, dict) else tr
for e in ev:
    if "dur" in e:
        print("  %-24s %8.1f us (at %.1f)" % (e["name"], e["dur"], e["ts"]))
