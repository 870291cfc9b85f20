"""Times the host-image call (b2pbr_composite_tile_host) and its parts on the GPU box (profiling aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import dplug_b200 as b2
from dplug_b200 import synth
W, H = 3840, 2160
torch.cuda.set_device(0)
d, m, z = synth.frame(W, H)
hosts = [torch.from_numpy(a).pin_memory().numpy() for a in (d, m, z)]
wfb = torch.empty((H, W, 4), dtype=torch.uint8).pin_memory().numpy()
c = b2.PBRCompositor(0); c.resizeBuffers(W, H, 64, 64); c.setSkybox(synth.skybox())
tiles = b2.BoxArray(b2.tileAreas([(0, 0, W, H)], 64, 64)); whole = b2.BoxArray([(0, 0, W, H)])
for _ in range(3):
    c.compositeTile(wfb, tiles, *hosts, updateRects=whole)
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 10
for _ in range(n):
    c.compositeTile(wfb, tiles, *hosts, updateRects=whole)
torch.cuda.synchronize()
print("e2e ms per frame", (time.perf_counter() - t0) / n * 1e3)
# raw PCIe: the three maps up, the frame down, alone and together
td = [torch.empty(a.shape, dtype=torch.from_numpy(a).dtype, device="cuda") for a in hosts]
th = [torch.from_numpy(a) for a in hosts]
out_d = torch.empty((H, W, 4), dtype=torch.uint8, device="cuda"); out_h = torch.from_numpy(wfb)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def up():
    with torch.cuda.stream(s1):
        for a, b in zip(td, th): a.copy_(b, non_blocking=True)
def down():
    with torch.cuda.stream(s2):
        out_h.copy_(out_d, non_blocking=True)
for name, fn in (("H2D 83 MB", lambda: up()), ("D2H 33 MB", lambda: down()), ("both", lambda: (up(), down()))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10): fn()
    torch.cuda.synchronize()
    print(name, "ms", (time.perf_counter() - t0) / 10 * 1e3)
# chunked uploads (the pipeline's schedule) alone, then next to compositor kernels on another stream
bounds = [0, 1088, 1664, 1920, 2048, 2112, 2160]
def up_chunked():
    with torch.cuda.stream(s1):
        for y0, y1 in zip(bounds[:-1], bounds[1:]):
            for a, b in zip(td, th): a[y0:y1].copy_(b[y0:y1], non_blocking=True)
def timeit(name, fn, n=10):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    print(name, "ms", (time.perf_counter() - t0) / n * 1e3)
timeit("H2D chunked (18 copies)", up_chunked)
c.diffuseMap.upload(0, d); c.materialMap.upload(0, m); c.depthMap.upload(0, z)
def up_with_kernels():
    up_chunked()
    for _ in range(5):
        c.regenerateMipmaps(whole); c.compositeTile(None, tiles)
timeit("H2D chunked + 5 frames of kernels on the compositor stream", up_with_kernels)
def up_with_down():
    up_chunked(); down()
timeit("H2D chunked + D2H", up_with_down)
