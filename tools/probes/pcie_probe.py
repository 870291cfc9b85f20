"""PCIe schedule probe (profiling aid): chunked H2D of the three maps with the finished rows going back concurrently."""
import sys, time
import torch
W, H = 3840, 2160
torch.cuda.set_device(0)
th = [torch.empty((H, W, 4), dtype=torch.uint8).pin_memory(), torch.empty((H, W, 4), dtype=torch.uint8).pin_memory(), torch.empty((H, W), dtype=torch.int16).pin_memory()]
td = [t.cuda() for t in th]
out_d = torch.empty((H, W, 4), dtype=torch.uint8, device="cuda"); out_h = torch.empty((H, W, 4), dtype=torch.uint8).pin_memory()
s1, s2, s3 = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
def run(bounds, lag=128, d2h=True, split_streams=False):
    evs = []
    for y0, y1 in zip(bounds[:-1], bounds[1:]):
        if split_streams:
            with torch.cuda.stream(s1):
                td[0][y0:y1].copy_(th[0][y0:y1], non_blocking=True)
                td[2][y0:y1].copy_(th[2][y0:y1], non_blocking=True)
            with torch.cuda.stream(s3):
                td[1][y0:y1].copy_(th[1][y0:y1], non_blocking=True)
                e3 = torch.cuda.Event(); e3.record(s3)
            s1.wait_event(e3)
        else:
            with torch.cuda.stream(s1):
                for a, b in zip(td, th): a[y0:y1].copy_(b[y0:y1], non_blocking=True)
        e = torch.cuda.Event(); e.record(s1); evs.append(e)
    if d2h:
        prev = 0
        for k, y1 in enumerate(bounds[1:]):
            done = H if y1 == H else max(0, (y1 - lag) // 64 * 64)
            if done > prev:
                s2.wait_event(evs[k])
                with torch.cuda.stream(s2):
                    out_h[prev:done].copy_(out_d[prev:done], non_blocking=True)
                prev = done
def timeit(name, fn, n=10):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    print("%-70s %.3f ms" % (name, (time.perf_counter() - t0) / n * 1e3))
def geo(minrows):
    b = [0]
    while b[-1] < H:
        left = H - b[-1]
        ch = max(minrows, ((left + 1) // 2 + 63) // 64 * 64)
        b.append(b[-1] + min(ch, left))
    return b
for name, b in (("1 chunk", [0, H]), ("geo min 64 (6 chunks)", geo(64)), ("geo min 256", geo(256)), ("2 chunks 1920+240", [0, 1920, H]),
                ("3 chunks 1280+640+240", [0, 1280, 1920, H]), ("equal 5 x 432", [0, 432, 864, 1296, 1728, H]), ("equal 9 x 240", list(range(0, H + 1, 240)))):
    timeit("H2D only, " + name + " " + str(len(b) - 1), lambda: run(b, d2h=False))
    timeit("H2D + pipelined D2H, " + name, lambda: run(b))
    timeit("H2D (2 streams) + pipelined D2H, " + name, lambda: run(b, split_streams=True))
