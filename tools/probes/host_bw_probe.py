"""Host-memory / PCIe ceiling with N ranks copying at once (torchrun): every rank moves what one C5 band moves per frame
(level-0 rows up, composited rows down, pinned buffers) and reports its own rates; rank 0 prints the aggregate.
Explains the e2e figure of bench.py --gpus N: the 8 GPUs of a box share the host's memory controllers and PCIe roots."""
import os, time, torch, torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
W, rows = 16384, 16384 // world
up = [torch.empty((rows, W, 4), dtype=torch.uint8).pin_memory(), torch.empty((rows, W, 4), dtype=torch.uint8).pin_memory(), torch.empty((rows, W), dtype=torch.int16).pin_memory()]
dev = [t.cuda() for t in up]
down_d = torch.empty((rows, W, 4), dtype=torch.uint8, device="cuda"); down_h = torch.empty((rows, W, 4), dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def once(h2d=True, d2h=True):
    if h2d:
        with torch.cuda.stream(s1):
            for a, b in zip(dev, up): a.copy_(b, non_blocking=True)
    if d2h:
        with torch.cuda.stream(s2):
            down_h.copy_(down_d, non_blocking=True)
def timed(**kw):
    once(**kw); torch.cuda.synchronize()
    if world > 1: dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): once(**kw)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    if world > 1:
        t = torch.tensor([dt], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); dt = float(t.item())
    return dt
nb_up, nb_down = sum(t.numel() * t.element_size() for t in up), down_h.numel()
for name, kw, nb in (("H2D only", dict(d2h=False), nb_up), ("D2H only", dict(h2d=False), nb_down), ("both", {}, nb_up + nb_down)):
    dt = timed(**kw)
    if rank == 0:
        print("[host_bw] %d ranks, %-8s: %.2f ms per rank-frame, %.1f GB/s per rank, %.1f GB/s aggregate" % (world, name, dt * 1e3, nb / dt / 1e9, world * nb / dt / 1e9), flush=True)
if world > 1: dist.destroy_process_group()
