#!/usr/bin/env python
"""Benchmark of the one hot path: full PBR re-composite of a frame (GUIGraphics.regenerateMipmaps over the dirty area
+ ICompositor.compositeTile over its 64x64 tiles).  Metric: composited Mpix/s (BASELINE.json).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload auto|c1|c2|c3|c4|c5]

Default workload ("auto"):
  N = 1  BASELINE configs[1]: synthetic 3840x2160 maps, full re-composite on one B200 (device-resident `value`, host-image
         `e2e`); the other single-GPU configs (C1 620x330, C3 8192^2 pyramid, C5 canvas on one GPU) ride along in `extra`.
  N > 1  BASELINE configs[4]: ONE 16384x16384 canvas cut into N row bands, one band per GPU, level-0 halo rows exchanged
         per frame inside the library (b2band_*, NCCL send/recv over NVLink) — strong scaling.  Every rank checks its
         band against the unbanded frame computed on its own GPU (`parity`); the run fails if a byte differs.
         The N independent C2 replicas of round 1 are kept as `secondary.c2_replicas`.
Launched by torchrun for N > 1 (one rank per GPU).  Prints ONE JSON line on rank 0.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "composited Mpix/s (full PBR frame)"
_OUT = sys.stdout
WORKLOADS = {
    # name: (W, H, description)
    "c1": (620, 330, "C1 examples/distort default size 620x330, full-frame composite"),
    "c2": (3840, 2160, "C2 synthetic 3840x2160 diffuse/material/depth, full re-composite"),
    "c3": (8192, 8192, "C3 emissive Mipmap bloom pyramid (generateNextLevel, 8 levels) on 8192x8192 RGBA"),
    "c4": (2048, 1024, "C4 batch of 256 plug-in-instance UIs at 2048x1024, instances dealt out over the GPUs (one step = the whole batch)"),
    "c5": (16384, 16384, "C5 16384x16384 single canvas row-banded over the GPUs, level-0 halo exchange per frame"),
}
MIN_SUSTAINED_S = 1.0


def config_for(workload, n_gpus):
    """the `config` object — identical for the b200 arm and the reference arm (arm-specific detail lives elsewhere)"""
    W, H, desc = WORKLOADS[workload]
    return {"workload": desc, "width": W, "height": H, "tile": "64x64", "n_gpus": n_gpus,
            "call_shape": "regenerateMipmaps(dirty = whole frame) + compositeTile(64x64 tile list) per frame (device-resident: as one b2pbr_redraw call; host images: update_rects = NULL as the D shim passes)",
            "l2": "inputs larger than the 126 MB L2 (independent frames visited round-robin / a canvas of several hundred MB per GPU)"}


def algorithmic_bytes_frame(W, H):
    """every input read once + every API-visible output written once."""
    b = W * H * (4 + 4 + 2 + 4)
    b += 4 * sum((W >> l) * (H >> l) for l in range(1, 6))
    b += 2 * sum((W >> l) * (H >> l) for l in range(1, 5))
    return b


class ClockSampler:
    """nvidia-smi clock / throttle sampler running during the timed region (B200_PROFILING.md recipe), every 20 ms."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, period_ms=20):
        self.index, self.period_ms = index, period_ms
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", str(self.period_ms)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def mark(self):
        return time.perf_counter()

    def stop(self, t0=None, t1=None):
        """summary of the samples taken in [t0, t1] (host clock; None = all)"""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for ts, ln in self.lines:
            if (t0 is not None and ts < t0) or (t1 is not None and ts > t1):
                continue
            p = [c.strip() for c in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2])); pw.append(float(p[3]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples_under_load": len(sm), "power_w_max": max(pw) if pw else None, "period_ms": self.period_ms}


def _peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy)"
    except Exception:
        return 6650.0, "fallback 6650 GB/s (B200_PROFILING.md)"


def _sha1(path):
    try:
        return hashlib.sha1(open(path, "rb").read()).hexdigest()[:12]
    except Exception:
        return None


def _traffic(key):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/dram_traffic.json)
    with its provenance; `stale` = the kernel source changed since the capture."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "dram_traffic.json")))[key]
        src = _sha1(os.path.join(ROOT, "dplug_b200", "csrc", "pbr_tex.cu"))
        return int(t["dram_bytes_read"] + t["dram_bytes_write"]), {"file": "profiles/dram_traffic.json", "capture": t.get("source"),
                                                                   "kernel_source_sha1_at_capture": t.get("kernel_source_sha1"),
                                                                   "kernel_source_sha1_now": src,
                                                                   "stale": t.get("kernel_source_sha1") != src}
    except Exception:
        return None, {"file": None}


def _set_affinity(local):
    """pin this rank's host thread to the CPUs closest to its GPU (NUMA-local pinned buffers follow by first touch)"""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        cpus = sorted(os.sched_getaffinity(0))
        return {"cpus": "%d-%d (%d)" % (cpus[0], cpus[-1], len(cpus)) if cpus else None}
    except Exception as e:
        return {"cpus": None, "error": type(e).__name__}


# ------------------------------------------------------------------------------------------------ reference arm
def _oracle_sample(W, H, threads, steps, warmup, budget_s):
    """times the CPU restatement on a bounded sample of the workload: the top `rows` rows of the frame are the dirty
    rectangle of every step (regenerateMipmaps over it + compositeTile over its tiles) — proportional work, so the
    Mpix/s of the sample is the Mpix/s of the frame."""
    from dplug_b200 import synth
    from tests import oracle
    g = oracle.Graphics(W, H, numThreads=threads)
    g.setSkybox(synth.skybox())
    blk = min(H, 512)
    d, m, z = synth.frame(W, blk)
    reps = (min(H, 4096) + blk - 1) // blk                   # content for the rows the sample can touch
    import numpy as np
    rowsHave = min(H, reps * blk)
    g.setInputRows(np.concatenate([d] * reps)[:rowsHave], np.concatenate([m] * reps)[:rowsHave], np.concatenate([z] * reps)[:rowsHave])
    probe = [(0, 0, W, min(H, 128))]
    t0 = time.perf_counter()
    g.regenerateMipmaps(probe); g.compositeGUI(probe)
    t_probe = time.perf_counter() - t0
    per_row = t_probe / probe[0][3]
    rows = int(budget_s / max(1, steps + warmup) / per_row) // 64 * 64
    rows = max(64, min(rows, rowsHave if rowsHave < H else H))
    band = [(0, 0, W, rows)]
    for _ in range(warmup):
        g.regenerateMipmaps(band); g.compositeGUI(band)
    t0 = time.perf_counter()
    for _ in range(steps):
        g.regenerateMipmaps(band); g.compositeGUI(band)
    dt = time.perf_counter() - t0
    sample = "full frame per step" if rows == H else "rows [0,%d) of the %dx%d frame per step (regenerateMipmaps + compositeTile over them)" % (rows, W, H)
    return W * rows * steps / dt / 1e6, dt / steps * 1e3, sample


CPU_NOTE = ("C++ restatement of legacypbr.d / mipmap.d (oracle/, g++ -O3 -msse4.1, SSE intrinsics only where bit-exactness needs them, "
            "persistent thread pool like core/dplug/core/thread.d); the reference's LDC build with its SSE pass loops is not "
            "buildable here (no D compiler) and would be faster per thread")


def run_reference(args, workload):
    """The reference arm: the CPU restatement of Dplug's compositor (the reference is pure D and cannot be built in this
    image, so kind = "port"), all host threads, same metric / unit / config as the b200 arm.  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    W, H, desc = WORKLOADS[workload]
    cores = os.cpu_count() or 1
    value, ms, sample = _oracle_sample(W, H, cores, args.steps, args.warmup, 120.0)
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": "Mpix/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
           "higher_is_better": True, "scaling": "strong" if workload == "c5" else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": config_for(workload, args.gpus),
           "cpu_baseline": {"value": value, "unit": "Mpix/s", "cores": cores, "kind": "port", "sample": sample, "note": CPU_NOTE},
           "e2e": {"value": value, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), file=_OUT, flush=True)


def cpu_baseline(W, H, budget_s=20.0):
    """Oracle timed on the box's host cores on a bounded sample (rank 0, N = 1): all threads, and the reference's own
    3-thread cap (graphics.d:82-87) beside it."""
    cores = os.cpu_count() or 1
    v, _, sample = _oracle_sample(W, H, cores, 3, 1, budget_s * 0.6)
    v3, _, _ = _oracle_sample(W, H, 3, 2, 1, budget_s * 0.4)
    return {"value": v, "unit": "Mpix/s", "cores": cores, "kind": "port", "sample": sample, "note": CPU_NOTE, "faithful_3_threads_value": v3}


# ------------------------------------------------------------------------------------------------ b200 arm helpers
class _DevView:
    """zero-copy view of device memory for torch (cuda array interface)"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2, "strides": None}


def _dist_init(local):
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return world, rank, dist


def _max_over_ranks(x, dist):
    import torch
    if dist is None:
        return x
    t = torch.tensor([x], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _timed(stream, step, steps, barrier, min_seconds=0.0):
    """CUDA-event time of `steps` calls of step(i) on `stream` (ms); with min_seconds the loop is repeated in blocks of
    `steps` until that much device time has passed and the per-step mean over everything is returned."""
    import torch
    total_ms, total_steps, i = 0.0, 0, 0
    host0 = time.perf_counter()
    while True:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            step(i); i += 1
        e1.record(stream)
        torch.cuda.synchronize()
        total_ms += e0.elapsed_time(e1); total_steps += steps
        if total_ms * 1e-3 >= min_seconds or time.perf_counter() - host0 > 20 * max(min_seconds, 0.05):
            break
    return total_ms, total_steps


def frame_bench(local, W, H, instances, steps, warmup, graph=True, batch=False, batch_streams=4, rank=0, world=1, dist=None,
                e2e=True, sustained=True, mip_fusion=1, redraw=True):
    """device-resident + host-image timing of full-frame composites of WxH UIs on one GPU (C1 / C2 / C4 shapes)"""
    import numpy as np
    import torch
    import dplug_b200 as b2
    from dplug_b200 import synth, _lib
    _lib.lib().b2_set_mip_fusion(mip_fusion)
    stream = torch.cuda.Stream()
    if batch:
        from dplug_b200 import sharding
        R = len(sharding.instances_for_rank(256, rank, world))
    else:
        R = instances
    tiles = b2.BoxArray(b2.tileAreas([(0, 0, W, H)], 64, 64))   # marshalled once, like the D slice the caller keeps
    whole = b2.BoxArray([(0, 0, W, H)])
    sky = synth.skybox()
    comps, hosts, distinct = [], [], {}
    streams = [torch.cuda.Stream() for _ in range(batch_streams)] if batch else []
    for k in range(R):
        seed = synth.SEED + 1000 * rank + (k % 8 if batch else k)   # the batch re-uses 8 distinct synthetic UIs per rank
        if seed not in distinct:
            distinct[seed] = synth.frame(W, H, seed)
        d, m, z = distinct[seed]
        c = b2.PBRCompositor(local)
        c.resizeBuffers(W, H, 64, 64)
        c.set_stream((streams[k % len(streams)] if batch else stream).cuda_stream)
        c.setSkybox(sky)
        c.diffuseMap.upload(0, d); c.materialMap.upload(0, m); c.depthMap.upload(0, z)
        comps.append(c)
        if k == 0:
            hosts = [torch.from_numpy(a).pin_memory().numpy() for a in (d, m, z)]
    wfb = torch.empty((H, W, 4), dtype=torch.uint8).pin_memory().numpy()
    graphs = {}

    def one(i):
        c = comps[i % R]
        if graph and (i % R) in graphs:
            c.graphLaunch(graphs[i % R])     # the same two calls, replayed from a CUDA graph
            return
        # GUIGraphics.doDraw stages C + D: regenerateMipmaps (whole frame dirty) then compositeGUI -> compositeTile, as the
        # one call (b2pbr_redraw; --redraw 0: the two separate calls — same device work)
        if redraw:
            c.redraw(whole, tiles)
        else:
            c.regenerateMipmaps(whole)
            c.compositeTile(None, tiles)

    def step(i):
        if batch:
            fork = torch.cuda.Event(); fork.record(stream)
            for st in streams:
                st.wait_event(fork)
            for k in range(R):
                one(k)
            for st in streams:
                join = torch.cuda.Event(); join.record(st); stream.wait_event(join)
        else:
            one(i)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier(); torch.cuda.synchronize()

    W_ = max(3, warmup) if batch else max(3, warmup, R)
    for i in range(W_):
        step(i)
    barrier()
    if graph:
        for k, c in enumerate(comps):
            c.graphBegin()
            if redraw:
                c.redraw(whole, tiles)
            else:
                c.regenerateMipmaps(whole); c.compositeTile(None, tiles)
            graphs[k] = c.graphEnd()
        for i in range(1 if batch else R):
            step(i)
        barrier()
    launches0 = sum(c.kernelLaunches() for c in comps)
    # ---- burst: EXACTLY `steps` steps, barrier + synchronize on both sides
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for i in range(steps):
        step(i)
    e1.record(stream)
    barrier()
    ms = _max_over_ranks(e0.elapsed_time(e1), dist)
    launches = sum(c.kernelLaunches() for c in comps) - launches0
    res = {"ms_per_step": ms / steps, "launches": int(launches), "R": R, "frames_per_step": (256 if batch else world)}
    # ---- sustained: the same step looped for >= 1 s of device time, clocks sampled every 20 ms during it
    if sustained:
        sampler = ClockSampler(local).start() if rank == 0 else None
        time.sleep(0.15)
        barrier()
        h0 = time.perf_counter()
        tot_ms, tot_steps = _timed(stream, step, max(steps, 50 if not batch else 2), barrier, MIN_SUSTAINED_S)
        h1 = time.perf_counter()
        res["sustained"] = {"ms_per_step": _max_over_ranks(tot_ms / tot_steps, dist), "steps": tot_steps, "seconds": tot_ms * 1e-3}
        if sampler:
            res["clocks"] = sampler.stop(h0, h1)
    # ---- dominant kernel: per-launch duration with CUDA events on the launching stream, in the same frame sequence
    # (the mip regeneration of the same frame runs before it, untimed, so the cache state is the timed region's)
    evs = []
    for i in range(min(max(steps, 20), 200)):
        c = comps[i % R]
        st = streams[(i % R) % len(streams)] if batch else stream
        c.regenerateMipmaps(whole)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st); c.compositeTile(None, tiles); b.record(st)
        evs.append((a, b))
        if batch:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    res["light_ms"] = sum(a.elapsed_time(b) for a, b in evs) / len(evs)
    # ---- end to end through the reference-facing call with HOST buffers (pinned), called EXACTLY like the D shim
    # (d/dplug/gui/cudacompositor.d): 64x64 tile list, update_rects = NULL; upload of the three maps, device mips,
    # composite and the copy of the frame back to the host are all inside the timed region
    if e2e:
        d, m, z = hosts
        for _ in range(2):
            for k in range(R if batch else 1):
                comps[k].compositeTile(wfb, tiles, d, m, z)
        barrier()
        e2e_steps = max(3, min(steps, 10 if W * H > 2000000 else 200))
        per_step = R if batch else 1
        window = 4 if batch else 1   # the batch keeps a few UIs' calls in flight (begin / end); a single UI waits for its frame
        wfbs = [wfb] + [torch.empty((H, W, 4), dtype=torch.uint8).pin_memory().numpy() for _ in range(window - 1)]
        t0 = time.perf_counter()
        n_done = 0
        while True:
            for _ in range(e2e_steps):
                if not batch:
                    comps[0].compositeTile(wfb, tiles, d, m, z)   # synchronises on the copy-back
                    continue
                for k in range(per_step):
                    if k >= window:
                        comps[k - window].compositeTileEnd()
                    comps[k].compositeTileBegin(wfbs[k % window], tiles, d, m, z)
                for k in range(max(0, per_step - window), per_step):
                    comps[k].compositeTileEnd()
            n_done += e2e_steps
            if time.perf_counter() - t0 >= (0.5 if not batch else 0.0):
                break
        torch.cuda.synchronize()
        e2e_s = _max_over_ranks(time.perf_counter() - t0, dist)
        res["e2e"] = {"value": res["frames_per_step"] * W * H * n_done / e2e_s / 1e6, "unit": "Mpix/s", "ms_per_frame": e2e_s / n_done / per_step * 1e3,
                      "h2d_bytes_per_step": int(d.nbytes + m.nbytes + z.nbytes) * per_step, "d2h_bytes_per_step": int(wfb.nbytes) * per_step,
                      "call": ("b2pbr_composite_tile_host_begin / _end, 4 UIs in flight" if batch else "b2pbr_composite_tile_host") +
                              " (ICompositor.compositeTile with pinned host images; update_rects = NULL and the 64x64 tile list, as d/dplug/gui/cudacompositor.d calls it)"}
        if not batch:
            # a dirty-rectangle frame through the same call: one widget-sized rectangle (1/16 of the area) changed
            r = (W // 4, H // 4, W // 2, H // 2)
            grown = [b2.growRect(r, 20, W, H)]
            dtiles = b2.BoxArray(b2.tileAreas(grown, 64, 64))
            for _ in range(3):
                comps[0].compositeTile(wfb, dtiles, d, m, z)
            t0 = time.perf_counter(); n = 0
            while time.perf_counter() - t0 < 0.25:
                comps[0].compositeTile(wfb, dtiles, d, m, z); n += 1
            dt = time.perf_counter() - t0
            res["e2e_dirty_rect"] = {"ms_per_call": dt / n * 1e3, "rect": list(r), "composited_mpix_per_s": b2.boxArea(grown[0]) * n / dt / 1e6 if hasattr(b2, "boxArea") else
                                     (grown[0][2] - grown[0][0]) * (grown[0][3] - grown[0][1]) * n / dt / 1e6}
    del comps
    return res


def c3_bench(steps, warmup, mip_fusion=1, which=("random", "survey")):
    """BASELINE config 3: the diffuse (emissive) pyramid recipe on one 8192x8192 RGBA image: level 1
    boxAlphaCovIntoPremul, levels 2..8 cubic (graphics.d:1378-1392 with Mipmap(8))."""
    import numpy as np
    import torch
    import dplug_b200 as b2
    from dplug_b200 import _lib, synth
    W, H, _ = WORKLOADS["c3"]
    stream = torch.cuda.Stream()
    _lib.lib().b2_set_mip_fusion(mip_fusion)
    R = 2   # two images (2 x 358 MB) alternate so the source is not L2-resident
    B = 4 * W * H + 4 * sum((W >> l) * (H >> l) for l in range(1, 9))
    peak, src = _peak()
    out = {"algorithmic_bytes_per_step": B, "peak": peak, "peak_source": src}

    def measure(fill):
        mips = []
        for k in range(R):
            m = b2.Mipmap(b2.RGBA8, 8, W, H)
            m.set_stream(stream.cuda_stream)
            fill(m, k)
            mips.append(m)
        for i in range(max(3, warmup)):
            mips[i % R].generateChain(2, 2)
        torch.cuda.synchronize()
        l0 = sum(m.kernelLaunches() for m in mips)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(steps):
            mips[i % R].generateChain(2, 2)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        launches = int(sum(m.kernelLaunches() for m in mips) - l0)
        tot_ms, tot_steps = _timed(stream, lambda i: mips[i % R].generateChain(2, 2), max(steps, 100), None, 0.5)
        del mips
        return {"ms_per_step": ms, "sustained_ms_per_step": tot_ms / tot_steps, "gpu_launches": launches,
                "value": W * H / (ms * 1e-3) / 1e6, "unit": "L0-Mpix/s", "achieved_gbs": B / (ms * 1e-3) / 1e9, "frac": B / (ms * 1e-3) / 1e9 / peak}

    rng = np.random.default_rng(1)

    def fill_random(m, k):
        base = rng.integers(0, 256, size=(H // 8, W, 4), dtype=np.uint8)
        for b in range(8):
            m.uploadRows(0, np.roll(base, k * 7 + b, axis=0), b * (H // 8))

    def fill_survey(m, k):
        rows = 1024
        blk = synth.diffuse(W, rows, seed=synth.SEED + k)     # one 1024-row block of the SURVEY 8d generator repeated down the image
        for y0 in range(0, H, rows):
            m.uploadRows(0, blk, y0)

    if "survey" in which:
        out["survey_diffuse"] = dict(measure(fill_survey), texels="SURVEY 8d diffuse map (emissive alpha on ~2 % of the area)")
    if "random" in which:
        out["random"] = dict(measure(fill_random), texels="uniformly random RGBA (random alpha: worst case of the premultiplying box)")
    return out


# ------------------------------------------------------------------------------------------------ C5: row bands
def _c5_blocks(W, blk=512):
    from dplug_b200 import synth
    return [(synth.diffuse(W, blk, synth.SEED + k), synth.material(W, blk, synth.SEED + k), synth.depth(W, blk, synth.SEED + k)) for k in range(2)]


def _c5_upload(c, blocks, y0, y1, phase, blk=512):
    """canvas content independent of the banding: the 512-row block b of the canvas is generator block (b + phase) % 2"""
    yy = y0
    while yy < y1:
        b = yy // blk
        off = yy - b * blk
        n = min(blk - off, y1 - yy)
        d, m, z = blocks[(b + phase) % 2]
        c.diffuseMap.uploadRows(0, d[off:off + n], yy); c.materialMap.uploadRows(0, m[off:off + n], yy); c.depthMap.uploadRows(0, z[off:off + n], yy)
        yy += n


def c5_bench(local, W, H, steps, warmup, rank, world, dist, breakdown=False, e2e=True, verify=True, sustained=True):
    """one canvas in `world` row bands (strong scaling).  Returns the timing of this rank's band plus the parity check
    of its output rows against the unbanded frame computed on the same GPU."""
    import numpy as np
    import torch
    import dplug_b200 as b2
    from dplug_b200 import synth, sharding as sh
    stream = torch.cuda.Stream()
    y0, y1 = sh.band_rows(H, rank, world)
    sky = synth.skybox()
    c = b2.PBRCompositor(local, band=(y0, y1))
    c.resizeBuffers(W, H, 64, 64)
    c.set_stream(stream.cuda_stream)
    c.setSkybox(sky)
    blocks = _c5_blocks(W)
    _c5_upload(c, blocks, y0, y1, 0)
    grp = sh.BandGroup.nccl(c, rank, world, H, dist)

    def step(i=0):
        grp.compositeFrame()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier(); torch.cuda.synchronize()

    for _ in range(max(3, warmup)):
        step()
    barrier()
    l0 = c.kernelLaunches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    barrier()
    ms = _max_over_ranks(e0.elapsed_time(e1), dist)
    res = {"ms_per_step": ms / steps, "launches": int(c.kernelLaunches() - l0), "band": [y0, y1], "halo_bytes_per_frame_this_rank": grp.haloBytesPerFrame(),
           "interior_rows": [grp.info()["interior_lo"], grp.info()["interior_hi"]]}
    if sustained:
        sampler = ClockSampler(local).start() if rank == 0 else None
        time.sleep(0.15)
        barrier()
        h0 = time.perf_counter()
        # a fixed number of frames on every rank (the ranks exchange rows every frame: they must agree on the count)
        n = max(steps, int(MIN_SUSTAINED_S / max(1e-4, ms / steps * 1e-3)) + 1)
        tot_ms, tot_steps = _timed(stream, step, n, barrier, 0.0)
        barrier()
        h1 = time.perf_counter()
        res["sustained"] = {"ms_per_step": _max_over_ranks(tot_ms / tot_steps, dist), "steps": tot_steps, "seconds": tot_ms * 1e-3}
        if sampler:
            res["clocks"] = sampler.stop(h0, h1)
    # ---- parity: this band's rows against the unbanded frame of the same canvas, computed on this GPU; two contents
    # (the second one right after the first, so a halo row left over from the previous frame would show)
    if verify:
        ref = b2.PBRCompositor(local)
        ref.resizeBuffers(W, H, 64, 64)
        ref.set_stream(stream.cuda_stream)
        ref.setSkybox(sky)
        whole = b2.BoxArray([(0, 0, W, H)]); tiles = b2.BoxArray(b2.tileAreas([(0, 0, W, H)], 64, 64))
        ok, sums = True, []
        single_ms = None
        for phase in (1, 0):
            _c5_upload(ref, blocks, 0, H, phase)
            _c5_upload(c, blocks, y0, y1, phase)
            ref.regenerateMipmaps(whole); ref.compositeTile(None, tiles)
            step()
            barrier()
            outs = []
            for comp, lo, hi in ((c, y0, y1), (ref, 0, H)):
                info = comp.outputInfo()
                t = torch.as_tensor(_DevView(info["pixels"], info["pitch"] * (hi - lo)), device="cuda:%d" % local).view(hi - lo, info["pitch"])
                outs.append(t)
            band_rows, ref_rows = outs[0][:, :W * 4], outs[1][y0:y1, :W * 4]
            ok = ok and bool(torch.equal(band_rows, ref_rows))
            sums.append(int(band_rows.to(torch.int64).sum().item()))
        if rank == 0:
            # the same canvas on ONE GPU (this one), timed like the bands: the strong-scaling reference of this run
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for _ in range(2):
                ref.regenerateMipmaps(whole); ref.compositeTile(None, tiles)
            a.record(stream)
            for _ in range(5):
                ref.regenerateMipmaps(whole); ref.compositeTile(None, tiles)
            b.record(stream); torch.cuda.synchronize()
            single_ms = a.elapsed_time(b) / 5
        okAll = ok
        if dist is not None:
            t = torch.tensor([1 if ok else 0], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MIN); okAll = bool(t.item())
            s = torch.tensor(sums, device="cuda", dtype=torch.int64); dist.all_reduce(s); sums = [int(v) for v in s.tolist()]
        res["parity"] = {"bands_equal_unbanded_frame": okAll, "checked": "every output byte of every band vs the unbanded 16384x16384 frame on the same GPU, two contents back to back",
                         "canvas_byte_sums": sums}
        if single_ms is not None:
            res["single_gpu_ms"] = single_ms
        del ref
    # ---- end to end: host rows in (pinned), composited rows out, through the band objects
    if e2e:
        rows = y1 - y0
        hd = torch.empty((rows, W, 4), dtype=torch.uint8).pin_memory().numpy()
        hm = torch.empty((rows, W, 4), dtype=torch.uint8).pin_memory().numpy()
        hz = torch.empty((rows, W), dtype=torch.uint16).pin_memory().numpy()
        yy = 0
        while yy < rows:
            b = (y0 + yy) // 512; off = (y0 + yy) - b * 512; n = min(512 - off, rows - yy)
            d, m, z = blocks[b % 2]
            hd[yy:yy + n] = d[off:off + n]; hm[yy:yy + n] = m[off:off + n]; hz[yy:yy + n] = z[off:off + n]
            yy += n
        out = torch.empty((rows, W, 4), dtype=torch.uint8).pin_memory().numpy()

        def e2e_step():
            c.diffuseMap.uploadRows(0, hd, y0); c.materialMap.uploadRows(0, hm, y0); c.depthMap.uploadRows(0, hz, y0)
            grp.compositeFrame()
            c.downloadRowsInto(out, y0, y1)      # synchronises

        for _ in range(2):
            e2e_step()
        barrier()
        n = 5
        t0 = time.perf_counter()
        for _ in range(n):
            e2e_step()
        barrier()
        e2e_s = _max_over_ranks(time.perf_counter() - t0, dist)
        res["e2e"] = {"value": W * H * n / e2e_s / 1e6, "unit": "Mpix/s", "ms_per_frame": e2e_s / n * 1e3,
                      "h2d_bytes_per_step": int(hd.nbytes + hm.nbytes + hz.nbytes) * world if world > 1 else int(hd.nbytes + hm.nbytes + hz.nbytes),
                      "d2h_bytes_per_step": int(out.nbytes) * world if world > 1 else int(out.nbytes),
                      "call": "per band: Mipmap.upload of the band's level-0 rows (pinned host) + b2band_composite_frame + b2pbr_download of its rows",
                      "affinity": _set_affinity(local)}
    grp.close()
    del c
    return res


# ------------------------------------------------------------------------------------------------ drivers
def run_default_single(args):
    """N = 1: C2 headline + extras"""
    import torch
    W, H, desc = WORKLOADS["c2"]
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    _set_affinity(local)
    r = frame_bench(local, W, H, args.instances, args.steps, args.warmup, graph=bool(args.graph), mip_fusion=args.mip_fusion, redraw=bool(args.redraw))
    peak, peak_src = _peak()
    B = algorithmic_bytes_frame(W, H)
    achieved = B / (r["light_ms"] * 1e-3) / 1e9
    traffic, tsrc = _traffic("k_pbr_pair@%dx%d" % (W, H))
    out = {
        "metric": METRIC, "value": W * H / (r["ms_per_step"] * 1e-3) / 1e6, "unit": "Mpix/s", "n_gpus": 1,
        "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_for("c2", 1),
        "details": {"instances_per_gpu": r["R"], "cuda_graph": bool(args.graph),
                    "l2": "%d independent frames (%.0f MB) visited round-robin" % (r["R"], r["R"] * B / 1e6)},
        "sustained": dict(r.get("sustained", {}), value=W * H / (r["sustained"]["ms_per_step"] * 1e-3) / 1e6) if "sustained" in r else None,
        "frame_roofline_frac": (B / (r["ms_per_step"] * 1e-3) / 1e9) / peak,
        "roofline": {"bound": "hbm", "kernel": "k_pbr_pair (fused lighting, 8 passes, two rows per lane on the packed FP32 pipe; mip and skybox samplers on the texture units)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "traffic_source": tsrc, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": B, "launch_ms": r["light_ms"],
                     "note": "the lighting pass is FP32-issue-bound (433 instructions per pixel against 16 B per pixel): "
                             "the HBM fraction is the contract number, the issue utilisation is in profiles/"},
        "e2e": r["e2e"], "e2e_dirty_rect": r.get("e2e_dirty_rect"),
        "gpu_launches": r["launches"], "clocks": r.get("clocks"),
    }
    if args.extras:
        extra = {}
        w1, h1, _ = WORKLOADS["c1"]
        r1 = frame_bench(local, w1, h1, 4, 400, 3, graph=bool(args.graph), sustained=False)
        extra["c1"] = {"workload": WORKLOADS["c1"][2], "ms_per_step": r1["ms_per_step"], "value": w1 * h1 / (r1["ms_per_step"] * 1e-3) / 1e6,
                       "e2e": r1["e2e"], "gpu_launches_per_frame": r1["launches"] / 400.0, "lighting_ms": r1["light_ms"]}
        extra["c3"] = dict(c3_bench(50, 3, args.mip_fusion), workload=WORKLOADS["c3"][2])
        w5, h5, _ = WORKLOADS["c5"]
        r5 = c5_bench(local, w5, h5, 10, 3, 0, 1, None, e2e=False, verify=False, sustained=False)
        extra["c5_one_gpu"] = {"workload": WORKLOADS["c5"][2], "ms_per_step": r5["ms_per_step"], "value": w5 * h5 / (r5["ms_per_step"] * 1e-3) / 1e6}
        out["extra"] = extra
    if not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(W, H)
    print(json.dumps(out), file=_OUT, flush=True)


def run_default_multi(args):
    """N > 1: C5 headline (strong scaling over row bands) + the C2 replicas as a secondary figure"""
    import torch
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    world, rank, dist = _dist_init(local)
    aff = _set_affinity(local)
    W, H, desc = WORKLOADS["c5"]
    r = c5_bench(local, W, H, args.steps, args.warmup, rank, world, dist, breakdown=args.breakdown)
    sec = None
    if args.extras:
        w2, h2, _ = WORKLOADS["c2"]
        r2 = frame_bench(local, w2, h2, args.instances, max(args.steps, 20), args.warmup, graph=bool(args.graph), rank=rank, world=world, dist=dist, sustained=False)
        sec = {"c2_replicas": {"workload": WORKLOADS["c2"][2] + ", one independent instance stream per GPU (weak scaling, no collective)",
                               "value": world * w2 * h2 / (r2["ms_per_step"] * 1e-3) / 1e6, "ms_per_step": r2["ms_per_step"], "e2e": r2["e2e"]}}
    halo = torch.tensor([r["halo_bytes_per_frame_this_rank"]], device="cuda", dtype=torch.int64)
    dist.all_reduce(halo)
    if rank == 0:
        if not r["parity"]["bands_equal_unbanded_frame"]:
            print(json.dumps({"error": "C5 parity failed: a band's output differs from the unbanded frame", "parity": r["parity"]}), file=_OUT, flush=True)
            dist.destroy_process_group()
            sys.exit(1)
        peak, peak_src = _peak()
        B = algorithmic_bytes_frame(W, H)
        single = r.get("single_gpu_ms")
        out = {
            "metric": METRIC, "value": W * H / (r["ms_per_step"] * 1e-3) / 1e6, "unit": "Mpix/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic (two 512-row generator blocks alternating down the canvas; the same canvas for every N)",
            "config": config_for("c5", world),
            "details": {"parallelism": "%d row bands of whole 64-row tiles, one per GPU; per frame one exchange of level-0 halo rows (94 diffuse + 28 depth rows per band edge) "
                                       "by ncclSend/ncclRecv inside the library (b2band_composite_frame), hidden behind the interior tiles" % world,
                        "halo_bytes_per_frame": int(halo.item()), "interior_rows_rank0": r["interior_rows"], "affinity": aff},
            "sustained": dict(r["sustained"], value=W * H / (r["sustained"]["ms_per_step"] * 1e-3) / 1e6) if "sustained" in r else None,
            "parity": r["parity"],
            "strong_scaling_in_this_run": {"one_gpu_ms_same_canvas": single, "speedup": (single / r["ms_per_step"]) if single else None, "n_gpus": world},
            "frame_roofline_frac": (B / (r["ms_per_step"] * 1e-3) / 1e9) / (peak * world), "peak_source": peak_src,
            "roofline": {"bound": "hbm", "kernel": "whole banded frame (k_pbr_pair + mip chain + halo exchange)", "achieved": B / (r["ms_per_step"] * 1e-3) / 1e9,
                         "peak": peak * world, "unit": "GB/s", "frac": (B / (r["ms_per_step"] * 1e-3) / 1e9) / (peak * world), "traffic": None,
                         "algorithmic_bytes_per_launch": B},
            "e2e": r.get("e2e"), "secondary": sec,
            "gpu_launches": r["launches"], "clocks": r.get("clocks"),
        }
        print(json.dumps(out), file=_OUT, flush=True)
    dist.destroy_process_group()


def run_workload(args):
    """explicit --workload c1 | c2 | c3 | c4 | c5 (profiling and the per-config numbers quoted in DESIGN.md)"""
    import torch
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    wl = args.workload
    W, H, desc = WORKLOADS[wl]
    peak, peak_src = _peak()
    if wl == "c3":
        which = ("random", "survey") if args.c3_data == "both" else (args.c3_data,)
        r = c3_bench(args.steps, args.warmup, args.mip_fusion, which)
        head = r.get("random") or r.get("survey_diffuse")
        print(json.dumps({"metric": "pyramid L0-Mpix/s (8-level emissive Mipmap chain)", "value": head["value"], "unit": "Mpix/s", "n_gpus": 1,
                          "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": head["ms_per_step"], "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config_for("c3", 1),
                          "roofline": {"bound": "hbm", "kernel": "mip chain", "achieved": head["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": head["frac"],
                                       "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": r["algorithmic_bytes_per_step"]},
                          "variants": r, "gpu_launches": head["gpu_launches"]}), file=_OUT, flush=True)
        return
    world, rank, dist = _dist_init(local)
    if wl == "c5":
        r = c5_bench(local, W, H, args.steps, args.warmup, rank, world, dist, breakdown=args.breakdown, e2e=not args.no_e2e, verify=not args.no_verify)
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": W * H / (r["ms_per_step"] * 1e-3) / 1e6, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps,
                              "warmup": max(3, args.warmup), "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong",
                              "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_for("c5", world), "result": r,
                              "gpu_launches": r["launches"], "clocks": r.get("clocks")}), file=_OUT, flush=True)
    else:
        batch = wl == "c4"
        r = frame_bench(local, W, H, args.instances, args.steps, args.warmup, graph=bool(args.graph), batch=batch, batch_streams=args.batch_streams,
                        rank=rank, world=world, dist=dist, mip_fusion=args.mip_fusion, e2e=not args.no_e2e)
        if rank == 0:
            B = algorithmic_bytes_frame(W, H)
            achieved = B / (r["light_ms"] * 1e-3) / 1e9
            out = {"metric": METRIC, "value": r["frames_per_step"] * W * H / (r["ms_per_step"] * 1e-3) / 1e6, "unit": "Mpix/s", "n_gpus": world,
                   "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                   "scaling": "strong" if batch else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_for(wl, world),
                   "details": {"instances_per_gpu": r["R"], "cuda_graph": bool(args.graph)}, "sustained": r.get("sustained"),
                   "roofline": {"bound": "hbm", "kernel": "k_pbr_pair (fused lighting, 8 passes, two rows per lane on the packed FP32 pipe; mip and skybox samplers on the texture units)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                                "frac": achieved / peak, "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": B, "launch_ms": r["light_ms"]},
                   "e2e": r.get("e2e"), "gpu_launches": r["launches"], "clocks": r.get("clocks")}
            if world == 1 and not args.no_cpu_baseline and wl in ("c1", "c2"):
                out["cpu_baseline"] = cpu_baseline(W, H)
            print(json.dumps(out), file=_OUT, flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    # stdout carries the ONE JSON line and nothing else: libraries that print to fd 1 (NCCL's version banner) go to stderr
    global _OUT
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto"] + sorted(WORKLOADS))
    ap.add_argument("--instances", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--extras", type=int, default=1, help="auto workload: also measure the other configs (extra / secondary keys)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-verify", action="store_true", help="c5: skip the parity check against the unbanded frame")
    ap.add_argument("--batch-streams", type=int, default=4, help="c4: streams the batch's independent UIs are dealt over")
    ap.add_argument("--breakdown", action="store_true", help="c5: per-phase device times of a step on stderr (every rank)")
    ap.add_argument("--c3-data", default="both", choices=["both", "random", "survey"])
    ap.add_argument("--mip-fusion", type=int, default=1, choices=[0, 1, 2],
                    help="0: one launch per generateLevel call; 1 (default): small levels fused; 2: every level fused")
    ap.add_argument("--redraw", type=int, default=1, help="1: one b2pbr_redraw call per frame; 0: regenerateMipmaps + compositeTile (the same device work)")
    ap.add_argument("--graph", type=int, default=1, help="replay each frame's device work from a CUDA graph (1) or launch kernel by kernel (0)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    auto = args.workload == "auto"
    if args.impl == "reference":
        wl = ("c2" if max(world, args.gpus) == 1 else "c5") if auto else args.workload
        if wl in ("c3", "c4"):
            wl = "c2"
        run_reference(args, wl)
    elif auto and world == 1:
        run_default_single(args)
    elif auto:
        run_default_multi(args)
    else:
        run_workload(args)


if __name__ == "__main__":
    main()
