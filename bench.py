#!/usr/bin/env python
"""Benchmark of the one hot path: full PBR re-composite of a frame (regenerateMipmaps over the whole frame +
compositeTile over all 64x64 tiles), BASELINE.json configs[1]: synthetic 3840x2160 maps on 1 x B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c2|c1|c3]

N > 1 is launched by torchrun (one rank per GPU); frames are independent plug-in UIs, so ranks share nothing
(weak scaling, no data-path collective).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (W, H, description)
    "c1": (620, 330, "C1 examples/distort default size 620x330, full-frame composite"),
    "c2": (3840, 2160, "C2 synthetic 3840x2160 diffuse/material/depth, full re-composite"),
    "c3": (8192, 8192, "C3 emissive Mipmap bloom pyramid (generateNextLevel, 8 levels) on 8192x8192 RGBA"),
    "c4": (2048, 1024, "C4 batch of 256 plug-in-instance UIs at 2048x1024, instances dealt out over the GPUs (one step = the whole batch)"),
    "c5": (16384, 16384, "C5 16384x16384 single canvas row-banded over the GPUs, level-0 halo exchange per frame"),
}


def algorithmic_bytes_frame(W, H):
    """every input read once + every API-visible output written once."""
    b = W * H * (4 + 4 + 2 + 4)
    b += 4 * sum((W >> l) * (H >> l) for l in range(1, 6))
    b += 2 * sum((W >> l) * (H >> l) for l in range(1, 5))
    return b


class ClockSampler:
    """nvidia-smi clock / throttle sampler running during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            p = [c.strip() for c in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_reference(args, W, H, desc):
    """The reference arm: the CPU restatement of Dplug's compositor (the reference is pure D and cannot be built
    in this image, so kind = "port"), all host threads, on the same config and metric."""
    import numpy as np
    from dplug_b200 import synth
    from tests import oracle
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    d, m, z = synth.frame(W, H)
    g = oracle.Graphics(W, H, numThreads=cores)
    g.setSkybox(synth.skybox())
    g.setInputs(d, m, z)
    whole = [(0, 0, W, H)]
    # size the per-step sample: full frame if the whole run fits in ~2 minutes, else a band of tile rows
    t0 = time.perf_counter()
    g.regenerateMipmaps(whole)
    g.compositeGUI([(0, 0, W, min(H, 128))])
    t_probe = time.perf_counter() - t0
    est_frame = t_probe * max(1.0, H / 128.0)
    budget = 120.0 / max(1, args.steps + args.warmup)
    rows = H if est_frame <= budget else max(64, int(H * budget / est_frame) // 64 * 64)
    band = [(0, 0, W, rows)]
    for _ in range(args.warmup):
        g.regenerateMipmaps(whole); g.compositeGUI(band)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        g.regenerateMipmaps(whole); g.compositeGUI(band)
    dt = time.perf_counter() - t0
    value = W * rows * args.steps / dt / 1e6
    sample = "full frame" if rows == H else "mipmaps of the full frame + tiles of rows [0,%d) per step" % rows
    out = {"impl": "reference", "metric": "composited Mpix/s (full PBR frame)", "value": value, "unit": "Mpix/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": desc, "tile": "64x64", "threads": cores},
           "cpu_baseline": {"value": value, "unit": "Mpix/s", "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def cpu_baseline(W, H, budget_s=20.0):
    """Oracle timed on the box's host cores on a bounded sample (rank 0, N = 1)."""
    from dplug_b200 import synth
    from tests import oracle
    cores = os.cpu_count() or 1
    d, m, z = synth.frame(W, H)
    res = {}
    for label, threads in (("all", cores), ("faithful3", 3)):
        g = oracle.Graphics(W, H, numThreads=threads)
        g.setSkybox(synth.skybox())
        g.setInputs(d, m, z)
        whole = [(0, 0, W, H)]
        g.regenerateMipmaps(whole)
        rows = min(H, 256 if threads == 3 else H)
        band = [(0, 0, W, rows)]
        g.compositeGUI([(0, 0, W, 64)])  # warm-up
        t0 = time.perf_counter()
        n = 0
        while True:
            g.regenerateMipmaps(whole); g.compositeGUI(band)
            n += 1
            if time.perf_counter() - t0 > budget_s / 2 or n >= 3:
                break
        dt = time.perf_counter() - t0
        res[label] = (W * rows * n / dt / 1e6, threads, rows, n)
    v, threads, rows, n = res["all"]
    return {"value": v, "unit": "Mpix/s", "cores": threads, "kind": "port",
            "sample": "%d x (mipmaps of the full %dx%d frame + compositeTile over rows [0,%d)), %d threads" % (n, W, H, rows, threads),
            "faithful_3_threads_value": res["faithful3"][0]}


def run_b200(args, W, H, desc):
    import numpy as np
    import torch
    import dplug_b200 as b2
    from dplug_b200 import synth, _lib
    _lib.lib().b2_set_mip_fusion(args.mip_fusion)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    stream = torch.cuda.Stream()

    # R independent plug-in UIs per GPU, visited round-robin, so every step works on inputs that the previous
    # steps evicted from the 126 MB L2 (R * ~133 MB of maps)
    batch = args.workload == "c4"   # one step = every instance of the batch once (strong scaling: 256 UIs in total)
    if batch:
        from dplug_b200 import sharding
        R = len(sharding.instances_for_rank(256, rank, world))
    else:
        R = args.instances
    tiles = b2.BoxArray(b2.tileAreas([(0, 0, W, H)], 64, 64))   # marshalled once, like the D slice the caller keeps
    whole = b2.BoxArray([(0, 0, W, H)])
    sky = synth.skybox()
    comps, hosts = [], []
    distinct = {}
    streams = [torch.cuda.Stream() for _ in range(args.batch_streams)] if batch else []
    for k in range(R):
        seed = synth.SEED + 1000 * rank + (k % 8 if batch else k)   # the batch re-uses 8 distinct synthetic UIs per rank
        if seed not in distinct:
            distinct[seed] = synth.frame(W, H, seed)
        d, m, z = distinct[seed]
        c = b2.PBRCompositor(local)
        c.resizeBuffers(W, H, 64, 64)
        # the batch: independent UIs are dealt over a few streams, so one UI's small mip launches and the tail of its
        # lighting kernel overlap the next UI's work
        c.set_stream((streams[k % len(streams)] if batch else stream).cuda_stream)
        c.setSkybox(sky)
        c.diffuseMap.upload(0, d); c.materialMap.upload(0, m); c.depthMap.upload(0, z)
        comps.append(c)
        if k == 0:
            hosts = [torch.from_numpy(a).pin_memory().numpy() for a in (d, m, z)]
    wfb = torch.empty((H, W, 4), dtype=torch.uint8).pin_memory().numpy()

    graphs = {}

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
            one(i % R)

    def one(i):
        c = comps[i % R]
        if args.graph and (i % R) in graphs:
            c.graphLaunch(graphs[i % R])     # the same two calls, replayed from a CUDA graph
            return
        c.regenerateMipmaps(whole)       # GUIGraphics.regenerateMipmaps, whole frame dirty
        c.compositeTile(None, tiles)     # GUIGraphics.compositeGUI -> ICompositor.compositeTile

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for i in range(max(3, args.warmup) if batch else max(3, args.warmup, R)):
        step(i)
    barrier()
    if args.graph:
        for k, c in enumerate(comps):
            c.graphBegin()
            c.regenerateMipmaps(whole)
            c.compositeTile(None, tiles)
            graphs[k] = c.graphEnd()
        for i in range(1 if batch else R):
            step(i)
        barrier()
    launches0 = sum(c.kernelLaunches() for c in comps)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for i in range(args.steps):
        step(i)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    launches = sum(c.kernelLaunches() for c in comps) - launches0
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    # dominant kernel: the lighting kernel.  Per-launch duration with CUDA events on the launching stream, in the same
    # frame sequence (the mip regeneration of the same frame runs before it, untimed, so the cache state is the
    # timed region's)
    evs = []
    for i in range(min(max(args.steps, 20) if batch else args.steps, 200)):
        c = comps[i % R]
        st = streams[(i % R) % len(streams)] if batch else stream   # the stream this compositor launches on
        c.regenerateMipmaps(whole)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        c.compositeTile(None, tiles)
        b.record(st)
        evs.append((a, b))
        if batch:
            torch.cuda.synchronize()     # timed alone, not next to another stream's frame
    torch.cuda.synchronize()
    light_ms = sum(a.elapsed_time(b) for a, b in evs) / len(evs)

    # end-to-end through the reference-facing call with HOST buffers (pinned): upload the three maps, device
    # mips, composite, copy the frame back — all inside the timed region
    c = comps[0]
    d, m, z = hosts
    for _ in range(2):
        for k in range(R if batch else 1):
            comps[k].compositeTile(wfb, tiles, d, m, z, updateRects=whole)
    barrier()
    e2e_steps = max(3, min(args.steps, 10 if W * H > 2000000 else 200))   # small frames: enough calls to time
    per_step = R if batch else 1      # the batch: every instance of this rank goes through the host-image call
    window = 4 if batch else 1   # the batch keeps a few UIs' calls in flight (begin / end); a single UI waits for its frame
    wfbs = [wfb] + [torch.empty((H, W, 4), dtype=torch.uint8).pin_memory().numpy() for _ in range(window - 1)]
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        if not batch:
            comps[0].compositeTile(wfb, tiles, d, m, z, updateRects=whole)   # synchronises on the copy-back
            continue
        for k in range(per_step):
            if k >= window:
                comps[k - window].compositeTileEnd()
            comps[k].compositeTileBegin(wfbs[k % window], tiles, d, m, z, updateRects=whole)
        for k in range(max(0, per_step - window), per_step):
            comps[k].compositeTileEnd()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    B = algorithmic_bytes_frame(W, H)
    achieved = B / (light_ms * 1e-3) / 1e9 if light_ms > 0 else 0.0
    frames_per_step = 256 if batch else world   # whole job, all ranks
    value = frames_per_step * W * H * args.steps / (ms * 1e-3) / 1e6
    out = {
        "metric": "composited Mpix/s (full PBR frame)", "value": value, "unit": "Mpix/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong" if batch else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic (8 distinct UIs per GPU, re-used across the batch)" if batch else "synthetic",
        "config": {"workload": desc, "tile": "64x64", "instances_per_gpu": R, "cuda_graph": bool(args.graph), **({"streams": len(streams)} if batch else {}),
                   "l2": "inputs larger than L2: %d independent frames (%.0f MB) visited round-robin" % (R, R * B / 1e6),
                   "parallelism": "independent frames per GPU, no collective" if world > 1 else "1 GPU"},
        "frame_roofline_frac": (B * (R if batch else 1) * args.steps / (ms * 1e-3) / 1e9) / peak,
        "roofline": {"bound": "hbm", "kernel": "k_pbr_fast (fused lighting, 8 passes)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": _traffic("k_pbr_fast@%dx%d" % (W, H)), "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": B, "launch_ms": light_ms},
        "e2e": {"value": frames_per_step * W * H * e2e_steps / e2e_s / 1e6, "unit": "Mpix/s",
                "h2d_bytes_per_step": int(d.nbytes + m.nbytes + z.nbytes) * per_step, "d2h_bytes_per_step": int(wfb.nbytes) * per_step,
                "call": "b2pbr_composite_tile_host_begin / _end, 4 UIs in flight (host images, pinned)" if batch else
                        "b2pbr_composite_tile_host (ICompositor.compositeTile with host images, pinned)"},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(W, H)
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


class _DevView:
    """zero-copy view of device memory for torch (cuda array interface)"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2, "strides": None}


def _traffic(key):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/), or None."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "dram_traffic.json")))[key]
        return int(t["dram_bytes_read"] + t["dram_bytes_write"])
    except Exception:
        return None


def _peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy)"
    except Exception:
        return 6650.0, "fallback 6650 GB/s (B200_PROFILING.md)"


def run_c3(args, W, H, desc):
    """BASELINE config 3: the diffuse (emissive) pyramid recipe on one 8192x8192 RGBA image: level 1
    boxAlphaCovIntoPremul, levels 2..8 cubic (graphics.d:1378-1392 with Mipmap(8)).  Headline = uniformly random
    texels (alpha random: the worst case for the alpha-premultiplying box, no texel takes the alpha == 0 shortcut);
    `variants.survey_diffuse` = the SURVEY section 8d diffuse map (emissive alpha on ~2 % of the area)."""
    import numpy as np
    import torch
    import dplug_b200 as b2
    from dplug_b200 import _lib, synth
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream()
    _lib.lib().b2_set_mip_fusion(args.mip_fusion)
    R = 2   # two images (2 x 358 MB) alternate so the source is not L2-resident
    B = 4 * W * H + 4 * sum((W >> l) * (H >> l) for l in range(1, 9))
    peak, src = _peak()

    def measure(fill):
        mips = []
        for k in range(R):
            m = b2.Mipmap(b2.RGBA8, 8, W, H)
            m.set_stream(stream.cuda_stream)
            fill(m, k)
            mips.append(m)
        for i in range(max(3, args.warmup)):
            mips[i % R].generateChain(2, 2)
        torch.cuda.synchronize()
        l0 = sum(m.kernelLaunches() for m in mips)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler = ClockSampler(0); sampler.start(); time.sleep(0.3)
        e0.record(stream)
        for i in range(args.steps):
            mips[i % R].generateChain(2, 2)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop()
        launches = int(sum(m.kernelLaunches() for m in mips) - l0)
        del mips
        return ms, clocks, launches

    rng = np.random.default_rng(1)
    base = rng.integers(0, 256, size=(H // 8, W, 4), dtype=np.uint8)

    def fill_random(m, k):
        for b in range(8):
            m.uploadRows(0, np.roll(base, k * 7 + b, axis=0), b * (H // 8))

    def fill_survey(m, k):
        rows = 512
        for y0 in range(0, H, rows):
            m.uploadRows(0, synth.diffuse(W, H, seed=synth.SEED + k, y0=y0, y1=min(H, y0 + rows)), y0)

    if args.c3_data == "survey":   # profiling aid: only the SURVEY data set
        ms, clocks, launches = measure(fill_survey)
        print(json.dumps({"workload": "c3", "texels": "survey", "ms_per_step": ms / args.steps, "gpu_launches": launches}), flush=True)
        return
    ms, clocks, launches = measure(fill_random)
    ms2 = ms if args.c3_data == "random" else measure(fill_survey)[0]
    achieved = B * args.steps / (ms * 1e-3) / 1e9
    achieved2 = B * args.steps / (ms2 * 1e-3) / 1e9
    fused = args.mip_fusion
    print(json.dumps({
        "metric": "pyramid L0-Mpix/s (8-level emissive Mipmap chain)", "value": W * H * args.steps / (ms * 1e-3) / 1e6, "unit": "Mpix/s",
        "n_gpus": 1, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": desc, "texels": "uniformly random RGBA (random alpha: worst case of the premultiplying box)",
                   "mip_fusion": fused, "l2": "inputs larger than L2: %d images alternate (%.0f MB each)" % (R, B / 1e6)},
        "roofline": {"bound": "hbm", "kernel": {0: "mip chain (k_box_rgba<premul> + 7 x k_cubic_rgba)", 1: "mip chain (k_box_rgba<premul>, 2 x k_cubic_rgba, k_mip_fused levels 4-6 and 7-8)", 2: "k_mip_fused x 3 (levels 1-3, 4-6, 7-8)"}[fused],
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": src,
                     "algorithmic_bytes_per_launch": B},
        "variants": {"survey_diffuse": {"texels": "SURVEY 8d diffuse map (emissive alpha on ~2 % of the area)", "ms_per_step": ms2 / args.steps,
                                        "value": W * H * args.steps / (ms2 * 1e-3) / 1e6, "achieved": achieved2, "frac": achieved2 / peak}},
        "gpu_launches": launches, "clocks": clocks}), flush=True)


def run_c5(args, W, H, desc):
    """BASELINE config 5: one canvas cut into row bands, one band per GPU (strong scaling).  Per frame: level-0
    halo rows are exchanged with the neighbouring bands (NCCL point-to-point over NVLink), every band regenerates
    the mip rows it reads and composites its tiles."""
    import numpy as np
    import torch
    import dplug_b200 as b2
    from dplug_b200 import synth, sharding as sh
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    stream = torch.cuda.current_stream()
    bands, plan = sh.halo_plan(W, H, world)
    y0, y1 = bands[rank]
    c = b2.PBRCompositor(local, band=(y0, y1) if world > 1 else None)
    c.resizeBuffers(W, H, 64, 64)
    c.set_stream(stream.cuda_stream)
    c.setSkybox(synth.skybox())
    # synthetic content: one 512-row block of the generator repeated down the band (generating 268 Mpix with
    # numpy would take minutes); every rank uploads ONLY its own rows, halos arrive through the exchange
    blk = 512
    d, m, z = (synth.diffuse(W, blk, synth.SEED + rank), synth.material(W, blk, synth.SEED + rank), synth.depth(W, blk, synth.SEED + rank))
    for yy in range(y0, y1, blk):
        n = min(blk, y1 - yy)
        c.diffuseMap.uploadRows(0, d[:n], yy); c.materialMap.uploadRows(0, m[:n], yy); c.depthMap.uploadRows(0, z[:n], yy)
    stores = {}
    for which in (sh.DIFFUSE, sh.DEPTH):
        w_, h_, pitch, ptr, first = c.map(which).levelInfo(0)
        slo, shi, _, _ = c.bandRows(which, 0)
        t = torch.as_tensor(_DevView(ptr, pitch * (shi - slo)), device="cuda:%d" % local).view(shi - slo, pitch)
        stores[which] = (t, first)
    tiles = b2.BoxArray(b2.tileAreas([(0, y0, W, y1)], 64, 64))
    whole = b2.BoxArray([(0, 0, W, H)])

    own_rects, halo_rects = sh.own_and_halo_rects(W, H, y0, y1)
    ia, ib = sh.interior_tile_rows(W, H, y0, y1)
    own_rects, halo_rects = b2.BoxArray(own_rects), (b2.BoxArray(halo_rects) if halo_rects else None)
    interior = b2.BoxArray(b2.tileAreas([(0, ia, W, ib)], 64, 64)) if ib > ia else None
    edge_rects = [r for r in ((0, y0, W, ia), (0, ib, W, y1)) if r[3] > r[1]]
    boundary = b2.BoxArray(b2.tileAreas(edge_rects, 64, 64)) if edge_rects else None
    side = torch.cuda.Stream()
    marks = []   # --breakdown: events between the phases of a step

    def mark():
        if args.breakdown:
            e = torch.cuda.Event(enable_timing=True); e.record(stream); marks.append(e)

    def step():
        mark()
        if world > 1:
            # One exchange step per frame, hidden behind the work that does not depend on it: the halo rows travel while
            # the mip rows that depend only on the band's own rows are generated and the interior tiles (those reading
            # no halo row and no mip texel the halo pass writes) are composited; the mip texels that depend on a halo
            # row are (re)generated on a second stream once the rows are there, then the boundary tiles follow.
            reqs = sh.post_halo_exchange(rank, world, plan, stores, dist)
            c.regenerateMipmaps(own_rects)
            own_done = torch.cuda.Event(); own_done.record(stream)
            mark()
            if interior is not None:
                c.compositeTile(None, interior)
            mark()
            with torch.cuda.stream(side):
                side.wait_event(own_done)
                sh.wait_halo_exchange(reqs)
                if halo_rects is not None:
                    c.regenerateMipmaps(halo_rects, stream=side.cuda_stream)
                halo_done = torch.cuda.Event(); halo_done.record(side)
            stream.wait_event(halo_done)
            if boundary is not None:
                c.compositeTile(None, boundary)
        else:
            c.regenerateMipmaps(whole)
            mark()
            c.compositeTile(None, tiles)
        mark()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(); torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    l0 = c.kernelLaunches()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start(); time.sleep(0.3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    if args.breakdown:
        per = len(marks) // (args.steps + max(3, args.warmup))
        last = marks[-per * args.steps:]
        names = ["post exchange + mips (own rows)", "lighting (interior tiles)", "halo pass tail + lighting (boundary tiles)"] if world > 1 else ["mips", "lighting"]
        acc = [0.0] * (per - 1)
        for k in range(args.steps):
            for j in range(per - 1):
                acc[j] += last[k * per + j].elapsed_time(last[k * per + j + 1])
        print("rank %d breakdown ms/step: %s" % (rank, ", ".join("%s %.3f" % (n, a / args.steps) for n, a in zip(names, acc))), file=sys.stderr, flush=True)
    if world > 1:
        t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
    if rank == 0:
        B = algorithmic_bytes_frame(W, H)
        peak, src = _peak()
        halo = sum((hi - lo) * W * (4 if which == sh.DIFFUSE else 2) for r in plan for (which, _, lo, hi) in plan[r])
        print(json.dumps({
            "metric": "composited Mpix/s (full PBR frame)", "value": W * H * args.steps / (ms * 1e-3) / 1e6, "unit": "Mpix/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic (512-row generator block repeated)",
            "config": {"workload": desc, "tile": "64x64", "parallelism": "%d row bands, level-0 halo exchange (%d bytes per frame in total)" % (world, halo),
                       "l2": "inputs larger than L2 (%.0f MB per band)" % (B / world / 1e6)},
            "frame_roofline_frac": (B * args.steps / (ms * 1e-3) / 1e9) / (peak * world), "peak_source": src,
            "gpu_launches": int(c.kernelLaunches() - l0), "clocks": clocks}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--instances", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--batch-streams", type=int, default=4, help="c4: streams the batch's independent UIs are dealt over")
    ap.add_argument("--breakdown", action="store_true", help="c5: per-phase device times of a step on stderr (every rank)")
    ap.add_argument("--c3-data", default="both", choices=["both", "random", "survey"])
    ap.add_argument("--mip-fusion", type=int, default=1, choices=[0, 1, 2],
                    help="0: one launch per generateLevel call; 1 (default): small levels fused; 2: every level fused")
    ap.add_argument("--graph", type=int, default=1, help="replay each frame's device work from a CUDA graph (1) or launch kernel by kernel (0)")
    args = ap.parse_args()
    W, H, desc = WORKLOADS[args.workload]
    if args.impl == "reference":
        if args.workload in ("c3", "c5"):
            W, H, desc = WORKLOADS["c2"]
        run_reference(args, W, H, desc)
    elif args.workload == "c3":
        run_c3(args, W, H, desc)
    elif args.workload == "c5":
        run_c5(args, W, H, desc)
    else:
        run_b200(args, W, H, desc)


if __name__ == "__main__":
    main()
