"""Multi-GPU sharding logic.  CPU: planning arithmetic + a world_size-2 gloo run of the halo exchange.
GPU: row-banded compositing on one device reproduces the unbanded frame bit for bit."""
import os
import socket
import subprocess
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_instances_and_bands(b2):
    from dplug_b200 import sharding as sh
    assert [len(sh.instances_for_rank(256, r, 8)) for r in range(8)] == [32] * 8          # BASELINE config 4
    assert sum((sh.instances_for_rank(10, r, 4) for r in range(4)), []) == list(range(10))
    assert [sh.band_rows(16384, r, 8) for r in range(8)] == [(2048 * r, 2048 * (r + 1)) for r in range(8)]
    rows = [sh.band_rows(2160, r, 4) for r in range(4)]
    assert rows[0][0] == 0 and rows[-1][1] == 2160 and all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
    assert all(a % 64 == 0 for a, _ in rows)


def test_halo_extents_match_survey_appendix_c(b2):
    """SURVEY Appendix C (verified there by brute-force dependency propagation): band [2048,4096) of a 16384-row
    canvas needs 94 extra level-0 diffuse rows on each side (level-5 bicubic bloom) and 28 depth rows (AO pyramid)."""
    from dplug_b200 import sharding as sh
    assert sh.need_rows(16384, 16384, 2048, 4096, sh.DIFFUSE, 0) == (2048 - 94, 4096 + 94)
    assert sh.need_rows(16384, 16384, 2048, 4096, sh.DEPTH, 0) == (2048 - 28, 4096 + 28)
    assert sh.need_rows(16384, 16384, 2048, 4096, sh.MATERIAL, 0) == (2048, 4096)
    # the first band clamps at the top edge: shadow reach and bloom reach stop at row 0
    assert sh.need_rows(16384, 16384, 0, 2048, sh.DIFFUSE, 0) == (0, 2048 + 94)
    bands, plan = sh.halo_plan(16384, 16384, 8)
    assert plan[0] == [(sh.DIFFUSE, 1, 2048, 2142), (sh.DEPTH, 1, 2048, 2076)]
    assert len(plan[3]) == 4
    with pytest.raises(ValueError):
        sh.halo_plan(620, 330, 8)   # bands thinner than the bloom reach


def _brute_force_need(W, H, y0, y1):
    """independent dependency propagation for the diffuse pyramid: mark level-0 rows touched by band rows"""
    hs = [H >> l for l in range(6)]
    need = [set() for _ in range(6)]
    for j in range(y0, y1):
        for l in (1, 2, 3):
            y = max((j + 0.5) / 2 ** l - 0.5, 0.0)
            iy = min(int(y), hs[l] - 1)
            need[l].update([iy, min(iy + 1, hs[l] - 1)])
        for l in (4, 5):
            y = (j + 0.5) / 2 ** l - 0.5
            for k in (-1, 0, 1, 2):
                need[l].add(max(0, min(int(np.floor(y)) + k, hs[l] - 1)))
    need[0].update(range(y0, y1))
    for l in range(5, 0, -1):
        for r in need[l]:
            if l >= 2:
                need[l - 1].update([max(2 * r - 1, 0), 2 * r, 2 * r + 1, min(2 * r + 2, hs[l - 1] - 1)])
            else:
                need[l - 1].update([2 * r, 2 * r + 1])
    return [(min(s), max(s) + 1) for s in need]


@pytest.mark.parametrize("H,y0,y1", [(2160, 0, 576), (2160, 576, 1088), (2160, 1600, 2160), (1000, 320, 640), (777, 128, 777)])
def test_need_rows_vs_brute_force(b2, H, y0, y1):
    from dplug_b200 import sharding as sh
    W = 512
    want = _brute_force_need(W, H, y0, y1)
    for l in range(6):
        assert sh.need_rows(W, H, y0, y1, sh.DIFFUSE, l) == want[l], l


def test_interior_tile_rows(b2):
    """host arithmetic: the tiles a band may composite before its halos exist (bench.py --workload c5)"""
    from dplug_b200 import sharding as sh
    W = H = 16384
    bands, _ = sh.halo_plan(W, H, 8)
    for r, (y0, y1) in enumerate(bands):
        a, b = sh.interior_tile_rows(W, H, y0, y1)
        assert (a == y0) == (r == 0) and (b == y1) == (r == 7)
        assert 0 <= a - y0 <= 256 and 0 <= y1 - b <= 256 and (a - y0) % 64 == 0 and (y1 - b) % 64 == 0
        # nothing the interior reads is a halo row
        for which in (sh.DIFFUSE, sh.DEPTH):
            lo, hi = sh.need_rows(W, H, a, b, which, 0)
            assert y0 <= lo and hi <= y1
    # a band thinner than two bloom reaches has no interior
    a, b = sh.interior_tile_rows(256, 1024, 320, 448)
    assert a == b


_WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from dplug_b200 import sharding as sh, synth
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
W, H = 256, 1024
bands, plan = sh.halo_plan(W, H, world)
y0, y1 = bands[rank]
stores, full = {}, {}
for which, gen, elem in ((sh.DIFFUSE, synth.diffuse, 4), (sh.DEPTH, synth.depth, 2)):
    lo, hi = sh.need_rows(W, H, y0, y1, which, 0)
    buf = np.full((hi - lo, W * elem), 0xEE, dtype=np.uint8)          # stored rows, halos poisoned
    own = gen(W, H, y0=y0, y1=y1)
    buf[y0 - lo:y1 - lo] = own.reshape(y1 - y0, -1).view(np.uint8)
    stores[which] = (torch.from_numpy(buf), lo)
    full[which] = gen(W, H, y0=lo, y1=hi).reshape(hi - lo, -1).view(np.uint8)
n = sh.exchange_halos(rank, world, plan, stores, dist)
ok = all(np.array_equal(stores[k][0].numpy(), full[k]) for k in stores)
t = torch.tensor([1 if ok else 0]); dist.all_reduce(t, op=dist.ReduceOp.MIN)
# instance sharding: every instance index is owned exactly once
mine = torch.zeros(10, dtype=torch.int64); mine[sh.instances_for_rank(10, rank, world)] = 1
dist.all_reduce(mine)
if rank == 0:
    print("RESULT", int(t.item()), n, mine.tolist())
dist.destroy_process_group()
'''


def test_halo_exchange_gloo_world2(b2, tmp_path):
    """world_size 2 on CPU (gloo): after the exchange every rank holds exactly the generator's rows in its halos."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % {"root": ROOT})
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=240) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    line = [l for l in outs[0][0].splitlines() if l.startswith("RESULT")][0].split(None, 3)
    assert line[1] == "1" and int(line[2]) == 4 and line[3] == str([1] * 10)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("world", [2, 3])
def test_banded_composite_equals_unbanded(b2, mode, world):
    """Bands composited one after the other on one device, each seeing only its own rows + halos, reproduce the
    unbanded frame exactly (global coordinates: toEye, blue noise, edge clamps — SURVEY Appendix C)."""
    from dplug_b200 import sharding as sh, synth
    W, H = 384, 960
    d, m, z = synth.frame(W, H)
    sky = synth.skybox(128, 96)
    ref = b2.PBRCompositor(0)
    ref.resizeBuffers(W, H)
    ref.setMode(mode)
    ref.setSkybox(sky)
    ref.diffuseMap.upload(0, d); ref.materialMap.upload(0, m); ref.depthMap.upload(0, z)
    ref.regenerateMipmaps([(0, 0, W, H)])
    ref.compositeGUI([(0, 0, W, H)])
    want = ref.download()
    bands, plan = sh.halo_plan(W, H, world)
    for r, (y0, y1) in enumerate(bands):
        c = b2.PBRCompositor(0, band=(y0, y1))
        c.resizeBuffers(W, H)
        c.setMode(mode)
        c.setSkybox(sky)
        for which, img in ((sh.DIFFUSE, d), (sh.MATERIAL, m), (sh.DEPTH, z)):
            slo, shi, lo, hi = c.bandRows(which, 0)
            assert (lo, hi) == sh.need_rows(W, H, y0, y1, which, 0) and slo <= lo and hi <= shi
            c.map(which).uploadRows(0, img[lo:hi], lo)          # own rows + the halo rows a neighbour would send
        c.regenerateMipmaps([(0, 0, W, H)])
        c.compositeGUI([(0, y0, W, y1)])
        got = c.downloadRows(y0, y1)
        assert np.array_equal(got, want[y0:y1]), (r, y0, y1)
        # the mip rows the band generated are the unbanded ones
        for l in range(1, 6):
            _, _, lo, hi = c.bandRows(sh.DIFFUSE, l)
            assert np.array_equal(c.diffuseMap.downloadRows(l, lo, hi), ref.diffuseMap.download(l)[lo:hi]), l
        # the overlapped order of bench.py --workload c5: the band's own rows are regenerated while the halo rows are
        # still in flight (poisoned here), the halo rectangles afterwards
        c2 = b2.PBRCompositor(0, band=(y0, y1))
        c2.resizeBuffers(W, H)
        c2.setMode(mode)
        c2.setSkybox(sky)
        own, halos = sh.own_and_halo_rects(W, H, y0, y1)
        for which, img in ((sh.DIFFUSE, d), (sh.MATERIAL, m), (sh.DEPTH, z)):
            slo, shi, lo, hi = c2.bandRows(which, 0)
            poison = np.full_like(img[lo:hi], 0xAB)
            poison[y0 - lo:y1 - lo] = img[y0:y1]
            c2.map(which).uploadRows(0, poison, lo)
        c2.regenerateMipmaps(own)
        ia, ib = sh.interior_tile_rows(W, H, y0, y1)
        assert y0 <= ia <= ib <= y1 and (ia - y0) % 64 == 0 and (r != 0 or ia == y0) and (r != world - 1 or ib == y1 or ib == ia)
        if ib > ia:
            c2.compositeGUI([(0, ia, W, ib)])                    # interior tiles: before the halos exist
        for which, img in ((sh.DIFFUSE, d), (sh.MATERIAL, m), (sh.DEPTH, z)):
            slo, shi, lo, hi = c2.bandRows(which, 0)
            c2.map(which).uploadRows(0, img[lo:hi], lo)          # "the halos arrive"
        if halos:
            c2.regenerateMipmaps(halos)
        edge = [rc for rc in ((0, y0, W, ia), (0, ib, W, y1)) if rc[3] > rc[1]]
        if edge:
            c2.compositeGUI(edge)
        assert np.array_equal(c2.downloadRows(y0, y1), want[y0:y1]), ("overlapped", r, y0, y1)
        for l in range(1, 6):
            _, _, lo, hi = c2.bandRows(sh.DIFFUSE, l)
            assert np.array_equal(c2.diffuseMap.downloadRows(l, lo, hi), ref.diffuseMap.download(l)[lo:hi]), ("overlapped", l)


@pytest.mark.parametrize("W,H,world", [(16384, 16384, 8), (16384, 16384, 2), (3840, 2160, 4), (384, 960, 3), (256, 1024, 2), (1000, 777, 2)])
def test_library_band_plan_matches_python_planning(b2, W, H, world):
    """b2_band_rows / b2_band_plan (what b2band_composite_frame schedules by) against the Python restatement."""
    import ctypes as C
    from dplug_b200 import sharding as sh, _lib
    for r in range(world):
        y0, y1 = C.c_int(), C.c_int()
        assert _lib.lib().b2_band_rows(H, world, r, 64, C.byref(y0), C.byref(y1)) == 0
        assert (y0.value, y1.value) == sh.band_rows(H, r, world)
        if y1.value <= y0.value:
            continue
        halos, interior = sh.band_plan(W, H, y0.value, y1.value)
        assert halos == sh.own_and_halo_rects(W, H, y0.value, y1.value)[1]
        assert interior == sh.interior_tile_rows(W, H, y0.value, y1.value)


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 3])
def test_band_group_peers_frames_equal_unbanded(b2, world):
    """b2band (PEERS transport, every band on device 0 here; on a multi-GPU box spread over the devices): several
    frames with CHANGING content through b2band_composite_frame; every band's rows equal the unbanded frame bit for
    bit.  Halos start poisoned, so a frame that reads a row before it arrived (or a stale row of the previous frame)
    fails."""
    from dplug_b200 import sharding as sh, synth, _lib
    W, H = 384, 960
    ndev = _lib.lib().b2_device_count()
    sky = synth.skybox(128, 96)
    ref = b2.PBRCompositor(0); ref.resizeBuffers(W, H); ref.setSkybox(sky)
    bands = []
    for r in range(world):
        y0, y1 = sh.band_rows(H, r, world)
        c = b2.PBRCompositor(r % ndev, band=(y0, y1)); c.resizeBuffers(W, H); c.setSkybox(sky)
        for which in (sh.DIFFUSE, sh.DEPTH):
            slo, shi, lo, hi = c.bandRows(which, 0)
            elem = 4 if which == sh.DIFFUSE else 2
            poison = np.full((shi - slo, W, 4) if elem == 4 else (shi - slo, W), 0xAB, dtype=np.uint8 if elem == 4 else np.uint16)
            c.map(which).uploadRows(0, poison, slo)
        bands.append((c, y0, y1))
    grp = sh.BandGroup.peers([c for c, _, _ in bands])
    assert grp.haloBytesPerFrame() > 0 and grp.info(0)["n_recvs"] == 2 and grp.info(1)["n_sends"] == (2 if world == 2 else 4)
    for frame in range(3):
        d, m, z = synth.frame(W, H, synth.SEED + 17 * frame)
        ref.diffuseMap.upload(0, d); ref.materialMap.upload(0, m); ref.depthMap.upload(0, z)
        ref.regenerateMipmaps([(0, 0, W, H)]); ref.compositeGUI([(0, 0, W, H)])
        want = ref.download()
        for c, y0, y1 in bands:       # every band uploads ONLY its own rows
            c.diffuseMap.uploadRows(0, d[y0:y1], y0); c.materialMap.uploadRows(0, m[y0:y1], y0); c.depthMap.uploadRows(0, z[y0:y1], y0)
        grp.compositeFrame()
        grp.sync()
        for c, y0, y1 in bands:
            assert np.array_equal(c.downloadRows(y0, y1), want[y0:y1]), (frame, y0, y1)
            for l in range(1, 6):
                _, _, lo, hi = c.bandRows(sh.DIFFUSE, l)
                assert np.array_equal(c.diffuseMap.downloadRows(l, lo, hi), ref.diffuseMap.download(l)[lo:hi]), (frame, l)
    # the exchange alone (a host that schedules the frame itself)
    d, m, z = synth.frame(W, H, synth.SEED + 99)
    for c, y0, y1 in bands:
        c.diffuseMap.uploadRows(0, d[y0:y1], y0); c.depthMap.uploadRows(0, z[y0:y1], y0)
    grp.exchangeBegin(); grp.exchangeEnd(); grp.sync()
    for c, y0, y1 in bands:
        _, _, lo, hi = c.bandRows(sh.DIFFUSE, 0)
        assert np.array_equal(c.diffuseMap.downloadRows(0, lo, hi), d[lo:hi])
        _, _, lo, hi = c.bandRows(sh.DEPTH, 0)
        assert np.array_equal(c.depthMap.downloadRows(0, lo, hi), z[lo:hi])
    grp.close()


_NCCL_WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
import dplug_b200 as b2
from dplug_b200 import sharding as sh, synth
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("gloo", rank=rank, world_size=world)      # host channel for the 128-byte NCCL id only
W, H = 640, 1536
sky = synth.skybox(128, 96)
y0, y1 = sh.band_rows(H, rank, world)
c = b2.PBRCompositor(rank, band=(y0, y1)); c.resizeBuffers(W, H); c.setSkybox(sky)
ref = b2.PBRCompositor(rank); ref.resizeBuffers(W, H); ref.setSkybox(sky)
grp = sh.BandGroup.nccl(c, rank, world, H, dist)
ok = True
for frame in range(4):
    d, m, z = synth.frame(W, H, synth.SEED + 31 * frame)
    ref.diffuseMap.upload(0, d); ref.materialMap.upload(0, m); ref.depthMap.upload(0, z)
    ref.regenerateMipmaps([(0, 0, W, H)]); ref.compositeGUI([(0, 0, W, H)])
    want = ref.download()
    c.diffuseMap.uploadRows(0, d[y0:y1], y0); c.materialMap.uploadRows(0, m[y0:y1], y0); c.depthMap.uploadRows(0, z[y0:y1], y0)
    grp.compositeFrame(); grp.sync()
    ok = ok and np.array_equal(c.downloadRows(y0, y1), want[y0:y1])
t = torch.tensor([1 if ok else 0]); dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("RESULT", int(t.item()), grp.haloBytesPerFrame())
grp.close()
dist.destroy_process_group()
'''


@pytest.mark.gpu
def test_band_group_nccl_world2(b2, tmp_path):
    """Two processes, one GPU each, halo rows over NCCL (ncclSend / ncclRecv inside the library): four frames with
    changing content, each band equal to the unbanded frame.  Needs two GPUs."""
    from dplug_b200 import _lib
    if _lib.lib().b2_device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    script = tmp_path / "nccl_worker.py"
    script.write_text(_NCCL_WORKER % {"root": ROOT})
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=600) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    line = [l for l in outs[0][0].splitlines() if l.startswith("RESULT")][0].split()
    assert line[1] == "1" and int(line[2]) > 0
