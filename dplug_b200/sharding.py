"""Multi-GPU sharding of the path.

* batch of plug-in UIs (BASELINE config 4): instances are independent -> contiguous blocks, no communication;
* one large canvas (config 5): horizontal row bands; the only exchange is the level-0 halo rows (diffuse and
  depth) between neighbouring bands, once per frame.  The exchange and the overlapped frame schedule live in the
  library behind the C ABI (`b2band_*`, include/b2pbr.h; `BandGroup` below is its ctypes wrapper): NCCL send/recv
  for one process per GPU, copy-engine peer copies for one process driving several GPUs.
  The planning functions below (`halo_plan`, `own_and_halo_rects`, `interior_tile_rows`) restate the library's
  planning arithmetic in Python; the tests check the two against each other, and `exchange_halos` is the same
  exchange over torch.distributed (gloo) for the world-size-2 CPU test.
"""
import ctypes as C
from . import _lib

DIFFUSE, MATERIAL, DEPTH = 0, 1, 2


def instances_for_rank(n_instances, rank, world):
    """contiguous block of instance indices owned by `rank`"""
    base, extra = divmod(n_instances, world)
    lo = rank * base + min(rank, extra)
    return list(range(lo, lo + base + (1 if rank < extra else 0)))


def band_rows(H, rank, world, align=64):
    """rows [y0, y1) of band `rank`: equal shares of the 64-row GUI tiles (gui/dplug/gui/graphics.d:59-60)"""
    tiles = (H + align - 1) // align
    base, extra = divmod(tiles, world)
    t0 = rank * base + min(rank, extra)
    t1 = t0 + base + (1 if rank < extra else 0)
    return min(t0 * align, H), min(t1 * align, H)


def need_rows(W, H, y0, y1, which, level=0):
    """rows [lo, hi) of a pyramid level that band [y0, y1) must hold valid data for"""
    lo, hi = C.c_int(), C.c_int()
    _lib.check(_lib.lib().b2_band_need_rows(W, H, y0, y1, which, level, C.byref(lo), C.byref(hi)))
    return lo.value, hi.value


def halo_plan(W, H, world, align=64):
    """For every rank: the level-0 rows it must receive, as {rank: [(map, src_rank, row_lo, row_hi), ...]}.
    Raises if a halo reaches beyond the adjacent band (bands thinner than the bloom reach)."""
    bands = [band_rows(H, r, world, align) for r in range(world)]
    plan = {r: [] for r in range(world)}
    for r, (y0, y1) in enumerate(bands):
        if y1 <= y0:
            continue
        for which in (DIFFUSE, DEPTH):
            lo, hi = need_rows(W, H, y0, y1, which, 0)
            if lo < y0:
                if r == 0 or lo < bands[r - 1][0]:
                    raise ValueError("band %d: halo reaches past the adjacent band" % r)
                plan[r].append((which, r - 1, lo, y0))
            if hi > y1:
                if r == world - 1 or hi > bands[r + 1][1]:
                    raise ValueError("band %d: halo reaches past the adjacent band" % r)
                plan[r].append((which, r + 1, y1, hi))
    return bands, plan


def post_halo_exchange(rank, world, plan, stores, dist):
    """Starts the exchange of the plan and returns the pending requests.  `stores[map] = (tensor2d, first_row)`: a
    (rows, pitch_bytes) uint8 torch tensor viewing the rows this rank stores for level 0 of `map` (global row
    `first_row` first).  Sends are derived from the other ranks' receive lists, so every rank posts matching
    isend / irecv pairs in one batch.  Only halo rows (rows outside the band) are written, so work that reads the
    band's own rows — `regenerateMipmaps(own_and_halo_rects(...)[0])` — may run while the rows are in flight."""
    ops = []
    for dst in range(world):
        for (which, src, lo, hi) in plan[dst]:
            if src == rank and dst != rank:
                t, first = stores[which]
                ops.append(dist.P2POp(dist.isend, t[lo - first:hi - first], dst))
            if dst == rank and src != rank:
                t, first = stores[which]
                ops.append(dist.P2POp(dist.irecv, t[lo - first:hi - first], src))
    return dist.batch_isend_irecv(ops) if ops else []


def wait_halo_exchange(reqs):
    for req in reqs:
        req.wait()
    return len(reqs)


def exchange_halos(rank, world, plan, stores, dist):
    """Executes the plan and waits for it; returns the number of point-to-point operations of this rank."""
    return wait_halo_exchange(post_halo_exchange(rank, world, plan, stores, dist))


def own_and_halo_rects(W, H, y0, y1):
    """Update rectangles of a band's frame in two steps: ([own rows], [halo rows above, halo rows below]).
    regenerateMipmaps(own) followed — once the halos have arrived — by regenerateMipmaps(halo) leaves every pyramid
    level exactly as one regenerateMipmaps over everything would: the second call recomputes, through the
    impactOnNextLevel chain, every mip texel that depends on a halo row (the same argument that makes the reference's
    dirty-rectangle updates correct, gui/dplug/gui/graphics.d:1378-1419)."""
    lo, hi = need_rows(W, H, y0, y1, DIFFUSE, 0)      # the widest halo (level-5 bicubic bloom)
    halos = []
    if lo < y0:
        halos.append((0, lo, W, y0))
    if hi > y1:
        halos.append((0, y1, W, hi))
    return [(0, y0, W, y1)], halos


_RECIPES = {DIFFUSE: (2, 1, 1, 1, 1), DEPTH: (0, 0, 1, 1)}   # qualities of levels 1.. (gui/dplug/gui/graphics.d:1380-1384, 1414)


def _rows_written(W, H, rect, which):
    """rows [lo, hi) of levels 1.. that regenerateMipmaps([rect]) writes (the impactOnNextLevel chain)"""
    from .boxlist import impactOnNextLevel
    out, area, w, h = [], rect, W, H
    for q in _RECIPES[which]:
        if w >> 1 == 0 or h >> 1 == 0:
            break
        area = impactOnNextLevel(q, area, w, h)
        out.append((area[1], area[3]))
        w, h = w >> 1, h >> 1
    return out


def interior_tile_rows(W, H, y0, y1, align=64):
    """Rows [a, b) of a band (whole GUI tiles) whose compositing reads neither a level-0 halo row nor a mip texel that
    regenerateMipmaps(halo rectangles) writes: these tiles can be composited after regenerateMipmaps(own rows) while
    the halo rows are still in flight and while the halo pass runs on another stream; the rest of the band
    ([y0, a) and [b, y1)) follows once the halo pass is done."""
    _, halos = own_and_halo_rects(W, H, y0, y1)
    a, b = y0, y1
    for rect in halos:
        top = rect[1] < y0
        written = {which: _rows_written(W, H, rect, which) for which in (DIFFUSE, DEPTH)}
        while a < b:
            cand = (a, min(a + align, b)) if top else (max(b - align, a), b)
            clash = False
            for which in (DIFFUSE, DEPTH):
                lo0, hi0 = need_rows(W, H, cand[0], cand[1], which, 0)
                clash = clash or lo0 < y0 or hi0 > y1
                for level, (wlo, whi) in enumerate(written[which], 1):
                    nlo, nhi = need_rows(W, H, cand[0], cand[1], which, level)
                    clash = clash or (nlo < whi and wlo < nhi)
            if not clash:
                break
            if top:
                a = cand[1]
            else:
                b = cand[0]
    return a, b


class BandGroup:
    """ctypes wrapper of b2band (include/b2pbr.h): the halo exchange and the frame schedule of row bands.

    BandGroup.nccl(band, rank, world, H, dist): one process per GPU; the NCCL unique id is created by rank 0 in the
    library and broadcast over `dist` (any torch.distributed group; only 128 bytes of host data travel).
    BandGroup.peers(bands): every band in this process (one thread drives several GPUs; copy-engine pushes)."""

    def __init__(self, handle, keep):
        self._h, self._keep, self._l = handle, keep, _lib.lib()

    @staticmethod
    def band_y(H, world, align=64):
        return [band_rows(H, r, world, align)[0] for r in range(world)] + [band_rows(H, world - 1, world, align)[1]]

    @classmethod
    def nccl(cls, band, rank, world, H, dist=None, align=64):
        l = _lib.lib()
        ident = C.create_string_buffer(128)
        if world > 1:
            import torch
            if rank == 0:
                _lib.check(l.b2_band_nccl_unique_id(ident))
            box = [bytes(ident.raw)]
            dist.broadcast_object_list(box, src=0)
            ident = C.create_string_buffer(box[0], 128)
        ys = (C.c_int * (world + 1))(*cls.band_y(H, world, align))
        h = C.c_void_p()
        _lib.check(l.b2band_create_nccl(C.byref(h), band.handle, rank, world, ys, ident, None))
        return cls(h, [band])

    @classmethod
    def peers(cls, bands):
        l = _lib.lib()
        arr = (C.c_void_p * len(bands))(*[b.handle for b in bands])
        h = C.c_void_p()
        _lib.check(l.b2band_create_peers(C.byref(h), arr, len(bands)))
        return cls(h, list(bands))

    def close(self):
        if self._h:
            self._l.b2band_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def exchangeBegin(self): _lib.check(self._l.b2band_exchange_begin(self._h))
    def exchangeEnd(self): _lib.check(self._l.b2band_exchange_end(self._h))
    def compositeFrame(self): _lib.check(self._l.b2band_composite_frame(self._h))
    def sync(self): _lib.check(self._l.b2band_sync(self._h))
    def haloBytesPerFrame(self): return int(self._l.b2band_halo_bytes_per_frame(self._h))

    def info(self, index=0):
        v = [C.c_int() for _ in range(5)]
        _lib.check(self._l.b2band_info(self._h, index, *[C.byref(x) for x in v]))
        return dict(zip(("rank", "interior_lo", "interior_hi", "n_sends", "n_recvs"), (x.value for x in v)))


def band_plan(W, H, y0, y1, align=64):
    """the library's planning of one band: (halo rectangles, (interior_lo, interior_hi))"""
    rects = (_lib.box2i * 2)()
    n, lo, hi = C.c_int(), C.c_int(), C.c_int()
    _lib.check(_lib.lib().b2_band_plan(W, H, y0, y1, align, rects, C.byref(n), C.byref(lo), C.byref(hi)))
    return [rects[k].astuple() for k in range(n.value)], (lo.value, hi.value)
