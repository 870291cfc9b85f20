"""Multi-GPU sharding of the path: host-side planning only, no device work.

* batch of plug-in UIs (BASELINE config 4): instances are independent -> round-robin blocks, no communication;
* one large canvas (config 5): horizontal row bands; the only exchange is the level-0 halo rows (diffuse and
  depth) between neighbouring bands, once per frame.  `halo_plan` derives the rows from the same arithmetic
  the library uses (b2_band_need_rows), `exchange_halos` moves them with torch.distributed point-to-point
  operations (NCCL over NVLink on GPUs, gloo in the CPU tests).
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


def exchange_halos(rank, world, plan, stores, dist):
    """Executes the plan.  `stores[map] = (tensor2d, first_row)`: a (rows, pitch_bytes) uint8 torch tensor viewing
    the rows this rank stores for level 0 of `map` (global row `first_row` first).  Sends are derived from the
    other ranks' receive lists, so every rank posts matching isend / irecv pairs in one batch."""
    ops = []
    for dst in range(world):
        for (which, src, lo, hi) in plan[dst]:
            if src == rank and dst != rank:
                t, first = stores[which]
                ops.append(dist.P2POp(dist.isend, t[lo - first:hi - first], dst))
            if dst == rank and src != rank:
                t, first = stores[which]
                ops.append(dist.P2POp(dist.irecv, t[lo - first:hi - first], src))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return len(ops)
