"""Deterministic synthetic inputs of the named shapes.

Counter-based: every texel is a function of (seed, map id, x, y) through splitmix64, so any row band can
be produced independently and the oracle and the GPU always receive identical bytes.
"""
import numpy as np

SEED = 0xD1B54A32D192ED03
_GOLD = np.uint64(0x9E3779B97F4A7C15)
MAP_DIFFUSE, MAP_MATERIAL, MAP_DEPTH, MAP_SKYBOX, MAP_EMISSIVE = 1, 2, 3, 4, 5


def splitmix64(x):
    with np.errstate(over="ignore"):
        z = (x + _GOLD).astype(np.uint64)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def _hash(seed, map_id, xs, ys, stride):
    with np.errstate(over="ignore"):
        key = np.uint64(seed) ^ (np.uint64(map_id) * _GOLD)
        idx = ys.astype(np.uint64)[:, None] * np.uint64(stride) + xs.astype(np.uint64)[None, :]
        return splitmix64(idx ^ key)


def _lattice_rgb(seed, map_id, W, y0, y1, cell=8):
    """bilinear up-sampling of a random lattice with `cell`-pixel spacing -> float (rows, W, 3) in 0..255"""
    xs = np.arange(W)
    ys = np.arange(y0, y1)
    gx0, gy0 = xs // cell, ys // cell
    fx = ((xs % cell) / float(cell))[None, :, None]
    fy = ((ys % cell) / float(cell))[:, None, None]
    stride = W // cell + 2

    def corner(gy, gx):
        h = _hash(seed, map_id, gx, gy, stride)
        return np.stack([(h >> np.uint64(8 * k)) & np.uint64(0xff) for k in range(3)], axis=-1).astype(np.float64)

    c00, c01 = corner(gy0, gx0), corner(gy0, gx0 + 1)
    c10, c11 = corner(gy0 + 1, gx0), corner(gy0 + 1, gx0 + 1)
    top = c00 * (1 - fx) + c01 * fx
    bot = c10 * (1 - fx) + c11 * fx
    return top * (1 - fy) + bot * fy


def diffuse(W, H, seed=SEED, y0=0, y1=None):
    """RGBA8: RGB = smooth lattice + +-8 noise; A (emissive) = 0 except LED blocks (~2 % area) with A in {128,255}."""
    y1 = H if y1 is None else y1
    xs, ys = np.arange(W), np.arange(y0, y1)
    rgb = _lattice_rgb(seed, MAP_DIFFUSE, W, y0, y1)
    nz = _hash(seed, MAP_DIFFUSE + 16, xs, ys, W)
    noise = np.stack([((nz >> np.uint64(8 * k)) & np.uint64(0xf)).astype(np.float64) - 8.0 for k in range(3)], axis=-1)
    rgb = np.clip(np.floor(rgb + noise), 0, 255).astype(np.uint8)
    # LED blocks: 16x16 cells, ~2 % of them lit
    cell = _hash(seed, MAP_EMISSIVE, xs // 16, ys // 16, W // 16 + 1)
    lit = (cell % np.uint64(50)) == 0
    level = np.where(((cell >> np.uint64(20)) & np.uint64(1)) == 1, 255, 128)
    a = np.where(lit, level, 0).astype(np.uint8)
    return np.ascontiguousarray(np.concatenate([rgb, a[..., None]], axis=-1))


def material(W, H, seed=SEED, y0=0, y1=None):
    """RGBA8: R roughness per 64x64 block in {32,128,220}; G metalness in {25,255}; B specular 128; A 255."""
    y1 = H if y1 is None else y1
    xs, ys = np.arange(W), np.arange(y0, y1)
    cell = _hash(seed, MAP_MATERIAL, xs // 64, ys // 64, W // 64 + 1)
    rough = np.array([32, 128, 220], dtype=np.uint8)[(cell % np.uint64(3)).astype(np.int64)]
    metal = np.where(((cell >> np.uint64(13)) & np.uint64(1)) == 1, 255, 25).astype(np.uint8)
    out = np.empty((y1 - y0, W, 4), dtype=np.uint8)
    out[..., 0] = rough
    out[..., 1] = metal
    out[..., 2] = 128
    out[..., 3] = 255
    return out


def depth(W, H, seed=SEED, y0=0, y1=None, kind="smooth"):
    """L16.  smooth: 15000 + four sinusoids (amplitude <= 12000 in total) + 2000-high bevel steps every 256 px.
    noise: uniform u16 (worst case for the variance / LOD logic)."""
    y1 = H if y1 is None else y1
    xs, ys = np.arange(W), np.arange(y0, y1)
    if kind == "noise":
        return (_hash(seed, MAP_DEPTH, xs, ys, W) & np.uint64(0xffff)).astype(np.uint16)
    X = xs[None, :].astype(np.float64)
    Y = ys[:, None].astype(np.float64)
    z = 15000.0 + 3000.0 * np.sin(X / 37.0) + 3000.0 * np.sin(Y / 53.0) \
        + 3000.0 * np.sin((X + Y) / 97.0) + 3000.0 * np.sin((X - 2.0 * Y) / 23.0)
    z = z + 2000.0 * (((xs // 256)[None, :] + (ys // 256)[:, None]) % 2)
    return np.clip(np.floor(z), 0, 65535).astype(np.uint16)


def skybox(W=1024, H=768, seed=SEED):
    """RGBA8 smooth gradient + noise; 1024x768 like examples/distort/gfx/skybox.jpg."""
    xs, ys = np.arange(W), np.arange(H)
    rgb = _lattice_rgb(seed, MAP_SKYBOX, W, 0, H, cell=32)
    nz = _hash(seed, MAP_SKYBOX + 16, xs, ys, W)
    noise = np.stack([((nz >> np.uint64(8 * k)) & np.uint64(0x1f)).astype(np.float64) - 16.0 for k in range(3)], axis=-1)
    rgb = np.clip(np.floor(rgb + noise), 0, 255).astype(np.uint8)
    a = np.full((H, W, 1), 255, dtype=np.uint8)
    return np.ascontiguousarray(np.concatenate([rgb, a], axis=-1))


def frame(W, H, seed=SEED, depth_kind="smooth"):
    """(diffuse, material, depth) level-0 maps of one plug-in UI"""
    return diffuse(W, H, seed), material(W, H, seed), depth(W, H, seed, kind=depth_kind)
