"""ctypes binding of the CPU oracle (oracle/libb2ref.so) for the tests — the checker, never the product.

Class and method names follow the reference (Mipmap, PBRCompositor, GUIGraphics slice) so the parity tests
drive both back ends with the same calls.
"""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "libb2ref.so")

RGBA8, L16 = 0, 1
Q_BOX, Q_CUBIC, Q_PREMUL = 0, 1, 2


class box2i(C.Structure):
    _fields_ = [("min_x", C.c_int), ("min_y", C.c_int), ("max_x", C.c_int), ("max_y", C.c_int)]

    def astuple(self):
        return (self.min_x, self.min_y, self.max_x, self.max_y)


class imageref(C.Structure):
    _fields_ = [("w", C.c_int), ("h", C.c_int), ("pitch", C.c_size_t), ("pixels", C.c_void_p)]


class params(C.Structure):
    _fields_ = [("light1_color", C.c_float * 3), ("light2_dir", C.c_float * 3), ("light2_color", C.c_float * 3),
                ("light3_dir", C.c_float * 3), ("light3_color", C.c_float * 3), ("ambient_light", C.c_float),
                ("skybox_amount", C.c_float), ("tonemap_threshold", C.c_float), ("tonemap_ratio", C.c_float),
                ("pass_active_mask", C.c_uint32)]


_lib = None


def build():
    subprocess.run(["make", "-C", ORACLE_DIR, "libb2ref.so"], check=True, stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith((".cpp", ".hpp", ".h", ".inc"))]
        if not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs):
            build()
        l = C.CDLL(LIB_PATH)
        V, I, F, D = C.c_void_p, C.c_int, C.c_float, C.c_double
        sig = {
            "b2ref_mip_create": (V, [I, I, I, I]), "b2ref_mip_destroy": (None, [V]), "b2ref_mip_num_levels": (I, [V]),
            "b2ref_mip_level": (I, [V, I, C.POINTER(imageref)]),
            "b2ref_mip_size_level": (I, [V, I, I, I, I, I, I, I]),
            "b2ref_mip_replicate_borders_touching": (I, [V, I, box2i]),
            "b2ref_mip_generate_next_level": (I, [V, I, box2i, I, C.POINTER(box2i)]),
            "b2ref_mip_generate_mipmaps": (I, [V, I]),
            "b2ref_mip_linear_sample": (I, [V, I, F, F, C.POINTER(F)]),
            "b2ref_mip_cubic_sample": (I, [V, I, F, F, C.POINTER(F)]),
            "b2ref_mip_linear_mipmap_sample": (I, [V, F, F, F, C.POINTER(F)]),
            "b2ref_default_params": (None, [C.POINTER(params)]),
            "b2ref_pbr_create": (V, [I]), "b2ref_pbr_destroy": (None, [V]), "b2ref_pbr_resize": (I, [V, I, I, I, I]),
            "b2ref_pbr_set_params": (I, [V, C.POINTER(params)]),
            "b2ref_pbr_set_skybox": (I, [V, V, I, I, C.c_size_t]),
            "b2ref_pbr_composite_tile": (I, [V, imageref, C.POINTER(box2i), I, V, V, V]),
            "b2ref_gfx_create": (V, [I]), "b2ref_gfx_destroy": (None, [V]), "b2ref_gfx_resize": (I, [V, I, I]),
            "b2ref_gfx_compositor": (V, [V]), "b2ref_gfx_map": (V, [V, I]), "b2ref_gfx_output": (I, [V, C.POINTER(imageref)]),
            "b2ref_gfx_regenerate_mipmaps": (I, [V, C.POINTER(box2i), I]),
            "b2ref_gfx_composite_gui": (I, [V, C.POINTER(box2i), I]),
            "b2ref_draw_layer": (I, [imageref] * 9 + [I, I, C.POINTER(box2i), I, I]),
            "b2ref_screenshot_qb": (C.c_size_t, [imageref, I, imageref, V, C.c_size_t]),
            "b2ref_screenshot_png_pixels": (I, [imageref, I, V, C.c_size_t]),
            "b2ref_resize_image": (I, [I, V, I, I, C.c_size_t, V, I, I, C.c_size_t]),
            "b2ref_present": (I, [imageref, imageref, I, imageref, I, I, C.POINTER(box2i), I, V, I]),
            "b2ref_tile_areas": (I, [C.POINTER(box2i), I, I, I, C.POINTER(box2i), I]),
            "b2ref_remove_overlapping_areas": (I, [C.POINTER(box2i), I, C.POINTER(box2i), I]),
            "b2ref_bounding_box": (box2i, [C.POINTER(box2i), I]), "b2ref_have_no_overlap": (I, [C.POINTER(box2i), I]),
            "b2ref_dirty_rect_list_add": (I, [C.POINTER(box2i), I, I, box2i]),
            "b2ref_impact_on_next_level": (box2i, [I, box2i, I, I]),
            "b2ref_linmap": (D, [D, D, D, D, D]), "b2ref_rsqrtss": (F, [F, I]), "b2ref_set_rsqrt_mode": (None, [I]),
            "b2ref_rsqrtss_table_vs_host": (C.c_long, []), "b2ref_pow_ps": (F, [F, F]), "b2ref_fastlog2": (F, [F]),
            "b2ref_plane_fitting_normal": (None, [C.POINTER(F), C.POINTER(F)]),
            "b2ref_image_layout": (None, [I, I, I, I, I, I, I, I, C.POINTER(I), C.POINTER(I), C.POINTER(I)]),
            "b2ref_kat_replicate_borders": (I, [C.POINTER(C.c_uint8), C.POINTER(I), C.POINTER(I), C.POINTER(I)]),
            "b2ref_compute_right_padding": (I, [I, I, I]), "b2ref_next_multiple_of": (C.c_size_t, [C.c_size_t, C.c_size_t]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


def boxes(rects):
    arr = (box2i * max(1, len(rects)))()
    for i, r in enumerate(rects):
        arr[i] = box2i(*[int(v) for v in r])
    return arr


def _view(ref, color):
    """numpy view (no copy) of an oracle image: (h, w, 4) u8 or (h, w) u16, honouring the pitch"""
    if ref.w == 0 or ref.h == 0:
        return np.zeros((ref.h, ref.w, 4) if color == RGBA8 else (ref.h, ref.w), dtype=np.uint8 if color == RGBA8 else np.uint16)
    nbytes = ref.pitch * (ref.h - 1) + ref.w * (4 if color == RGBA8 else 2)
    buf = (C.c_uint8 * nbytes).from_address(ref.pixels)
    a = np.frombuffer(buf, dtype=np.uint8)
    if color == RGBA8:
        return np.lib.stride_tricks.as_strided(a, shape=(ref.h, ref.w, 4), strides=(ref.pitch, 4, 1))
    a16 = np.frombuffer(buf, dtype=np.uint16, count=nbytes // 2)
    return np.lib.stride_tricks.as_strided(a16, shape=(ref.h, ref.w), strides=(ref.pitch, 2))


class Mipmap:
    def __init__(self, color, maxLevel, w, h, _handle=None):
        self.l = lib()
        self.color = color
        self._owned = _handle is None
        self.h = C.c_void_p(_handle) if _handle is not None else C.c_void_p(self.l.b2ref_mip_create(color, maxLevel, w, h))

    def __del__(self):
        if getattr(self, "_owned", False) and self.h:
            self.l.b2ref_mip_destroy(self.h)
            self.h = None

    def numLevels(self):
        return self.l.b2ref_mip_num_levels(self.h)

    def level(self, level):
        ref = imageref()
        assert self.l.b2ref_mip_level(self.h, level, C.byref(ref)) == 0
        return _view(ref, self.color)

    def levelRef(self, level):
        ref = imageref()
        assert self.l.b2ref_mip_level(self.h, level, C.byref(ref)) == 0
        return ref

    def sizeLevel(self, level, w, h, border, rowAlign, xMult, trailing):
        assert self.l.b2ref_mip_size_level(self.h, level, w, h, border, rowAlign, xMult, trailing) == 0

    def replicateBordersTouching(self, level, rect):
        assert self.l.b2ref_mip_replicate_borders_touching(self.h, level, box2i(*rect)) == 0

    def upload(self, level, image):
        self.level(level)[...] = image

    def download(self, level):
        return np.array(self.level(level))

    def generateNextLevel(self, quality, rect, level):
        out = box2i()
        assert self.l.b2ref_mip_generate_next_level(self.h, quality, box2i(*rect), level, C.byref(out)) == 0
        return out.astuple()

    def generateMipmaps(self, quality):
        assert self.l.b2ref_mip_generate_mipmaps(self.h, quality) == 0

    def _sample(self, fn, level, x, y, lvl_float=False):
        x = np.asarray(x, dtype=np.float32).ravel()
        y = np.asarray(y, dtype=np.float32).ravel()
        level = np.broadcast_to(np.asarray(level, dtype=np.float32), x.shape)
        out = np.zeros((x.size, 4), dtype=np.float32)
        tmp = (C.c_float * 4)()
        for i in range(x.size):
            lv = C.c_float(float(level[i])) if lvl_float else int(level[i])
            fn(self.h, lv, C.c_float(float(x[i])), C.c_float(float(y[i])), tmp)
            out[i] = tmp[:]
        return out if self.color == RGBA8 else out[:, 0]

    def linearSample(self, level, x, y):
        return self._sample(self.l.b2ref_mip_linear_sample, level, x, y)

    def cubicSample(self, level, x, y):
        return self._sample(self.l.b2ref_mip_cubic_sample, level, x, y)

    def linearMipmapSample(self, level, x, y):
        return self._sample(self.l.b2ref_mip_linear_mipmap_sample, level, x, y, lvl_float=True)


class DirtyRectList:
    """DirtyRectList.addRect on a Python-held list (boxlist.d:236-292)"""

    def __init__(self):
        self.rects = []

    def addRect(self, rect):
        cap = len(self.rects) * 4 + 8
        arr = (box2i * cap)(*[box2i(*r) for r in self.rects])
        n = lib().b2ref_dirty_rect_list_add(arr, len(self.rects), cap, box2i(*rect))
        assert n <= cap
        self.rects = [arr[i].astuple() for i in range(n)]


def draw_layer(dst, src, opacity, x, y, rects, mode):
    """dst / src / opacity: (diffuse, depth, material) numpy triples; dst is updated in place."""
    def ref(a):
        return imageref(a.shape[1], a.shape[0], a.strides[0], a.ctypes.data)
    op = opacity if opacity is not None else src
    rc = lib().b2ref_draw_layer(*[ref(a) for a in dst], *[ref(a) for a in src], *[ref(a) for a in op], x, y, boxes(rects), len(rects), mode)
    assert rc == 0


def present(composited, rendered, pixelFormat, window, userOrigin, rects, table=None, redrawBorders=False):
    """doDraw stages E-H on host arrays ((h, w, 4) uint8); `rendered` and `window` are updated in place."""
    def ref(a):
        return imageref(a.shape[1], a.shape[0], a.strides[0], a.ctypes.data)
    t = None if table is None else np.ascontiguousarray(table, dtype=np.uint8).reshape(768)
    rc = lib().b2ref_present(ref(composited), ref(rendered), pixelFormat, ref(window), userOrigin[0], userOrigin[1], boxes(rects), len(rects),
                             None if t is None else t.ctypes.data_as(C.c_void_p), 1 if redrawBorders else 0)
    assert rc == 0
    return window


def encodeScreenshotAsQB(rendered, pixelFormat, depth):
    """screencap.d:21-84 on host arrays: rendered (h, w, 4) uint8 in the window's pixel format, depth (h, w) uint16"""
    def ref(a):
        return imageref(a.shape[1], a.shape[0], a.strides[0], a.ctypes.data)
    n = lib().b2ref_screenshot_qb(ref(rendered), pixelFormat, ref(depth), None, 0)
    out = np.empty(n, dtype=np.uint8)
    assert lib().b2ref_screenshot_qb(ref(rendered), pixelFormat, ref(depth), out.ctypes.data_as(C.c_void_p), n) == n
    return out.tobytes()


def screenshotPixels(rendered, pixelFormat):
    """the pixel pass of encodeScreenshotAsPNG (screencap.d:88-131)"""
    out = np.empty_like(rendered)
    assert lib().b2ref_screenshot_png_pixels(imageref(rendered.shape[1], rendered.shape[0], rendered.strides[0], rendered.ctypes.data), pixelFormat,
                                             out.ctypes.data_as(C.c_void_p), out.strides[0]) == 0
    return out


RESIZE_DIFFUSE, RESIZE_MATERIAL, RESIZE_DEPTH, RESIZE_GENERIC_L16, RESIZE_SMOOTHER = range(5)


def resizeImage(kind, src, out_w, out_h):
    """ImageResizer over the in-tree stb_image_resize port (resizer.d / stb_image_resize.d, V2 = false): src is (h, w, 4)
    uint8 for the RGBA kinds, (h, w) uint16 for the L16 kinds"""
    src = np.ascontiguousarray(src)
    out = np.zeros((out_h, out_w) + src.shape[2:], dtype=src.dtype)
    rc = lib().b2ref_resize_image(kind, src.ctypes.data_as(C.c_void_p), src.shape[1], src.shape[0], src.strides[0],
                                  out.ctypes.data_as(C.c_void_p), out_w, out_h, out.strides[0])
    assert rc == 0
    return out


def default_params():
    p = params()
    lib().b2ref_default_params(C.byref(p))
    return p


class Graphics:
    """The GUIGraphics slice: doResize sizing + regenerateMipmaps + compositeGUI (graphics.d:1115-1137,1338-1423)."""

    def __init__(self, W, H, numThreads=3):
        self.l = lib()
        self.g = C.c_void_p(self.l.b2ref_gfx_create(numThreads))
        assert self.l.b2ref_gfx_resize(self.g, W, H) == 0
        self.W, self.H = W, H
        self.pbr = C.c_void_p(self.l.b2ref_gfx_compositor(self.g))
        self.diffuseMap = Mipmap(RGBA8, 0, 0, 0, _handle=self.l.b2ref_gfx_map(self.g, 0))
        self.materialMap = Mipmap(RGBA8, 0, 0, 0, _handle=self.l.b2ref_gfx_map(self.g, 1))
        self.depthMap = Mipmap(L16, 0, 0, 0, _handle=self.l.b2ref_gfx_map(self.g, 2))

    def __del__(self):
        if getattr(self, "g", None):
            self.l.b2ref_gfx_destroy(self.g)
            self.g = None

    def set_params(self, p):
        q = params()
        C.memmove(C.byref(q), C.byref(p), C.sizeof(params))  # same layout as the product's struct
        assert self.l.b2ref_pbr_set_params(self.pbr, C.byref(q)) == 0

    def setSkybox(self, image):
        image = np.ascontiguousarray(image, dtype=np.uint8)
        assert self.l.b2ref_pbr_set_skybox(self.pbr, image.ctypes.data_as(C.c_void_p), image.shape[1], image.shape[0], image.strides[0]) == 0

    def setInputs(self, diffuse, material, depth):
        self.diffuseMap.upload(0, diffuse)
        self.materialMap.upload(0, material)
        self.depthMap.upload(0, depth)

    def setInputRows(self, diffuse, material, depth, row0=0):
        """level-0 rows [row0, row0 + n) only (bounded samples of large canvases)"""
        n = diffuse.shape[0]
        self.diffuseMap.level(0)[row0:row0 + n] = diffuse
        self.materialMap.level(0)[row0:row0 + n] = material
        self.depthMap.level(0)[row0:row0 + n] = depth

    def regenerateMipmaps(self, rects):
        assert self.l.b2ref_gfx_regenerate_mipmaps(self.g, boxes(rects), len(rects)) == 0

    def compositeGUI(self, rects):
        assert self.l.b2ref_gfx_composite_gui(self.g, boxes(rects), len(rects)) == 0

    def output(self):
        ref = imageref()
        assert self.l.b2ref_gfx_output(self.g, C.byref(ref)) == 0
        return _view(ref, RGBA8)

    def fullFrame(self):
        whole = [(0, 0, self.W, self.H)]
        self.regenerateMipmaps(whole)
        self.compositeGUI(whole)
        return np.array(self.output())
