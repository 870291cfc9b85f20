"""Device PBRCompositor (mirror of gui/dplug/gui/legacypbr.d:67-266 and ICompositor, compositor.d:29-69)."""
import ctypes as C
import numpy as np
from . import _lib
from .mipmap import Mipmap, RGBA8, L16
from .boxlist import tileAreas

PBR_TILE_MAX_WIDTH = 64   # gui/dplug/gui/graphics.d:59
PBR_TILE_MAX_HEIGHT = 64  # gui/dplug/gui/graphics.d:60

DIFFUSE, MATERIAL, DEPTH = 0, 1, 2
PF_RGBA8, PF_BGRA8, PF_ARGB8 = 0, 1, 2
MODE_FAST, MODE_EXACT = 0, 1


def liftGammaGainContrastTable(r=(0.0, 1.0, 1.0, 0.0), g=None, b=None):
    """UIColorCorrection.setLiftGammaGainContrastRGB: (lift, gamma, gain, contrast) per channel -> (3, 256) uint8."""
    g = r if g is None else g
    b = r if b is None else b
    v = (C.c_float * 12)(*[float(x) for x in (tuple(r) + tuple(g) + tuple(b))])
    out = np.zeros((3, 256), dtype=np.uint8)
    _lib.lib().b2_lift_gamma_gain_contrast_table(v, out.ctypes.data_as(C.c_void_p))
    return out


def default_params():
    p = _lib.params()
    _lib.lib().b2pbr_default_params(C.byref(p))
    return p


def _ref(arr):
    return _lib.imageref(arr.shape[1], arr.shape[0], arr.strides[0], arr.ctypes.data)


class Layer:
    """Device-resident cache of a widget's PBR render (UIBufferedElementPBR buffers / PBRBackgroundGUI maps)."""
    DIFFUSE, DEPTH, MATERIAL, DIFFUSE_OPACITY, DEPTH_OPACITY, MATERIAL_OPACITY = range(6)
    BLIT, BLEND = 0, 1

    def __init__(self, w, h, withOpacity=True, device=0):
        self._l = _lib.lib()
        h_ = C.c_void_p()
        _lib.check(self._l.b2layer_create(C.byref(h_), device, w, h, 1 if withOpacity else 0))
        self._h, self.w, self.h = h_, w, h

    def close(self):
        if self._h:
            self._l.b2layer_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, plane, image, rects=None):
        dt = np.uint16 if plane == self.DEPTH else np.uint8
        image = np.ascontiguousarray(image, dtype=dt)
        if rects is None:
            rects = [(0, 0, self.w, self.h)]
        _lib.check(self._l.b2layer_upload(self._h, plane, image.ctypes.data_as(C.c_void_p), image.strides[0], _lib.boxes(rects), len(rects)))


    def resizePlane(self, plane, kind, image):
        """ImageResizer into one plane of this device-resident cache (PBRBackgroundGUI.resizeFun,
        pbrbackgroundgui.d:172-199): `image` is the widget's full-size source, `kind` a resizer.RESIZE_* constant."""
        image = np.ascontiguousarray(image, dtype=np.uint16 if plane == self.DEPTH else np.uint8)
        _lib.check(self._l.b2layer_resize_plane(self._h, plane, kind, image.ctypes.data_as(C.c_void_p), image.shape[1], image.shape[0], image.strides[0]))


class PBRCompositor:
    """`PBRCompositor(device)`; owns the device twins of the maps GUIGraphics lends to compositeTile."""

    PASS_NORMAL, PASS_AO, PASS_OBLIQUE_SHADOW, PASS_DIRECTIONAL = 0, 1, 2, 3
    PASS_SPECULAR, PASS_SKYBOX, PASS_EMISSIVE, PASS_CLAMP = 4, 5, 6, 7

    def __init__(self, device=0, band=None):
        self._l = _lib.lib()
        h = C.c_void_p()
        if band is None:
            _lib.check(self._l.b2pbr_create(C.byref(h), device))
        else:
            _lib.check(self._l.b2pbr_create_band(C.byref(h), device, band[0], band[1]))
        self._h = h
        self.device = device
        self.width = self.height = 0

    def close(self):
        if self._h:
            self._l.b2pbr_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def set_stream(self, cuda_stream):
        _lib.check(self._l.b2pbr_set_stream(self._h, C.c_void_p(cuda_stream)))

    # --- ICompositor -------------------------------------------------------------------------
    def resizeBuffers(self, width, height, areaMaxWidth=PBR_TILE_MAX_WIDTH, areaMaxHeight=PBR_TILE_MAX_HEIGHT):
        _lib.check(self._l.b2pbr_resize(self._h, width, height, areaMaxWidth, areaMaxHeight))
        self.width, self.height = width, height

    def compositeTile(self, wfb, areas, diffuseMap=None, materialMap=None, depthMap=None, updateRects=None):
        """ICompositor.compositeTile.  With numpy level-0 images this is the host drop-in (upload dirty rects,
        regenerate device mips, composite, copy `areas` back into `wfb`); with device `Mipmap`s (or None = the
        compositor's own twins) it composites device-resident data and `wfb` may be None."""
        n = len(areas)
        arr = _lib.boxes(areas)
        if isinstance(diffuseMap, np.ndarray):
            upd = _lib.boxes(updateRects) if updateRects is not None else None
            _lib.check(self._l.b2pbr_composite_tile_host(self._h, _ref(wfb), arr, n, upd, 0 if updateRects is None else len(updateRects),
                                                         _ref(diffuseMap), _ref(materialMap), _ref(depthMap)))
            return wfb
        hd = diffuseMap.handle if diffuseMap is not None else None
        hm = materialMap.handle if materialMap is not None else None
        hz = depthMap.handle if depthMap is not None else None
        _lib.check(self._l.b2pbr_composite_tile(self._h, arr, n, hd, hm, hz))
        if wfb is not None:
            self.download(areas, out=wfb)
        return wfb

    def compositeTileBegin(self, wfb, areas, diffuseMap, materialMap, depthMap, updateRects=None):
        """compositeTile with host images, enqueued only (page-locked images): several compositors — a batch of plug-in
        UIs — can have their calls in flight together.  The arrays must stay alive and unmodified until compositeTileEnd."""
        arr = _lib.boxes(areas)
        upd = _lib.boxes(updateRects) if updateRects is not None else None
        self._inflight = (wfb, diffuseMap, materialMap, depthMap, arr, upd)
        _lib.check(self._l.b2pbr_composite_tile_host_begin(self._h, _ref(wfb), arr, len(areas), upd, 0 if updateRects is None else len(updateRects),
                                                           _ref(diffuseMap), _ref(materialMap), _ref(depthMap)))

    def compositeTileEnd(self):
        _lib.check(self._l.b2pbr_composite_tile_host_end(self._h))
        wfb = self._inflight[0]
        self._inflight = None
        return wfb

    # --- legacy setters, legacypbr.d:75-123 -----------------------------------------------------
    def params(self):
        p = _lib.params()
        _lib.check(self._l.b2pbr_get_params(self._h, C.byref(p)))
        return p

    def set_params(self, p):
        _lib.check(self._l.b2pbr_set_params(self._h, C.byref(p)))

    def _set(self, name, value):
        p = self.params()
        if isinstance(value, (tuple, list, np.ndarray)):
            setattr(p, name, (C.c_float * 3)(*[float(v) for v in value]))
        else:
            setattr(p, name, value)
        self.set_params(p)

    def light1Color(self, c): self._set("light1_color", c)
    def light2Dir(self, d): self._set("light2_dir", d)
    def light2Color(self, c): self._set("light2_color", c)
    def light3Dir(self, d): self._set("light3_dir", d)
    def light3Color(self, c): self._set("light3_color", c)
    def skyboxAmount(self, a): self._set("skybox_amount", float(a))
    def ambientLight(self, a): self._set("ambient_light", float(a))
    def tonemapThreshold(self, v): self._set("tonemap_threshold", float(v))
    def tonemapRatio(self, v): self._set("tonemap_ratio", float(v))

    def setMode(self, mode):
        """MODE_FAST (default): throughput kernel, +-1 LSB; MODE_EXACT: reference operation order, bit-identical."""
        _lib.check(self._l.b2pbr_set_mode(self._h, int(mode)))

    def setPassActive(self, index, active):
        """getPass(index).setActive(active) — compositor.d:170-173,222-225."""
        p = self.params()
        m = p.pass_active_mask
        p.pass_active_mask = (m | (1 << index)) if active else (m & ~(1 << index))
        self.set_params(p)

    def setSkybox(self, image):
        """PassSkyboxReflections.setSkybox — legacypbr.d:861-870."""
        if image is None:
            _lib.check(self._l.b2pbr_set_skybox(self._h, None, 0, 0, 0))
            return
        image = np.ascontiguousarray(image, dtype=np.uint8)
        _lib.check(self._l.b2pbr_set_skybox(self._h, image.ctypes.data_as(C.c_void_p), image.shape[1], image.shape[0], image.strides[0]))

    # --- device twins -------------------------------------------------------------------------
    def map(self, which):
        h = self._l.b2pbr_map(self._h, which)
        if not h:
            raise _lib.B2Error("map before resizeBuffers")
        return Mipmap(RGBA8 if which != DEPTH else L16, 0, 0, 0, _handle=h)

    @property
    def diffuseMap(self): return self.map(DIFFUSE)
    @property
    def materialMap(self): return self.map(MATERIAL)
    @property
    def depthMap(self): return self.map(DEPTH)

    def skyboxMap(self):
        h = self._l.b2pbr_skybox(self._h)
        return Mipmap(RGBA8, 0, 0, 0, _handle=h) if h else None

    # --- GUIGraphics slice ----------------------------------------------------------------------
    def regenerateMipmaps(self, rects, stream=None):
        """GUIGraphics.regenerateMipmaps — gui/dplug/gui/graphics.d:1354-1423.  `stream` (a raw CUDA stream handle):
        enqueue on that stream instead of the compositor's."""
        if stream is None:
            _lib.check(self._l.b2pbr_regenerate_mipmaps(self._h, _lib.boxes(rects), len(rects)))
        else:
            _lib.check(self._l.b2pbr_regenerate_mipmaps_on(self._h, _lib.boxes(rects), len(rects), C.c_void_p(stream)))

    def redraw(self, updateRects, areas):
        """regenerateMipmaps(updateRects) + compositeTile(areas) as one call (GUIGraphics.doDraw runs them back to back,
        graphics.d:812-821)."""
        _lib.check(self._l.b2pbr_redraw(self._h, _lib.boxes(updateRects), len(updateRects), _lib.boxes(areas), len(areas)))

    def compositeGUI(self, rectsToCompositeDisjointed, wfb=None):
        """GUIGraphics.compositeGUI — gui/dplug/gui/graphics.d:1338-1349."""
        tiles = tileAreas(rectsToCompositeDisjointed, PBR_TILE_MAX_WIDTH, PBR_TILE_MAX_HEIGHT)
        return self.compositeTile(wfb, tiles)

    def compositeRaw(self, areas):
        """SimpleRawCompositor — compositor.d:260-298."""
        _lib.check(self._l.b2pbr_composite_raw(self._h, _lib.boxes(areas), len(areas)))

    def download(self, rects=None, out=None, pixelFormat=None):
        if out is None:
            out = np.zeros((self.height, self.width, 4), dtype=np.uint8)
        if rects is None:
            rects = [(0, 0, self.width, self.height)]
        arr = _lib.boxes(rects)
        if pixelFormat is None:
            _lib.check(self._l.b2pbr_download(self._h, out.ctypes.data_as(C.c_void_p), out.strides[0], arr, len(rects)))
        else:
            _lib.check(self._l.b2pbr_download_swizzled(self._h, pixelFormat, out.ctypes.data_as(C.c_void_p), out.strides[0], arr, len(rects)))
        return out

    def downloadRows(self, row0, row1):
        """rows [row0, row1) of the composited buffer (band objects hold only their own rows)"""
        out = np.zeros((row1 - row0, self.width, 4), dtype=np.uint8)
        if row1 > row0:
            origin = out.ctypes.data - row0 * out.strides[0]
            _lib.check(self._l.b2pbr_download(self._h, C.c_void_p(origin), out.strides[0], _lib.boxes([(0, row0, self.width, row1)]), 1))
        return out

    def downloadRowsInto(self, out, row0, row1):
        """rows [row0, row1) of the composited buffer into `out` ((row1 - row0, width, 4) uint8, e.g. pinned); synchronises"""
        origin = out.ctypes.data - row0 * out.strides[0]
        _lib.check(self._l.b2pbr_download(self._h, C.c_void_p(origin), out.strides[0], _lib.boxes([(0, row0, self.width, row1)]), 1))
        return out

    def outputInfo(self):
        """device ImageRef of the composited buffer: {"w", "h", "pitch", "pixels"} (pixels = device address of the first stored row)"""
        r = _lib.imageref()
        _lib.check(self._l.b2pbr_output(self._h, C.byref(r)))
        return {"w": r.w, "h": r.h, "pitch": int(r.pitch), "pixels": int(r.pixels)}

    def bandRows(self, which, level=0):
        """(stored_lo, stored_hi, need_lo, need_hi) rows of a pyramid level (see b2pbr_band_rows)."""
        v = [C.c_int() for _ in range(4)]
        _lib.check(self._l.b2pbr_band_rows(self._h, which, level, *[C.byref(x) for x in v]))
        return tuple(x.value for x in v)

    def graphBegin(self):
        """start recording this object's device work into a CUDA graph"""
        _lib.check(self._l.b2pbr_graph_begin(self._h))

    def graphEnd(self):
        gid = C.c_int()
        _lib.check(self._l.b2pbr_graph_end(self._h, C.byref(gid)))
        return gid.value

    def graphLaunch(self, gid):
        _lib.check(self._l.b2pbr_graph_launch(self._h, gid))

    def drawLayer(self, layer, x, y, dirtyRects, mode=Layer.BLEND):
        """the widget's onDrawPBR on the device: blit (PBRBackgroundGUI) or opacity-masked blend (UIBufferedElementPBR)"""
        _lib.check(self._l.b2pbr_draw_layer(self._h, layer._h, x, y, _lib.boxes(dirtyRects), len(dirtyRects), mode))

    # --- post-composite chain, GUIGraphics.doDraw stages E-H (graphics.d:825-868) ---------------------------
    def setColorCorrection(self, table):
        """UIColorCorrection transfer tables: (3, 256) uint8 (red, green, blue) or None."""
        if table is None:
            _lib.check(self._l.b2pbr_set_color_correction(self._h, None))
            return
        t = np.ascontiguousarray(table, dtype=np.uint8).reshape(768)
        _lib.check(self._l.b2pbr_set_color_correction(self._h, t.ctypes.data_as(C.c_void_p)))

    def present(self, window, pixelFormat, userOrigin, rects, redrawBorders=False):
        """blit + colour correction + reorderComponents + resizeContent into `window` ((h, w, 4) uint8)."""
        _lib.check(self._l.b2pbr_present(self._h, pixelFormat, window.ctypes.data_as(C.c_void_p), window.strides[0], window.shape[1],
                                         window.shape[0], userOrigin[0], userOrigin[1], _lib.boxes(rects), len(rects), 1 if redrawBorders else 0))
        return window

    # --- screenshot encoders (gui/dplug/gui/screencap.d:21-136) on the rendered buffer of doDraw stage G.bis -------
    def encodeScreenshotAsQB(self, pixelFormat):
        """encodeScreenshotAsQB (screencap.d:21-84): the .qb voxel file (W x 26 x H voxels) as bytes."""
        n = int(self._l.b2pbr_screenshot_qb_size(self._h))
        buf = np.empty(n, dtype=np.uint8)
        _lib.check(self._l.b2pbr_screenshot_qb(self._h, pixelFormat, buf.ctypes.data_as(C.c_void_p), n))
        return buf.tobytes()

    def screenshotPixels(self, pixelFormat):
        """the image encodeScreenshotAsPNG hands to its encoder (screencap.d:88-131): (h, w, 4) uint8, RGBA"""
        out = np.empty((self.height, self.width, 4), dtype=np.uint8)
        _lib.check(self._l.b2pbr_screenshot_rgba(self._h, pixelFormat, out.ctypes.data_as(C.c_void_p), out.strides[0]))
        return out

    def encodeScreenshotAsPNG(self, pixelFormat):
        """encodeScreenshotAsPNG (screencap.d:88-136): a PNG file as bytes."""
        n = int(self._l.b2pbr_screenshot_png_size(self._h))
        buf = np.empty(n, dtype=np.uint8)
        written = C.c_size_t(0)
        _lib.check(self._l.b2pbr_screenshot_png(self._h, pixelFormat, buf.ctypes.data_as(C.c_void_p), n, C.byref(written)))
        return buf[: written.value].tobytes()

    def sync(self):
        _lib.check(self._l.b2pbr_sync(self._h))

    def profileEnable(self, on=True):
        _lib.check(self._l.b2pbr_profile_enable(self._h, 1 if on else 0))

    def traceJSON(self):
        need = C.c_size_t()
        _lib.check(self._l.b2pbr_trace_json(self._h, None, 0, C.byref(need)))
        buf = C.create_string_buffer(need.value)
        _lib.check(self._l.b2pbr_trace_json(self._h, buf, need.value, None))
        return buf.value.decode()

    def kernelLaunches(self):
        return int(self._l.b2pbr_kernel_launches(self._h))
