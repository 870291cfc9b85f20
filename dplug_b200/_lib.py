"""ctypes binding of libb2pbr.so (include/b2pbr.h).

The library is CUDA-only: importing this module never falls back to a CPU implementation.  If the
shared object is missing the import fails loudly; if no GPU is present, object creation raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb2pbr.so")


class B2Error(RuntimeError):
    pass


class box2i(C.Structure):
    """box2i (math/dplug/math/box.d:29-30)."""
    _fields_ = [("min_x", C.c_int), ("min_y", C.c_int), ("max_x", C.c_int), ("max_y", C.c_int)]

    def astuple(self):
        return (self.min_x, self.min_y, self.max_x, self.max_y)


class imageref(C.Structure):
    """ImageRef!T (graphics/dplug/graphics/image.d:127-131)."""
    _fields_ = [("w", C.c_int), ("h", C.c_int), ("pitch", C.c_size_t), ("pixels", C.c_void_p)]


class params(C.Structure):
    _fields_ = [("light1_color", C.c_float * 3), ("light2_dir", C.c_float * 3), ("light2_color", C.c_float * 3),
                ("light3_dir", C.c_float * 3), ("light3_color", C.c_float * 3), ("ambient_light", C.c_float),
                ("skybox_amount", C.c_float), ("tonemap_threshold", C.c_float), ("tonemap_ratio", C.c_float),
                ("pass_active_mask", C.c_uint32)]


# name -> (restype, argtypes); every symbol include/b2pbr.h declares
SIGNATURES = {
    "b2_last_error": (C.c_char_p, []),
    "b2_abi_version": (C.c_int, []),
    "b2_device_count": (C.c_int, []),
    "b2mip_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "b2mip_create_band": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "b2mip_destroy": (None, [C.c_void_p]),
    "b2mip_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "b2mip_num_levels": (C.c_int, [C.c_void_p]),
    "b2mip_width": (C.c_int, [C.c_void_p]),
    "b2mip_height": (C.c_int, [C.c_void_p]),
    "b2mip_level": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(imageref), C.POINTER(C.c_int)]),
    "b2mip_upload": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(box2i), C.c_int]),
    "b2mip_download": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(box2i), C.c_int]),
    "b2mip_generate_next_level": (C.c_int, [C.c_void_p, C.c_int, box2i, C.c_int, C.POINTER(box2i)]),
    "b2mip_generate_level": (C.c_int, [C.c_void_p, C.c_int, C.c_int, box2i]),
    "b2mip_generate_mipmaps": (C.c_int, [C.c_void_p, C.c_int]),
    "b2mip_generate_chain": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "b2mip_generate_levels": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), box2i, C.POINTER(box2i)]),
    "b2_set_mip_fusion": (C.c_int, [C.c_int]),
    "b2_set_mip_wave": (C.c_int, [C.c_int]),
    "b2_mip_wave_plan": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), box2i, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int)]),
    "b2_dirty_rects_create": (C.c_int, [C.POINTER(C.c_void_p)]),
    "b2_dirty_rects_destroy": (None, [C.c_void_p]),
    "b2_dirty_rects_add": (C.c_int, [C.c_void_p, box2i]),
    "b2_dirty_rects_is_empty": (C.c_int, [C.c_void_p]),
    "b2_dirty_rects_pull_all": (C.c_int, [C.c_void_p, C.POINTER(box2i), C.c_int]),
    "b2_have_no_overlap": (C.c_int, [C.POINTER(box2i), C.c_int]),
    "b2mip_sample": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "b2mip_sync": (C.c_int, [C.c_void_p]),
    "b2mip_kernel_launches": (C.c_uint64, [C.c_void_p]),
    "b2_impact_on_next_level": (box2i, [C.c_int, box2i, C.c_int, C.c_int]),
    "b2_tile_areas": (C.c_int, [C.POINTER(box2i), C.c_int, C.c_int, C.c_int, C.POINTER(box2i), C.c_int]),
    "b2_remove_overlapping_areas": (C.c_int, [C.POINTER(box2i), C.c_int, C.POINTER(box2i), C.c_int]),
    "b2_grow_rect": (box2i, [box2i, C.c_int, C.c_int, C.c_int]),
    "b2pbr_default_params": (None, [C.POINTER(params)]),
    "b2pbr_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "b2pbr_create_band": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int]),
    "b2pbr_destroy": (None, [C.c_void_p]),
    "b2pbr_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "b2pbr_resize": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]),
    "b2pbr_set_params": (C.c_int, [C.c_void_p, C.POINTER(params)]),
    "b2pbr_get_params": (C.c_int, [C.c_void_p, C.POINTER(params)]),
    "b2pbr_set_mode": (C.c_int, [C.c_void_p, C.c_int]),
    "b2pbr_set_skybox": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t]),
    "b2pbr_map": (C.c_void_p, [C.c_void_p, C.c_int]),
    "b2pbr_skybox": (C.c_void_p, [C.c_void_p]),
    "b2pbr_output": (C.c_int, [C.c_void_p, C.POINTER(imageref)]),
    "b2pbr_regenerate_mipmaps": (C.c_int, [C.c_void_p, C.POINTER(box2i), C.c_int]),
    "b2pbr_regenerate_mipmaps_on": (C.c_int, [C.c_void_p, C.POINTER(box2i), C.c_int, C.c_void_p]),
    "b2pbr_composite_tile": (C.c_int, [C.c_void_p, C.POINTER(box2i), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "b2pbr_redraw": (C.c_int, [C.c_void_p, C.POINTER(box2i), C.c_int, C.POINTER(box2i), C.c_int]),
    "b2pbr_composite_tile_host": (C.c_int, [C.c_void_p, imageref, C.POINTER(box2i), C.c_int, C.POINTER(box2i), C.c_int,
                                            imageref, imageref, imageref]),
    "b2pbr_composite_tile_host_begin": (C.c_int, [C.c_void_p, imageref, C.POINTER(box2i), C.c_int, C.POINTER(box2i), C.c_int,
                                                  imageref, imageref, imageref]),
    "b2pbr_composite_tile_host_end": (C.c_int, [C.c_void_p]),
    "b2pbr_download": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(box2i), C.c_int]),
    "b2pbr_download_swizzled": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(box2i), C.c_int]),
    "b2_band_need_rows": (C.c_int, [C.c_int] * 6 + [C.POINTER(C.c_int)] * 2),
    "b2pbr_band_rows": (C.c_int, [C.c_void_p, C.c_int, C.c_int] + [C.POINTER(C.c_int)] * 4),
    "b2_band_rows": (C.c_int, [C.c_int] * 4 + [C.POINTER(C.c_int)] * 2),
    "b2_band_plan": (C.c_int, [C.c_int] * 5 + [C.POINTER(box2i)] + [C.POINTER(C.c_int)] * 3),
    "b2_band_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "b2band_create_nccl": (C.c_int, [C.POINTER(C.c_void_p), C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_void_p, C.c_void_p]),
    "b2band_create_peers": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int]),
    "b2band_destroy": (None, [C.c_void_p]),
    "b2band_exchange_begin": (C.c_int, [C.c_void_p]),
    "b2band_exchange_end": (C.c_int, [C.c_void_p]),
    "b2band_composite_frame": (C.c_int, [C.c_void_p]),
    "b2band_sync": (C.c_int, [C.c_void_p]),
    "b2band_halo_bytes_per_frame": (C.c_uint64, [C.c_void_p]),
    "b2band_info": (C.c_int, [C.c_void_p, C.c_int] + [C.POINTER(C.c_int)] * 5),
    "b2pbr_graph_begin": (C.c_int, [C.c_void_p]),
    "b2pbr_graph_end": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "b2pbr_graph_launch": (C.c_int, [C.c_void_p, C.c_int]),
    "b2layer_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int]),
    "b2layer_destroy": (None, [C.c_void_p]),
    "b2layer_upload": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(box2i), C.c_int]),
    "b2pbr_draw_layer": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(box2i), C.c_int, C.c_int]),
    "b2_resize_image": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_int, C.c_int, C.c_size_t]),
    "b2_resize_axis_table": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_float), C.c_int]),
    "b2layer_resize_plane": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_size_t]),
    "b2pbr_screenshot_qb_size": (C.c_size_t, [C.c_void_p]),
    "b2pbr_screenshot_qb": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]),
    "b2pbr_screenshot_rgba": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]),
    "b2pbr_screenshot_png_size": (C.c_size_t, [C.c_void_p]),
    "b2pbr_screenshot_png": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "b2pbr_present": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(box2i), C.c_int, C.c_int]),
    "b2pbr_set_color_correction": (C.c_int, [C.c_void_p, C.c_void_p]),
    "b2_lift_gamma_gain_contrast_table": (None, [C.POINTER(C.c_float), C.c_void_p]),
    "b2pbr_sync": (C.c_int, [C.c_void_p]),
    "b2pbr_composite_raw": (C.c_int, [C.c_void_p, C.POINTER(box2i), C.c_int]),
    "b2pbr_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "b2pbr_trace_json": (C.c_int, [C.c_void_p, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "b2pbr_kernel_launches": (C.c_uint64, [C.c_void_p]),
    "b2_host_register": (C.c_int, [C.c_void_p, C.c_size_t]),
    "b2_host_unregister": (C.c_int, [C.c_void_p]),
}

_lib = None


def lib():
    """Loads libb2pbr.so (once).  Raises if it has not been built: there is no other implementation."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B2Error("%s not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(the CUDA library is the only implementation of this package)" % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc):
    if rc != 0:
        raise B2Error("b2pbr error %d: %s" % (rc, lib().b2_last_error().decode()))


class BoxArray:
    """A rectangle list already marshalled for the C ABI (reuse it across frames: tile lists are long)."""

    def __init__(self, rects):
        self.n = len(rects)
        self.arr = boxes(rects)
        self.rects = list(rects)

    def __len__(self):
        return self.n

    def __iter__(self):
        return iter(self.rects)


def boxes(rects):
    """list of (minx,miny,maxx,maxy) -> ctypes array of box2i"""
    if isinstance(rects, BoxArray):
        return rects.arr
    arr = (box2i * max(1, len(rects)))()
    for i, r in enumerate(rects):
        arr[i] = r if isinstance(r, box2i) else box2i(*[int(v) for v in r])
    return arr
