/**
 * CUDA drop-in for Dplug's PBR compositor: binds libb2pbr.so (include/b2pbr.h).
 *
 * NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no D toolchain (no dmd / ldc2 / gdc / dub).
 * It depends only on declarations that exist in Dplug: ICompositor, CompositorCreationContext
 * (gui/dplug/gui/compositor.d:29-76), Mipmap / OwnedImage / ImageRef (graphics/dplug/graphics/mipmap.d,
 * image.d), box2i (math/dplug/math/box.d) and IProfiler (core/dplug/core/profiler.d).
 *
 * Install it by overriding GUIGraphics.buildCompositor (gui/dplug/gui/graphics.d:125-129):
 *
 *     override ICompositor buildCompositor(CompositorCreationContext* context)
 *     {
 *         return mallocNew!CudaPBRCompositor(context);
 *     }
 */
module dplug.gui.cudacompositor;

import dplug.core.nogc;
import dplug.core.profiler;
import dplug.math.box;
import dplug.math.vector;
import dplug.graphics;
import dplug.gui.compositor;
import dplug.window.window : WindowPixelFormat;

nothrow @nogc:

// ---- include/b2pbr.h -------------------------------------------------------------------------------------
extern(C) nothrow @nogc
{
    struct b2pbr;
    struct b2mip;
    struct b2_box2i { int min_x, min_y, max_x, max_y; }           // == box2i, math/dplug/math/box.d:29-30
    struct b2_imageref { int w, h; size_t pitch; void* pixels; }   // == ImageRef!T, image.d:127-131
    struct b2pbr_params
    {
        float[3] light1_color, light2_dir, light2_color, light3_dir, light3_color;
        float ambient_light, skybox_amount, tonemap_threshold, tonemap_ratio;
        uint pass_active_mask;
    }
    const(char)* b2_last_error();
    int b2pbr_create(b2pbr** o, int device);
    void b2pbr_destroy(b2pbr*);
    int b2pbr_resize(b2pbr*, int width, int height, int areaMaxWidth, int areaMaxHeight);
    int b2pbr_set_params(b2pbr*, const(b2pbr_params)*);
    int b2pbr_get_params(const(b2pbr)*, b2pbr_params*);
    void b2pbr_default_params(b2pbr_params*);
    int b2pbr_set_skybox(b2pbr*, const(ubyte)* rgba, int w, int h, size_t pitch);
    int b2pbr_composite_tile_host(b2pbr*, b2_imageref wfb, const(b2_box2i)* areas, int nAreas,
                                  const(b2_box2i)* updateRects, int nUpdate,
                                  b2_imageref diffuseL0, b2_imageref materialL0, b2_imageref depthL0);
    int b2_host_register(void* ptr, size_t bytes);
    int b2_host_unregister(void* ptr);
    // screenshot encoders (gui/dplug/gui/screencap.d) on the device-resident rendered buffer, ImageResizer on the device
    size_t b2pbr_screenshot_qb_size(const(b2pbr)*);
    int b2pbr_screenshot_qb(b2pbr*, int pixelFormat, void* out_, size_t cap);
    size_t b2pbr_screenshot_png_size(const(b2pbr)*);
    int b2pbr_screenshot_png(b2pbr*, int pixelFormat, void* out_, size_t cap, size_t* written);
    int b2_resize_image(int device, int kind, const(void)* in_, int inW, int inH, size_t inPitch, void* out_, int outW, int outH, size_t outPitch);
    int b2pbr_profile_enable(b2pbr*, int on);
    int b2pbr_trace_json(b2pbr*, char* buf, size_t cap, size_t* needed);

    // ---- multi-GPU row bands (one canvas cut into horizontal bands, one band object per GPU)
    struct b2band;
    int b2pbr_create_band(b2pbr** o, int device, int y0, int y1);
    b2mip* b2pbr_map(b2pbr*, int which);   // 0 diffuse, 1 material, 2 depth
    int b2mip_upload(b2mip*, int level, const(void)* host, size_t hostPitch, const(b2_box2i)* rects, int n);
    int b2pbr_download(b2pbr*, void* host, size_t hostPitch, const(b2_box2i)* rects, int n);
    int b2_band_rows(int H, int nranks, int rank, int align_, int* y0, int* y1);
    int b2_band_nccl_unique_id(void* id128);
    int b2band_create_nccl(b2band** o, b2pbr* band, int rank, int nranks, const(int)* bandY, const(void)* id128, void* ncclComm);
    int b2band_create_peers(b2band** o, const(b2pbr*)* bands, int nranks);
    void b2band_destroy(b2band*);
    int b2band_exchange_begin(b2band*);
    int b2band_exchange_end(b2band*);
    int b2band_composite_frame(b2band*);
    int b2band_sync(b2band*);
}

/// ICompositor backed by the B200 library.  Same calls, same arguments, same threading contract as
/// PBRCompositor (gui/dplug/gui/legacypbr.d:67-266); the mip levels the passes sample are regenerated on
/// the device from level 0, so the CPU-side levels GUIGraphics.regenerateMipmaps fills are simply not read
/// (a GUIGraphics subclass may override regenerateMipmaps, graphics.d:1354, to skip that CPU work).
final class CudaPBRCompositor : ICompositor
{
public:
nothrow:
@nogc:

    this(CompositorCreationContext* context, int device = 0)
    {
        // context.threadPool is not used: tiles are scheduled over the GPU's SMs
        int rc = b2pbr_create(&_h, device);
        assert(rc == 0, "b2pbr_create failed (no CUDA device? this compositor has no CPU fallback)");
        b2pbr_default_params(&_params);
    }

    ~this()
    {
        if (_h) b2pbr_destroy(_h);
    }

    /// compositor.d:43-46
    override void resizeBuffers(int width, int height, int areaMaxWidth, int areaMaxHeight)
    {
        int rc = b2pbr_resize(_h, width, height, areaMaxWidth, areaMaxHeight);
        assert(rc == 0);
    }

    /// compositor.d:63-68.  `areas` are the tiles of the PBR dirty rects grown by the update margin
    /// (graphics.d:930-936), i.e. a superset of what changed in level 0: they are uploaded, the device
    /// mips regenerated for them, the tiles composited and copied back into `wfb`.
    override void compositeTile(ImageRef!RGBA wfb, const(box2i)[] areas,
                                Mipmap!RGBA diffuseMap, Mipmap!RGBA materialMap, Mipmap!L16 depthMap,
                                IProfiler profiler)
    {
        static b2_imageref toC(T)(OwnedImage!T img)
        {
            return b2_imageref(img.w, img.h, cast(size_t) img.pitchInBytes(), cast(void*) img.scanlinePtr(0));
        }
        b2_imageref out_ = b2_imageref(wfb.w, wfb.h, wfb.pitch, cast(void*) wfb.pixels);
        static assert(box2i.sizeof == b2_box2i.sizeof);
        // the host's profiler sees the call as one span, like a pass of the reference (profiler.d:29-64); the per-stage
        // device times are available as Chrome Trace JSON (b2pbr_profile_enable + traceJSON below, same format as
        // TraceProfiler, profiler.d:177-207)
        version(Dplug_ProfileUI) if (profiler !is null) profiler.category("gpu").begin("b2pbr compositeTile");
        // update_rects = null: the tile list IS what changed (the library coalesces it: a full-frame list becomes one
        // rectangle and takes the fused mip chain and the pipelined transfer)
        int rc = b2pbr_composite_tile_host(_h, out_, cast(const(b2_box2i)*) areas.ptr, cast(int) areas.length,
                                           null, 0,
                                           toC(diffuseMap.levels[0]), toC(materialMap.levels[0]), toC(depthMap.levels[0]));
        version(Dplug_ProfileUI) if (profiler !is null) profiler.end();
        assert(rc == 0);
    }

    /// Device-side stage times of the calls since enableDeviceTrace(true), as Chrome Trace Event JSON (what TraceProfiler
    /// writes, profiler.d:177-207); `buf` receives at most buf.length bytes, the return value is the size needed.
    void enableDeviceTrace(bool on) { b2pbr_profile_enable(_h, on ? 1 : 0); }
    size_t deviceTraceJSON(char[] buf)
    {
        size_t needed;
        b2pbr_trace_json(_h, buf.ptr, buf.length, &needed);
        return needed;
    }

    /// encodeScreenshotAsQB / encodeScreenshotAsPNG (gui/dplug/gui/screencap.d:21-136) for an `onScreenshot` override
    /// (graphics.d:140-147): built on the device from the rendered frame and the depth map; free with `free(slice.ptr)`.
    /// WindowPixelFormat (window/dplug/window/window.d:155-160: BGRA8, ARGB8, RGBA8) -> B2_PF_* (RGBA8 = 0, BGRA8 = 1, ARGB8 = 2)
    static int toB2PixelFormat(WindowPixelFormat pf)
    {
        final switch (pf)
        {
            case WindowPixelFormat.RGBA8: return 0;
            case WindowPixelFormat.BGRA8: return 1;
            case WindowPixelFormat.ARGB8: return 2;
        }
    }
    ubyte[] encodeScreenshotAsQB(WindowPixelFormat pf)
    {
        import core.stdc.stdlib : malloc;
        size_t n = b2pbr_screenshot_qb_size(_h);
        ubyte* p = cast(ubyte*) malloc(n);
        if (p is null || b2pbr_screenshot_qb(_h, toB2PixelFormat(pf), p, n) != 0) return null;
        return p[0 .. n];
    }
    ubyte[] encodeScreenshotAsPNG(WindowPixelFormat pf)
    {
        import core.stdc.stdlib : malloc;
        size_t n = b2pbr_screenshot_png_size(_h), written;
        ubyte* p = cast(ubyte*) malloc(n);
        if (p is null || b2pbr_screenshot_png(_h, toB2PixelFormat(pf), p, n, &written) != 0) return null;
        return p[0 .. written];
    }

    // <LEGACY> setters, legacypbr.d:75-123
    void light1Color(vec3f c)   { _params.light1_color = [c.x, c.y, c.z]; push(); }
    void light2Dir(vec3f d)     { _params.light2_dir   = [d.x, d.y, d.z]; push(); }
    void light2Color(vec3f c)   { _params.light2_color = [c.x, c.y, c.z]; push(); }
    void light3Dir(vec3f d)     { _params.light3_dir   = [d.x, d.y, d.z]; push(); }
    void light3Color(vec3f c)   { _params.light3_color = [c.x, c.y, c.z]; push(); }
    void skyboxAmount(float a)  { _params.skybox_amount = a; push(); }
    void ambientLight(float a)  { _params.ambient_light = a; push(); }
    void tonemapThreshold(float v) { _params.tonemap_threshold = v; push(); }
    void tonemapRatio(float v)  { _params.tonemap_ratio = v; push(); }

    /// getPass(n).setActive(active), compositor.d:170-173,222-225
    void setPassActive(int pass, bool active)
    {
        if (active) _params.pass_active_mask |= (1u << pass);
        else _params.pass_active_mask &= ~(1u << pass);
        push();
    }

    /// PassSkyboxReflections.setSkybox (legacypbr.d:861-870): the pixels are copied to the device; like the
    /// reference this call takes ownership of the image and frees it.
    void setSkybox(OwnedImage!RGBA image)
    {
        int rc = b2pbr_set_skybox(_h, cast(const(ubyte)*) image.scanlinePtr(0), image.w, image.h, image.pitchInBytes());
        assert(rc == 0);
        image.destroyFree();
    }

private:
    b2pbr* _h;
    b2pbr_params _params;
    void push() { int rc = b2pbr_set_params(_h, &_params); assert(rc == 0); }
}

/// One large canvas composited by several GPUs of one box: horizontal bands of whole 64-row tiles, one band object per
/// GPU, all driven from this process (PEERS transport: halo rows are pushed GPU to GPU by the copy engines).  For one
/// process per GPU use b2band_create_nccl instead (see INTEGRATION.md section 4).  The caller owns the host images of
/// the WHOLE canvas, exactly as GUIGraphics does; every frame the bands upload their own rows, exchange the halo rows,
/// composite, and the rows come back into `wfb`.
final class CudaBandedCompositor
{
nothrow:
@nogc:
    this(int width, int height, int numDevices)
    {
        _w = width; _h = height; _n = numDevices;
        _bands = mallocSlice!(b2pbr*)(numDevices);
        _rows = mallocSlice!(int[2])(numDevices);
        foreach (r; 0 .. numDevices)
        {
            int y0, y1;
            b2_band_rows(height, numDevices, r, 64, &y0, &y1);
            _rows[r] = [y0, y1];
            int rc = b2pbr_create_band(&_bands[r], r, y0, y1);
            assert(rc == 0);
            rc = b2pbr_resize(_bands[r], width, height, 64, 64);
            assert(rc == 0);
        }
        int rc = b2band_create_peers(&_group, cast(const(b2pbr*)*) _bands.ptr, numDevices);
        assert(rc == 0, "bands thinner than the bloom reach (94 rows), or no peer access");
    }

    ~this()
    {
        if (_group) b2band_destroy(_group);
        foreach (b; _bands) if (b) b2pbr_destroy(b);
        freeSlice(_bands);
        freeSlice(_rows);
    }

    /// GUIGraphics.regenerateMipmaps + compositeGUI for the whole canvas (full redraw)
    void compositeFrame(ImageRef!RGBA wfb, OwnedImage!RGBA diffuse, OwnedImage!RGBA material, OwnedImage!L16 depth)
    {
        foreach (r; 0 .. _n)
        {
            b2_box2i own = b2_box2i(0, _rows[r][0], _w, _rows[r][1]);
            b2mip_upload(b2pbr_map(_bands[r], 0), 0, diffuse.scanlinePtr(0), diffuse.pitchInBytes(), &own, 1);
            b2mip_upload(b2pbr_map(_bands[r], 1), 0, material.scanlinePtr(0), material.pitchInBytes(), &own, 1);
            b2mip_upload(b2pbr_map(_bands[r], 2), 0, depth.scanlinePtr(0), depth.pitchInBytes(), &own, 1);
        }
        int rc = b2band_composite_frame(_group);   // enqueues every band's frame, halo exchange included
        assert(rc == 0);
        foreach (r; 0 .. _n)
        {
            b2_box2i own = b2_box2i(0, _rows[r][0], _w, _rows[r][1]);
            b2pbr_download(_bands[r], wfb.pixels, wfb.pitch, &own, 1);   // synchronises that band's stream
        }
    }

private:
    int _w, _h, _n;
    b2pbr*[] _bands;
    int[2][] _rows;
    b2band* _group;
}
