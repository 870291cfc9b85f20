/+ dub.sdl:
    name "dump_fixtures"
    dependency "dplug:gui" path="../../../reference"
    dependency "dplug:graphics" path="../../../reference"
    dependency "dplug:core" path="../../../reference"
    dependency "dplug:math" path="../../../reference"
+/
/**
    Reference-held fixtures for the B200 compositor (oracle pinning).

    Runs the REAL dplug:gui PBRCompositor and dplug:graphics Mipmap (no restatement) on level-0 maps exported by
    tools/d/export_inputs.py and writes every mip level and the composited frame as raw little-endian files that
    tools/d/import_outputs.py turns into tests/golden/*.npz with the keys tests/test_golden.py reads.  It also dumps the
    .qb screenshot of the frame and the three maps resized to 1/2 and 3/2 (SURVEY 8f rows 2 and 4), which the importer
    compares with the CPU restatement.  A maintainer with
    LDC runs (adjust the dplug path in the dub recipe above):

        python tools/d/export_inputs.py /tmp/fx
        dub run --single tools/d/dump_fixtures.d --compiler=ldc2 -- /tmp/fx
        python tools/d/import_outputs.py /tmp/fx          # rewrites tests/golden/ from the reference's own output

    This image has no D compiler (dmd, ldc2, gdc, dub absent), so the program is shipped uncompiled; it follows the call
    sequence of GUIGraphics.doResize / regenerateMipmaps / compositeGUI (gui/dplug/gui/graphics.d:1115-1137, 1338-1423).
*/
module dump_fixtures;

import core.stdc.stdio;
import core.stdc.stdlib;
import core.stdc.string;

import dplug.core.nogc;
import dplug.core.thread;
import dplug.core.profiler;
import dplug.math.box;
import dplug.graphics.color;
import dplug.graphics.image;
import dplug.graphics.mipmap;
import dplug.gui.boxlist;
import dplug.gui.compositor;
import dplug.gui.legacypbr;
import dplug.gui.screencap;
import dplug.graphics.resizer;
import dplug.window.window : WindowPixelFormat;

enum int PBR_TILE_MAX_WIDTH = 64;   // gui/dplug/gui/graphics.d:59-60
enum int PBR_TILE_MAX_HEIGHT = 64;

void readRaw(const(char)* dir, const(char)* name, void* dst, size_t rowBytes, int rows, size_t pitch)
{
    char[1024] path;
    snprintf(path.ptr, path.length, "%s/%s", dir, name);
    FILE* f = fopen(path.ptr, "rb");
    assert(f, "missing input file (run tools/d/export_inputs.py first)");
    foreach (y; 0 .. rows)
        fread(cast(ubyte*) dst + y * pitch, 1, rowBytes, f);
    fclose(f);
}

void writeRaw(const(char)* dir, const(char)* name, const(void)* src, size_t rowBytes, int rows, size_t pitch)
{
    char[1024] path;
    snprintf(path.ptr, path.length, "%s/%s", dir, name);
    FILE* f = fopen(path.ptr, "wb");
    foreach (y; 0 .. rows)
        fwrite(cast(const(ubyte)*) src + y * pitch, 1, rowBytes, f);
    fclose(f);
}

void dumpCase(const(char)* root, const(char)* caseName)
{
    char[1024] dir;
    snprintf(dir.ptr, dir.length, "%s/%s", root, caseName);
    int W, H, SW, SH;
    {
        char[1024] path;
        snprintf(path.ptr, path.length, "%s/dims.txt", dir.ptr);
        FILE* f = fopen(path.ptr, "r");
        assert(f);
        fscanf(f, "%d %d %d %d", &W, &H, &SW, &SH);
        fclose(f);
    }

    // GUIGraphics.doResize, graphics.d:1115-1137
    Mipmap!RGBA diffuseMap = mallocNew!(Mipmap!RGBA)();
    Mipmap!RGBA materialMap = mallocNew!(Mipmap!RGBA)();
    Mipmap!L16 depthMap = mallocNew!(Mipmap!L16)();
    diffuseMap.size(5, W, H);
    depthMap.size(4, W, H);
    depthMap.levels[0].size(W, H, 1, 1, 1, 2);
    materialMap.size(0, W, H);
    OwnedImage!RGBA composited = mallocNew!(OwnedImage!RGBA)();
    composited.size(W, H, 0, 16, 1, 0);

    readRaw(dir.ptr, "diffuse.raw", diffuseMap.levels[0].scanlinePtr(0), W * 4, H, diffuseMap.levels[0].pitchInBytes());
    readRaw(dir.ptr, "material.raw", materialMap.levels[0].scanlinePtr(0), W * 4, H, materialMap.levels[0].pitchInBytes());
    readRaw(dir.ptr, "depth.raw", depthMap.levels[0].scanlinePtr(0), W * 2, H, depthMap.levels[0].pitchInBytes());

    // compositor with the reference's thread pool (the output does not depend on the thread count)
    CompositorCreationContext context;
    context.threadPool = mallocNew!ThreadPool(3);
    PBRCompositor compositor = mallocNew!PBRCompositor(&context);
    compositor.resizeBuffers(W, H, PBR_TILE_MAX_WIDTH, PBR_TILE_MAX_HEIGHT);
    if (SW > 0)
    {
        OwnedImage!RGBA sky = mallocNew!(OwnedImage!RGBA)(SW, SH);
        readRaw(dir.ptr, "skybox.raw", sky.scanlinePtr(0), SW * 4, SH, sky.pitchInBytes());
        (cast(PassSkyboxReflections) compositor.getPass(PBRCompositor.PASS_SKYBOX)).setSkybox(sky);   // takes ownership
    }

    // GUIGraphics.regenerateMipmaps over the whole frame, graphics.d:1354-1423
    box2i whole = box2i(0, 0, W, H);
    {
        box2i area = whole;
        foreach (level; 1 .. diffuseMap.numLevels())
            area = diffuseMap.generateNextLevel(level == 1 ? Mipmap!RGBA.Quality.boxAlphaCovIntoPremul : Mipmap!RGBA.Quality.cubic, area, level);
        depthMap.levels[0].replicateBordersTouching(whole);
        area = whole;
        foreach (level; 1 .. depthMap.numLevels())
            area = depthMap.generateNextLevel(level >= 3 ? Mipmap!L16.Quality.cubic : Mipmap!L16.Quality.box, area, level);
    }

    // GUIGraphics.compositeGUI, graphics.d:1338-1349
    Vec!box2i rects = makeVec!box2i();
    Vec!box2i tiles = makeVec!box2i();
    rects.pushBack(whole);
    tileAreas(rects[], PBR_TILE_MAX_WIDTH, PBR_TILE_MAX_HEIGHT, tiles);
    IProfiler profiler = createProfiler();
    compositor.compositeTile(composited.toRef(), tiles[], diffuseMap, materialMap, depthMap, profiler);

    writeRaw(dir.ptr, "composited.raw", composited.scanlinePtr(0), W * 4, H, composited.pitchInBytes());
    foreach (level; 1 .. diffuseMap.numLevels())
    {
        char[64] name;
        snprintf(name.ptr, name.length, "diffuse_l%d.raw", level);
        auto L = diffuseMap.levels[level];
        writeRaw(dir.ptr, name.ptr, L.scanlinePtr(0), L.w * 4, L.h, L.pitchInBytes());
    }
    foreach (level; 1 .. depthMap.numLevels())
    {
        char[64] name;
        snprintf(name.ptr, name.length, "depth_l%d.raw", level);
        auto L = depthMap.levels[level];
        writeRaw(dir.ptr, name.ptr, L.scanlinePtr(0), L.w * 2, L.h, L.pitchInBytes());
    }
    // SURVEY 8(f) rows 2 and 4: the screenshot encoder and the resizer on the same data.
    // encodeScreenshotAsQB (gui/dplug/gui/screencap.d:21-84) on the rendered buffer of an RGBA8 window: the composited
    // frame with alpha = 255 (reorderComponents, graphics.d:1425-1452).
    {
        OwnedImage!RGBA rendered = mallocNew!(OwnedImage!RGBA)(W, H);
        foreach (y; 0 .. H)
        {
            RGBA* src = composited.scanlinePtr(y);
            RGBA* dst = rendered.scanlinePtr(y);
            foreach (x; 0 .. W)
                dst[x] = RGBA(src[x].r, src[x].g, src[x].b, 255);
        }
        ubyte[] qb = encodeScreenshotAsQB(rendered.toRef(), WindowPixelFormat.RGBA8, depthMap.levels[0].toRef());
        writeRaw(dir.ptr, "screenshot.qb", qb.ptr, qb.length, 1, qb.length);
        free(qb.ptr);
        rendered.destroyFree();
    }
    // ImageResizer (graphics/dplug/graphics/resizer.d) as PBRBackgroundGUI.resizeFun calls it (pbrbackgroundgui.d:172-199):
    // down to half size and up to 3/2.  NOTE: the result depends on DPLUG_USE_STB_IMAGE_RESIZE_V2 (stb_image_resize.d:157);
    // the B200 path implements the `false` (in-tree port) branches — build dplug:graphics with that enum flipped to pin it.
    {
        ImageResizer resizer;
        static immutable int[2][2] ratios = [[1, 2], [3, 2]];
        foreach (k, ratio; ratios)
        {
            const int ow = W * ratio[0] / ratio[1], oh = H * ratio[0] / ratio[1];
            OwnedImage!RGBA d2 = mallocNew!(OwnedImage!RGBA)(ow, oh), m2 = mallocNew!(OwnedImage!RGBA)(ow, oh);
            OwnedImage!L16 z2 = mallocNew!(OwnedImage!L16)(ow, oh);
            resizer.resizeImageDiffuse(diffuseMap.levels[0].toRef(), d2.toRef());
            resizer.resizeImageMaterial(materialMap.levels[0].toRef(), m2.toRef());
            resizer.resizeImageDepth(depthMap.levels[0].toRef(), z2.toRef());
            char[64] name;
            snprintf(name.ptr, name.length, "resized%d_diffuse.raw", cast(int) k);
            writeRaw(dir.ptr, name.ptr, d2.scanlinePtr(0), ow * 4, oh, d2.pitchInBytes());
            snprintf(name.ptr, name.length, "resized%d_material.raw", cast(int) k);
            writeRaw(dir.ptr, name.ptr, m2.scanlinePtr(0), ow * 4, oh, m2.pitchInBytes());
            snprintf(name.ptr, name.length, "resized%d_depth.raw", cast(int) k);
            writeRaw(dir.ptr, name.ptr, z2.scanlinePtr(0), ow * 2, oh, z2.pitchInBytes());
            d2.destroyFree(); m2.destroyFree(); z2.destroyFree();
        }
    }
    printf("%s: %d x %d done\n", caseName, W, H);
}

int main(string[] args)
{
    assert(args.length >= 2, "usage: dump_fixtures <dir written by export_inputs.py>");
    const(char)* root = args[1].ptr;   // dub passes zero-terminated argv strings through
    char[1024] path;
    snprintf(path.ptr, path.length, "%s/cases.txt", root);
    FILE* f = fopen(path.ptr, "r");
    assert(f);
    char[256] name;
    while (fscanf(f, "%255s", name.ptr) == 1)
        dumpCase(root, name.ptr);
    fclose(f);
    return 0;
}
