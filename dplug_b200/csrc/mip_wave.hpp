// Host / device interface of the whole-pyramid wavefront kernel (mip_wave.cu).
#pragma once
#include <vector>
#include "common.cuh"

namespace b2 {

constexpr int kWaveMaxSteps = 12;     // Mipmap(12) is the deepest pyramid of the path (the skybox, legacypbr.d:868)

enum { WAVE_BOX_RGBA = 0, WAVE_BOX_PREMUL = 1, WAVE_CUBIC_RGBA = 2, WAVE_BOX_L16 = 3, WAVE_CUBIC_L16 = 4 };

struct alignas(16) WaveItem {         // read by the kernel as two 128-bit loads
    int stepKind;                     // step (index of the generated level inside the chain) | body kind << 16
    int by;                           // virtual block row (8 destination rows) inside the step's update rectangle
    int bx0, bx1;                     // virtual block columns [bx0, bx1)
    int depLo, depHi;                 // completion counters (indices into ctl) of the source row chunks; none if depHi < depLo
    int depTarget;                    // value they must reach: the number of items of a source row chunk
    int done;                         // the counter this item's row chunk owns
};
static_assert(sizeof(WaveItem) == 32, "WaveItem is two int4");

struct WaveArgs {
    DevLevel lv[kWaveMaxSteps + 1];   // lv[0] = the chain's source level, lv[k + 1] = the level step k writes
    Box upd[kWaveMaxSteps];           // update rectangle of step k (generateLevel's `updateRect`)
    const WaveItem* items;
    int nItems;
    int* ctl;                         // [0] queue head, [1] CTAs that have left, [2 ..] row-chunk counters; all zero between launches
    int nCtl;
    unsigned long long nz;            // (-0.0f, -0.0f): see mip_math.cuh, mulToAdd2
};

// items in launch order; returns 0, or -1 when the chain cannot be planned
int wave_build_plan(const DevLevel* lv, const Box* upd, const int* q, int n, bool rgba, std::vector<WaveItem>& items, int* nCtl);
void launch_mip_wave(const WaveArgs& a, cudaStream_t s);

}  // namespace b2
