// Device-side interface between the C ABI (capi.cu) and the sm_100a kernels (kernels.cu).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace iqo_b200 {

// One axis as the kernels see it (device pointers; layout documented in plan.hpp).
struct AxisDev {
    const int32_t *first;  // [D]
    const int32_t *row;    // [D]
    const int32_t *coef;   // [numRows][N]
    const int32_t *deno;   // [numRows]
    int N;
    int S;
    int D;
};

// One launch = destination rows [dstRow0, dstRow0+dstRows) of nFrames frames.
struct ResizeArgs {
    AxisDev x, y;
    const uint8_t *src;        // source buffer row 0 == global source row srcRow0
    uint8_t *dst;              // addresses global destination row dstRow0
    long long srcPitch, dstPitch;
    long long srcFrameStride, dstFrameStride;
    int nFrames;
    int srcRow0, srcRows;      // rows of the full image present in `src` (row-band mode; else 0, S)
    int dstRow0, dstRows;
    int shift;                 // 20 (Lanczos) or 23 (Area, Linear)
    int lanczos;               // border rows/columns use the truncating divisions
    int workSigned;
};

// Tile geometry of the generic kernel, chosen by the host from the plan.
struct GenericGeom {
    int tileW, tileH;  // destination tile
    int workW;         // widest source-column window of any tile (elements)
    size_t smemBytes;
};

GenericGeom chooseGenericGeom(const int32_t *firstX, int N, int S, int D);
cudaError_t launchGeneric(const ResizeArgs &a, const GenericGeom &g, cudaStream_t stream);
cudaError_t initKernels();  // sets function attributes once per device

unsigned long long launchCount();

}  // namespace iqo_b200
