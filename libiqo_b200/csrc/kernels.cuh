// Device-side interface between the C ABI (capi.cu) and the sm_100a kernels (kernels.cu).
#pragma once

#include <cuda.h>
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

// Arguments of the specialised 2:1 x 2:1 single-phase Lanczos kernel (see plan.hpp HalfPlan).
struct HalfArgs {
    const uint8_t *src;
    uint8_t *dst;
    long long srcPitch, dstPitch, srcFrameStride, dstFrameStride;
    int SW, SH, DW, DH;
    int nFrames;
    int tileRows;             // destination rows per tile (even, <= 64)
    int tileShift;            // tiles start at 120*i - tileShift destination columns (0, or 4 for TMA)
    int dstVec;               // destination rows may be written with 8-byte stores
    // streaming variant (plan.hpp HalfPlan, s* fields): qmin / NG / cwY / borderY / borderX hold its tables
    int bandPairs;            // destination row pairs per warp
    int delta;                // source-row groups start at rows 4g + delta
    int zmask;                // zero main-phase coefficient words the kernel skips
    int NXH, skipHi0;
    uint32_t cwXo[6];
    // vertical
    int qmin, NG;
    uint32_t cwY[2][3];
    const uint32_t *borderY;  // [numRowsY][3]
    const int32_t *rowY;      // [DH]
    const int32_t *denoY;     // [numRowsY]
    int mbY, meY;
    int workBias;
    // horizontal
    int NWX, symmetric;
    int endsHi;               // symmetric form: the end-tap pair word has a non-zero high byte plane
    uint32_t cwX[7];
    uint32_t cwXs[4];
    int accInit;
    int mbX, meX;
    const uint32_t *magicY;   // [numRowsY] multiply-high constants of the border-row divisions
    const int32_t *borderX;   // [border columns][9]: 7 pair words, denominator*64, accumulator init
    int NX;
    uint32_t zero;            // always 0 (see the pair sums in resizeHalfKernel)
};

// tmap == NULL: source rows are read with global loads; otherwise a 3-D tensor map (x, y, frame)
// over the source frames with box 256 x boxRows x 1 (at most 65535 frames per launch either way).
cudaError_t launchHalf(const HalfArgs &a, const CUtensorMap *tmap, int boxRows, cudaStream_t stream);
int halfSourceRowsMax();
// Streaming variant: a warp per (120-pixel column strip, band of a.bandPairs row pairs, frame).
// Needs 16-byte aligned source rows and SW % 8 == 0; at most 65535 frames and bands per launch.
// tmap != NULL: the source FIFO is fed by TMA through a 3-D tensor map over a 16-bit view of the frames
// (x / 2, y, frame) with box 136 x halfStreamBoxRows(NG) x 1; NULL: by cp.async chunks.
cudaError_t launchHalfStream(const HalfArgs &a, const CUtensorMap *tmap, cudaStream_t stream);
int halfStreamBoxRows(int NG);
bool halfStreamHasKernel(int NG, int NXH);

// Arguments of the general packed kernel (see plan.hpp PackedPlan).
struct PackedArgs {
    const uint8_t *src;
    uint8_t *dst;
    long long srcPitch, dstPitch, srcFrameStride, dstFrameStride;
    int SW, SH, DW, DH;
    int nFrames;
    int srcRow0, dstRow0, dstRows;  // row-band mode (else 0, 0, DH)
    int tileW, tileH, wordsPerRow;  // destination tile and the 32-bit words of one shared W row
    int shift;                      // 20 or 23
    int isSigned;                   // Lanczos: signed high coefficient plane, border divisions
    int workBias;
    // vertical
    const int32_t *firstY, *ntapY, *coefOffY, *coefY, *rowY, *denoY;
    const uint32_t *magicY;
    int ntMax;                      // widest vertical slice (row stride of the staged coefficients)
    // horizontal: per destination column {first source column, offset of its NP coefficient pair
    // words in cwX, accumulator init, divisor (0 = shift)}
    const int4 *recX;
    const uint32_t *cwX;
    int NX, NP;                     // NP is already padded to an instantiated size
};

struct PackedGeom {
    int tileW, tileH, wordsPerRow;
    size_t smemBytes;
};
int packedPadNP(int np);  // smallest instantiated pair-word count >= np, 0 if none
PackedGeom choosePackedGeom(const int32_t *firstXClamped, int N, int S, int D, int npt, int ntMax);
cudaError_t launchPacked(const PackedArgs &a, cudaStream_t stream);

// Area 2:1 x 2:1 streaming kernel (two-tap single-phase tables, coefficients < 256 vertically).
// Needs 16-byte aligned source rows and an even-width multiple of 16 source; at most 65535 frames.
cudaError_t launchArea2(const uint8_t *src, uint8_t *dst, long long srcPitch, long long dstPitch, long long srcFrameStride,
                        long long dstFrameStride, int DW, int DH, int nFrames, const int32_t cy[2], const int32_t cx[2],
                        cudaStream_t stream);

// Linear up-sampling by K = 2 or 3 on X (any Linear ratio on Y): streaming kernel.  Needs SW % 4 == 0 and
// 4-byte aligned source / destination rows; DH <= 65535, nFrames <= 65535.
// Area reductions at 3:2, 4:3, 2:1, 5:2, 3:1, 4:1 on X (streaming, no shared memory)
bool areaDownHasKernel(int RS, int RD, int NX);
int areaDownItemColumns(int RS, int RD);
cudaError_t launchAreaDown(int RS, int RD, int NX, int NXeff, const uint8_t *src, uint8_t *dst, long long srcPitch, long long dstPitch,
                           long long srcFrameStride, long long dstFrameStride, int SW, int SH, int DW, int DH, int nFrames, int NY,
                           int NYeff, const int32_t *firstY, const int32_t *rowY, const int32_t *coefY, const int32_t *cx,
                           cudaStream_t stream);
cudaError_t launchLinearUp(int RS, int RD, const uint8_t *src, uint8_t *dst, long long srcPitch, long long dstPitch,
                           long long srcFrameStride, long long dstFrameStride, int SW, int SH, int DW, int DH, int nFrames,
                           const int32_t *firstY, const int32_t *rowY, const int32_t *coefY, const int q1X[8],
                           cudaStream_t stream);

// Arguments of the streaming 2:1 small-kernel path (plan.hpp SmallPlan).
struct SmallArgs {
    const uint8_t *src;
    uint8_t *dst;
    long long srcPitch, dstPitch, srcFrameStride, dstFrameStride;
    int SW, SH, DW, DH;
    int nFrames;
    int TY, cy0, NW, wbase;
    int32_t cY[4];
    uint32_t cwX[3];
    int accInit, workBias;
    int mbX, meX, mbY, meY;
    const int32_t *rowsY;      // [numRowsY][4]
    const uint32_t *magicY;    // [numRowsY]
    AxisDev gx, gy;            // generic tables: border pixels are recomputed from them
};
cudaError_t launchSmall(const SmallArgs &a, cudaStream_t stream);

// Arguments of the rational-ratio streaming kernel (plan.hpp RatioPlan).
struct RatioArgs {
    const uint8_t *src;
    uint8_t *dst;
    long long srcPitch, dstPitch, srcFrameStride, dstFrameStride;
    int SW, SH, DW, DH;
    int nFrames;
    int RS, RD, NX;
    int tailZeros;             // the last tap of every phase is zero (the kernel skips it)
    int bandRows;              // destination rows per warp (multiple of 8)
    int groupsPerStrip, c0;
    int workBias, accInit;
    int dstVec;                // destination rows may be written with 8-byte stores
    const int32_t *rowRec;     // [DH][8]
    uint32_t cwX[4][2][7];     // [phase][parity][pair word]
    int odd;                   // W pairs are columns (2m-1, 2m): first[0] is odd
    int mbX, meX;
    AxisDev gx, gy;            // generic tables: border columns are recomputed from them
};
bool ratioHasKernel(int RS, int RD, int NX, int odd);
cudaError_t launchRatio(const RatioArgs &a, cudaStream_t stream);

// Arguments of the general Lanczos streaming kernel (plan.hpp LStreamPlan + PackedPlan's horizontal tables).
struct LStreamArgs {
    const uint8_t *src;            // buffer row 0 == global source row srcRow0
    uint8_t *dst;                  // addresses global destination row dstRow0
    long long srcPitch, dstPitch, srcFrameStride, dstFrameStride;
    int SW, DW;
    int nFrames;
    int srcRow0, srcRows, dstRow0, dstRows;
    int bandRows;                  // destination rows per warp (multiple of 8)
    int stripW;
    int workBias;
    const int32_t *rowRec;         // [DH][16]
    const int4 *recX;              // PackedPlan: {first column, coefficient word offset, accumulator init, divisor}
    const uint32_t *cwX;
    int NP;
};
bool lstreamHasKernel(int NP);
cudaError_t launchLStream(const LStreamArgs &a, cudaStream_t stream);

// Arguments of the tensor-path Lanczos kernel (plan.hpp MmaPlan).  The source is read through a 3-D tensor map over a
// 16-bit view of the frames (x / 2, y, frame) with box (wcols / 2) x 16 x 1.
struct MmaArgs {
    uint8_t *dst;                  // addresses global destination row dstRow0
    long long dstPitch, dstFrameStride;
    int DW;
    int srcRow0;                   // global source row of the tensor map's row 0 (row-band mode; else 0)
    int dstRow0, dstRows;          // destination rows of this launch (global indices)
    int bandBlocks;                // 16-row blocks per warp
    int stripTiles, wcols;
    int vKMax, hKMax;
    int nChunks;                   // 8-row chunks of the source FIFO (the plan's: the row map is built for it)
    int warps;                     // warps per CTA (1, 2 or 4): they share the strip's FIFO / W / tables
    int workBias;
    int mbY, meY, mbX, meX;
    int dstVec;                    // destination rows can take 16-byte (2) / 8-byte (1) stores, else 0
    const int2 *vBlock;
    const uint4 *vFrag;
    const int2 *vRow;
    const int32_t *vRowMap;        // [blocks][vKMax * 32]: FIFO byte offset of the source row of every k slot
    int isSigned;                  // Lanczos (signed coefficients, 20-bit shift, border divisions) or Area / Linear
    const int32_t *stripXs;
    const int2 *hTile;
    const uint4 *hFrag;
    const int2 *hCol;
    int nFrames;
};
size_t mmaRatioSmemBytes(int wcols, int nChunks);
bool mmaRatioHasKernel(int RS, int RD, int NX, int odd);
size_t mmaSmemBytes(int wcols, int stripTiles, int nChunks, int hKMax);
cudaError_t launchMma(const MmaArgs &a, const CUtensorMap &tmap, cudaStream_t stream);
// tensor-path vertical pass + the 3:2 kernel's compile-time dp2a horizontal pass (cfg1)
cudaError_t launchMmaRatio(const MmaArgs &a, const CUtensorMap &tmap, const RatioArgs &ra, cudaStream_t stream);

// Float ("SIMD-semantics") mode, plan.hpp FloatPlan: the generic tile organisation with float tables and a float work tile.
struct FloatArgs {
    AxisDev x, y;                  // first / row maps and sizes of the integer plan (coef / deno unused)
    const float *coefX, *coefY;    // [numRows][N]
    const float *denoX, *denoY;    // [numRows], 0: no division
    const uint8_t *src;
    uint8_t *dst;
    long long srcPitch, dstPitch, srcFrameStride, dstFrameStride;
    int nFrames;
    int srcRow0, srcRows, dstRow0, dstRows;
};
cudaError_t launchFloat(const FloatArgs &a, const GenericGeom &g, cudaStream_t stream);

GenericGeom chooseGenericGeom(const int32_t *firstX, int N, int S, int D);
cudaError_t launchGeneric(const ResizeArgs &a, const GenericGeom &g, cudaStream_t stream);
cudaError_t initKernels();  // sets function attributes once per device

unsigned long long launchCount();

}  // namespace iqo_b200
