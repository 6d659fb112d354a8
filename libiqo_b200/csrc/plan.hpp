// Host-side planner: everything the reference computes in *ResizerImpl<Generic>::init
// plus the per-destination-index maps its resize() derives with LinearIterator,
// re-derived in closed form.  Pure C++, no CUDA.
#pragma once

#include <stddef.h>
#include <stdint.h>

#include <string>
#include <vector>

namespace iqo_b200 {

enum Kind { kLanczos = 0, kArea = 1, kLinear = 2 };

enum PlanError {
    kPlanOk = 0,
    kPlanBadArg = -1,       // zero size / degree / pxScale (the reference divides by zero in gcd)
    kPlanUnsupported = -2,  // Lanczos image smaller than its kernel (reference iterators desynchronise)
    kPlanDegenerate = -3,   // a border denominator is 0 or a coefficient leaves int16 (reference: SIGFPE / UB)
    kPlanTooLarge = -4      // sizes beyond the 32-bit index maps of this implementation
};

// One axis (X: along a row, Y: across rows) of a resampler.
//
// Every destination index d of the axis is described by
//   first[d] : source index of tap 0 (may lie outside [0,S): such taps carry coefficient 0
//              in every case the reference defines; loads clamp to the edge)
//   row[d]   : which coefficient row to use.  Rows [0, rD) are the phase tables of the
//              reference (d mod rD).  Rows >= rD are position-specific rows appended by the
//              planner: Lanczos border rows (out-of-range taps zeroed, denominators recorded)
//              and Linear's replicated first/last index.
//   deno[r]  : 0 for ordinary rows.  For Lanczos border rows the sum of the in-range
//              coefficients, used by the truncating divisions of resizeYborder /
//              resizeXborder (src/IQOLanczosResizerImpl_Generic.cpp:488,572).
struct AxisPlan {
    int64_t S, D;          // source / destination length
    int64_t rS, rD;        // gcd-reduced
    int N;                 // taps per row
    int bias;              // fixed-point one of this axis (64/256 for Y, 16384/32768 for X)
    bool identity;         // S == D: pass-through (one tap of weight `bias`)
    int64_t mainBegin, mainEnd;
    int numRows;
    std::vector<int32_t> coef;   // numRows x N
    std::vector<int32_t> deno;   // numRows
    std::vector<int32_t> first;  // D
    std::vector<int32_t> row;    // D
    int32_t coefMin, coefMax;    // over all rows
    int32_t posSumMax, negSumMin; // max over rows of the sum of positive / negative coefficients
};

struct Plan {
    Kind kind;
    unsigned degree;
    size_t pxScale;
    AxisPlan x, y;
    int shift;        // total fixed-point bits removed after the X pass (20 Lanczos, 23 Area/Linear)
    bool workSigned;  // the vertical pass stores int16 (Lanczos) or uint16 (Area, Linear)
    std::string error;
};

// Builds the plan; returns a PlanError and fills plan.error with a message on failure.
int buildPlan(Plan &plan, int kind, unsigned degree, size_t srcW, size_t srcH, size_t dstW, size_t dstH, size_t pxScale);

// number of taps (exported for the tests)
int lanczosNumCoefs(int degree, uint64_t rS, uint64_t rD, uint64_t pxScale);
int areaNumCoefs(uint64_t rS, uint64_t rD);

}  // namespace iqo_b200

namespace iqo_b200 {

// Parameters of the specialised kernel for 2:1 down-sampling on both axes with single-phase
// Lanczos tables (BASELINE configs 3 and 4).  See kernels.cu (resizeHalfKernel) for the layout.
struct HalfPlan {
    bool eligible;
    std::string why;          // reason when not eligible
    // vertical pass: u8 x s8 dp4a over groups of 4 source rows aligned to multiples of 4
    int qmin;                 // destination row pair k uses source-row groups k+qmin .. k+qmin+NG-1
    int NG;                   // 1..3
    uint32_t cwY[2][3];       // [row parity][group] packed s8 coefficients of the main phase
    std::vector<uint32_t> borderY;  // [numRowsY][3] packed words of every coefficient row (row 0 = main, unused)
    int workBias;             // added to every intermediate so that it is a non-negative u16
    // horizontal pass: dp2a over "natural" 16-bit pairs (columns 2m, 2m+1)
    int wa;                   // destination column d uses pair words d+wa .. d+wa+NWX-1
    int NWX;                  // N/2 + 1  (<= 7)
    bool symmetric;           // palindromic table: mirrored pairs are pre-added
    uint32_t cwX[7];          // general form, per pair word: bytes (lo(c_a), lo(c_b), hi(c_a), hi(c_b))
    uint32_t cwXs[4];         // symmetric form: [0..m-2] summed pairs, [m-1] centre, [m] the two end taps
    int accInit;              // rounding constant minus the bias contribution
    // border rows: floor(2^32 / deno) + 1 per coefficient row (0 for ordinary rows): the truncating
    // division by the small denominator becomes a multiply-high (exact for |numerator| <= 2^21)
    std::vector<uint32_t> magicY;
    // border columns in order (left ones, then right ones), 9 words each:
    // [0..6] pair words of the masked taps (general form), [7] denominator * 64, [8] accumulator init
    std::vector<int32_t> borderX;

    // ---- streaming variant (kernels.cu: resizeHalfStreamKernel) ----
    // vertical: the 4-row groups start at source rows 4g + sDelta; the offset is chosen so that as
    // few coefficient words as possible are non-zero (Lanczos2: 5 instead of 6)
    bool sEligible;
    int sDelta, sQmin, sNG;
    uint32_t sCwY[2][3];
    int sZ;                   // bit (parity * 3 + group): that main-phase word is zero and skipped by the kernel
    std::vector<uint32_t> sBorderY;  // [numRowsY][3]
    // horizontal: 16-bit pairs (columns 2m-1, 2m), so that the NX taps of a pixel fill exactly NX/2
    // pair words: word i holds taps 2i (low half) and 2i+1 (high half)
    int NXH;                  // NX / 2 (2, 4 or 6)
    uint32_t cwXo[6];         // bytes (lo(c_2i), lo(c_2i+1), hi(c_2i), hi(c_2i+1)); symmetric tables use the first NXH/2
    bool skipHi0;             // symmetric form: both taps of word 0 fit the low byte plane
    std::vector<int32_t> borderXo;   // [border columns][8]: 6 pair words, denominator * 64, accumulator init
};

void buildHalfPlan(const Plan &plan, HalfPlan &h);

}  // namespace iqo_b200

namespace iqo_b200 {

// Parameters of the general "packed" kernel (kernels.cu: resizePackedKernel): any kind, ratio and
// phase count, as long as no out-of-image tap carries a non-zero coefficient (true for every
// input the reference defines) and the vertical sums fit 16-bit lanes.
struct PackedPlan {
    bool eligible;
    std::string why;
    int workBias;                   // added to every intermediate: lanes stay in [0, 65535]
    // vertical pass, per destination row: taps [0, ntapY) of coefficient slice coefOffY apply to
    // source rows firstY .. firstY+ntapY-1 (all inside the image)
    std::vector<int32_t> firstY, ntapY, coefOffY;
    std::vector<uint32_t> magicY;   // per coefficient row: multiply-high constant of the border division
    // horizontal pass, per destination column: first source column (>= 0) and coefficient row
    std::vector<int32_t> firstX;
    int NP;                         // pair words per packed row (padded to `padNP`)
    int ntMax;                      // longest vertical slice
    std::vector<uint32_t> cwX;      // [numRowsX][2 (parity of the window start)][NP]: bytes lo_a, lo_b, hi_a, hi_b
    // per destination column: {first source column, offset of its row/parity in cwX,
    // rounding constant minus bias * (sum of the row), 0 or denominator * 64 (Lanczos border)}
    std::vector<int32_t> recX;      // [D][4]
};

// padNP: pair-word count the kernel is instantiated for (>= N/2 + 1)
void buildPackedPlan(const Plan &plan, PackedPlan &q, int padNP);

}  // namespace iqo_b200

namespace iqo_b200 {

// 2:1 x 2:1 single-phase Lanczos whose kernel has at most four non-zero taps per axis (the
// pxScale=2 chroma planes of YUV420, Lanczos1): streamed without shared memory
// (kernels.cu: resizeHalfSmallKernel).
struct SmallPlan {
    bool eligible;
    std::string why;
    int TY, cy0;                 // destination row y reads source rows 2y + cy0 .. 2y + cy0 + TY - 1
    int32_t cY[4];
    std::vector<int32_t> rowsY;  // [numRowsY][4]: every coefficient row at those tap positions (border rows masked)
    std::vector<uint32_t> magicY;
    int TX, cx0;                 // destination column d reads source columns 2d + cx0 .. 2d + cx0 + TX - 1
    int NW, wbase;               // ... = pair words d + wbase .. d + wbase + NW - 1
    uint32_t cwX[3];             // bytes lo_a, lo_b, hi_a, hi_b per pair word
    int accInit, workBias;
};

void buildSmallPlan(const Plan &plan, SmallPlan &s);

}  // namespace iqo_b200

namespace iqo_b200 {

// Lanczos at a rational ratio whose destination period divides 8 (3:2, 1:2, 3:4 ...) with a horizontal
// kernel of at most 10 taps whose first tap sits on an even offset (kernels.cu: resizeRatioStreamKernel).
// A warp walks down a column strip like the 2:1 streaming kernel; the vertical pass is driven by
// one record per destination row, the horizontal pass by the compile-time tap pattern of 8 pixels.
struct RatioPlan {
    bool eligible;
    std::string why;
    int RS, RD, NX;              // gcd-reduced horizontal ratio, horizontal taps
    int tailZeros;               // 1 when the last tap of every phase (border rows included) is zero
    int odd;                     // first[0] is odd: the kernel pairs columns (2m-1, 2m) instead of (2m, 2m+1)
    int GS;                      // source columns per group of 8 destination pixels (8 RS / RD)
    int c0;                      // first[0] on X (destination pixel 8G + p starts at GS*G + floor(p RS / RD) + c0)
    int groupsPerStrip;          // 8-pixel groups per warp strip
    int workBias;
    // vertical: per destination row {first 4-row group, groups (<= 4), 4 packed s8 words, denominator, magic}
    std::vector<int32_t> rowRec; // [DH][8]
    // horizontal: [phase (RD)][parity (2)][7] pair words, bytes (lo_a, lo_b, hi_a, hi_b)
    std::vector<uint32_t> cwX;
    int accInit;
};

void buildRatioPlan(const Plan &plan, RatioPlan &r);

}  // namespace iqo_b200

namespace iqo_b200 {

// Lanczos at any ratio (kernels.cu: resizeLanczosStreamKernel): the record-driven dp4a vertical pass of the
// rational-ratio kernel with the column-resident horizontal pass (and tables: recX, cwX, workBias) of the packed one.
struct LStreamPlan {
    bool eligible;
    std::string why;
    int stripW;                   // destination columns per warp strip (its source window fits 256 columns)
    int maxGroups;                // most 4-row groups a destination row reads (<= 8)
    // per destination row, 16 words: first 4-row group, groups, border denominator, multiply-high constant,
    // 8 packed s8 coefficient words, 4 unused
    std::vector<int32_t> rowRec;
};

void buildLStreamPlan(const Plan &plan, const PackedPlan &packed, LStreamPlan &g);

}  // namespace iqo_b200

namespace iqo_b200 {

// Lanczos at any ratio with both passes on the integer tensor path (kernels.cu: resizeLanczosMmaKernel).
// Both passes are banded integer matrix products evaluated with mma.sync.m16n8k32 (s32 accumulators):
//   vertical    W[16 dst rows x 8 columns] = A (coefficients, s8, 16 x 32 k) * B (source bytes, u8, 32 source rows x 8 columns),
//               k = consecutive source rows starting at the block's first row, 1 - 3 k-steps per 16-row block;
//   horizontal  out[16 rows x 8 dst columns] = A (bytes of W, u8, 16 x 32 k) * B (coefficient byte planes, 32 source columns x 8),
//               four products per k-step (low / high byte of W times low / high byte plane of the 14-bit coefficients).
// The planner lays the coefficient operands out in the register order of the mma fragments, one 16-byte (vertical) or
// 8-byte (horizontal) piece per lane, so that the kernel loads them with one coalesced vector load.
struct MmaPlan {
    bool eligible;
    std::string why;
    int workBias;                   // added to the intermediate: non-negative 16-bit values
    // vertical: 16-row destination blocks (global block index = dst row / 16)
    int vKMax;                      // most k-steps of a block (<= kMmaMaxKSteps)
    int nChunks;                    // 8-row chunks of the kernel's source FIFO (holds any block's rows)
    std::vector<int32_t> vBlock;    // [blocks][2]: first source row of the block's k range (may be negative), rows read from it
    std::vector<uint32_t> vFrag;    // [blocks][vKMax][32 lanes][4]: A fragments (coefficient bytes: s8 Lanczos, u8 Area / Linear)
    std::vector<int32_t> vRowMap;   // [blocks][vKMax * 32]: where the source row of every k slot sits in the kernel's FIFO (byte offset:
                                    // chunk floor(row / 8) occupies slot chunk mod nChunks, rows are wcols bytes).  Slots 0 .. rows-1
                                    // are the block's rows in order; an Area / Linear weight of 256 does not fit a byte and is split 255 + 1 over the row's
                                    // own slot and an extra slot that points at the same row; unused slots repeat the first row
    bool isSigned;                  // Lanczos: signed coefficients, 20-bit shift, border divisions; else unsigned, 23-bit shift
    std::vector<int32_t> vRow;      // [blocks * 16][2]: Lanczos border denominator (0: ordinary row), multiply-high constant
    // horizontal: 8-column destination tiles, strips of stripTiles tiles per warp
    int hKMax;
    int stripTiles;                 // tiles per strip (even, so that strips start on 16-byte boundaries of the destination)
    int wcols;                      // source columns staged per strip (multiple of 16, <= 512)
    std::vector<int32_t> stripXs;   // [strips]: source column of W element 0 (multiple of 16, may be negative)
    std::vector<int32_t> hTile;     // [tiles][2]: first source column of the tile's k range (multiple of 8), k-steps
    std::vector<uint32_t> hFrag;    // [tiles][hKMax][32 lanes][4]: B fragments {low plane b0, b1, high plane b0, b1}
    std::vector<int32_t> hCol;      // [tiles * 8][2]: accumulator init (rounding - bias * sum), divisor (0: shift by 20)
};
const int kMmaMaxKSteps = 3;
const int kMmaChunkRows = 8;      // source rows per TMA request of the kernel's FIFO

// wcols: staged source columns per strip the kernel was launched for (208 unless tuned)
void buildMmaPlan(const Plan &plan, MmaPlan &m, int wcols);

}  // namespace iqo_b200

namespace iqo_b200 {

// Optional "SIMD-semantics" float mode: what the reference's AVX-512 / AVX2 implementations compute instead
// of the Generic fixed-point path -- float tables normalised by their float sum (src/IQOLanczosResizerImpl_AVX512.cpp:
// 179-185), float FMA accumulation in tap order (:385-431, :547-590), round-to-nearest-even and saturation (:47-60) --
// with one deliberate difference: border rows AND border columns divide by the sum of the coefficients of the taps that
// are inside the image (the reference's resizeXborder adds the masked-out coefficients to its denominator, :507-519).
// Rows of `coef` follow AxisPlan's numbering (phase rows, then one row per border index); deno[r] == 0: no division.
struct FloatAxis {
    std::vector<float> coef;   // numRows x N
    std::vector<float> deno;   // numRows
};
struct FloatPlan {
    bool eligible;
    std::string why;
    FloatAxis x, y;
};
void buildFloatPlan(const Plan &plan, FloatPlan &f);

}  // namespace iqo_b200
