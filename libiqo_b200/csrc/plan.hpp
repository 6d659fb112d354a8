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
