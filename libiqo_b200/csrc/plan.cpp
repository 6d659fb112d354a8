// Host planner (see plan.hpp).  Compile WITHOUT -ffast-math / FMA contraction: the float
// rounding steps below decide the integer coefficients bit-for-bit.
#include "plan.hpp"

#include <math.h>
#include <stdio.h>

#include <algorithm>

namespace iqo_b200 {

namespace {

uint64_t gcdU64(uint64_t a, uint64_t b)
{
    while (b) {
        uint64_t r = a % b;
        a = b;
        b = r;
    }
    return a;
}

int64_t floorDiv(int64_t a, int64_t b)  // b > 0
{
    int64_t q = a / b;
    return (a % b != 0 && a < 0) ? q - 1 : q;
}

// The reference's "adjustCoefs" (src/IQOLanczosResizerImpl_Generic.cpp:341-367,
// src/IQOAreaResizerImpl_Generic.cpp:222-248): q_i = floor(c_i * one / sum + 0.5) in float,
// then the currently largest float weight absorbs +-1 until the integers sum to `one`.
// Returns false when a value cannot be held by the reference's (u)int16 table.
bool quantiseRow(std::vector<float> &w, float sum, int one, bool isUnsigned, int32_t *q)
{
    const float lo = isUnsigned ? 0.0f : -32768.0f;
    const float hi = isUnsigned ? 65535.0f : 32767.0f;
    int total = 0;
    for (size_t i = 0; i < w.size(); ++i) {
        float r = floorf(w[i] * float(one) / sum + 0.5f);
        if (!(r >= lo && r <= hi)) return false;
        q[i] = int32_t(r);
        total += q[i];
    }
    for (int step = (total < one) ? 1 : -1; total != one; total += step) {
        size_t j = size_t(std::max_element(w.begin(), w.end()) - w.begin());
        q[j] += step;
        w[j] = 0.0f;
    }
    return true;
}

// ---- Lanczos window (src/IQOLanczosResizerImpl_Generic.cpp:10-29) ----
double lanczosWeight(int degree, double x)
{
    const double kPi = 3.14159265358979;  // the reference's literal, not M_PI
    const double ax = fabs(x);
    if (fmod(ax, 1.0) < 1e-5) return ax < 1e-5 ? 1.0 : 0.0;
    if (double(degree) <= ax) return 0.0;
    const double u = kPi * x;
    const double v = kPi * (x / degree);
    return (sin(u) / u) * (sin(v) / v);
}

// Float weights of phase t (src/IQOLanczosResizerImpl_Generic.cpp:111-191), returns float sum.
float lanczosPhase(int degree, uint64_t rS, uint64_t rD, uint64_t t, uint64_t px, std::vector<float> &w)
{
    double origin;
    uint64_t num, den;  // sample spacing num/den
    if (rS > rD) {
        const int widen = std::max(1, int(px) / degree);
        const uint64_t sub = ((rD - (t * rS) % rD) * px) % rS;
        origin = -degree * widen - 0.5 * px + 0.5 * rD * px / rS + sub / double(rS);
        num = rD * px;
        den = rS;
    } else {
        origin = -degree + 1.0 - fmod((t * rS) / double(rD), 1.0);
        num = rD;
        den = rD;
    }
    float sum = 0;
    for (size_t i = 0; i < w.size(); ++i) {
        w[i] = float(lanczosWeight(degree, origin + (i * num) / double(den)));
        sum += w[i];
    }
    return sum;
}

// ---- Area weights (src/IQOAreaResizerImpl_Generic.cpp:74-97) ----
float areaPhase(uint64_t rS, uint64_t rD, uint64_t t, std::vector<float> &w)
{
    const double end = ((t + 1) * rS) / double(rD);
    double pos = (t * rS) / double(rD);
    float sum = 0;
    for (size_t i = 0; i < w.size(); ++i) {
        const double next = std::min(end, floor(pos) + 1.0);
        w[i] = float(next - pos);
        sum += w[i];
        pos = next;
    }
    return sum;
}

void finishStats(AxisPlan &a)
{
    a.coefMin = a.coefMax = 0;
    a.posSumMax = a.negSumMin = 0;
    for (int r = 0; r < a.numRows; ++r) {
        int32_t pos = 0, neg = 0;
        for (int i = 0; i < a.N; ++i) {
            int32_t c = a.coef[size_t(r) * a.N + i];
            a.coefMin = std::min(a.coefMin, c);
            a.coefMax = std::max(a.coefMax, c);
            (c > 0 ? pos : neg) += c;
        }
        a.posSumMax = std::max(a.posSumMax, pos);
        a.negSumMin = std::min(a.negSumMin, neg);
    }
}

void initAxis(AxisPlan &a, size_t S, size_t D, int bias)
{
    a.S = int64_t(S);
    a.D = int64_t(D);
    uint64_t g = gcdU64(S, D);
    a.rS = int64_t(S / g);
    a.rD = int64_t(D / g);
    a.bias = bias;
    a.identity = (S == D);
    a.mainBegin = 0;
    a.mainEnd = a.D;
    a.first.resize(D);
    a.row.resize(D);
}

// Pass-through axis: reference computes work = src*bias (Y) or (work + bias/2) >> log2(bias) (X);
// both equal a single tap of weight `bias` pushed through the ordinary fixed-point pipeline.
void identityAxis(AxisPlan &a)
{
    a.N = 1;
    a.numRows = 1;
    a.coef.assign(1, a.bias);
    a.deno.assign(1, 0);
    for (int64_t d = 0; d < a.D; ++d) {
        a.first[d] = int32_t(d);
        a.row[d] = 0;
    }
    finishStats(a);
}

int lanczosAxis(AxisPlan &a, int degree, uint64_t px, std::string &err, const char *name)
{
    if (a.identity) {
        identityAxis(a);
        return kPlanOk;
    }
    a.N = lanczosNumCoefs(degree, a.rS, a.rD, px);
    const int N = a.N, half = N / 2;
    // phase tables
    a.coef.assign(size_t(a.rD) * N, 0);
    std::vector<float> w(N);
    for (int64_t t = 0; t < a.rD; ++t) {
        float sum = lanczosPhase(degree, a.rS, a.rD, t, px, w);
        if (!quantiseRow(w, sum, a.bias, false, &a.coef[size_t(t) * N])) {
            err = std::string("Lanczos ") + name + " table: coefficient outside int16 (reference behaviour undefined)";
            return kPlanDegenerate;
        }
    }
    // main range (src/IQOLanczosResizerImpl_Generic.cpp:390-393,529-532)
    a.mainBegin = ((half - 1) * a.D + a.S - 1) / a.S;
    a.mainEnd = std::max<int64_t>(0, (a.S - half) * a.D / a.S);
    if (a.mainBegin > a.mainEnd) {
        err = std::string("Lanczos ") + name + " axis: source shorter than the kernel (reference unsupported)";
        return kPlanUnsupported;
    }
    a.deno.assign(size_t(a.rD), 0);
    a.numRows = int(a.rD);
    for (int64_t d = 0; d < a.D; ++d) {
        const int64_t f = d * a.S / a.D + 1 - half;
        const int64_t t = d % a.rD;
        a.first[d] = int32_t(f);
        if (d >= a.mainBegin && d < a.mainEnd) {
            a.row[d] = int32_t(t);
            continue;
        }
        // border index: private row with the out-of-range taps removed
        int32_t den = 0;
        const size_t base = a.coef.size();
        a.coef.resize(base + N);
        for (int i = 0; i < N; ++i) {
            const bool inside = (f + i >= 0 && f + i < a.S);
            const int32_t c = inside ? a.coef[size_t(t) * N + i] : 0;
            a.coef[base + i] = c;
            den += c;
        }
        // the Y denominator is an int16 in the reference; |den| <= sum|c| which we bound below
        if (den == 0) {
            err = std::string("Lanczos ") + name + " border: in-range coefficients sum to 0 (reference divides by zero)";
            return kPlanDegenerate;
        }
        a.deno.push_back(den);
        a.row[d] = a.numRows++;
    }
    finishStats(a);
    return kPlanOk;
}

void areaAxis(AxisPlan &a)
{
    if (a.identity) {
        identityAxis(a);
        return;
    }
    a.N = areaNumCoefs(a.rS, a.rD);
    a.numRows = int(a.rD);
    a.coef.assign(size_t(a.rD) * a.N, 0);
    a.deno.assign(size_t(a.rD), 0);
    std::vector<float> w(a.N);
    for (int64_t t = 0; t < a.rD; ++t) {
        float sum = areaPhase(a.rS, a.rD, t, w);
        quantiseRow(w, sum, a.bias, true, &a.coef[size_t(t) * a.N]);
    }
    for (int64_t d = 0; d < a.D; ++d) {
        a.first[d] = int32_t(d * a.S / a.D);  // src/IQOAreaResizerImpl_Generic.cpp:280,351
        a.row[d] = int32_t(d % a.rD);
    }
    finishStats(a);
}

void linearAxis(AxisPlan &a)
{
    if (a.identity) {
        identityAxis(a);
        return;
    }
    a.N = 2;
    a.numRows = int(a.rD) + 1;
    a.coef.assign(size_t(a.numRows) * 2, 0);
    a.deno.assign(size_t(a.numRows), 0);
    for (int64_t t = 0; t < a.rD; ++t) {
        // src/IQOLinearResizerImpl_Generic.cpp:61-68,202-206
        double whole;
        float c1 = float(modf((double(t) + 0.5) * a.rS / a.rD + 0.5, &whole));
        float c0 = 1.0f - c1;
        int32_t q0 = int32_t(uint16_t(floorf(c0 * a.bias + 0.5f)));
        a.coef[size_t(t) * 2] = q0;
        a.coef[size_t(t) * 2 + 1] = int32_t(uint16_t(a.bias - q0));
    }
    // replicated edge: one tap of full weight (resizeYborder / resizeXborder, :290-299,355-366)
    const int edgeRow = int(a.rD);
    a.coef[size_t(edgeRow) * 2] = a.bias;
    a.mainBegin = std::min<int64_t>(1, a.D);
    a.mainEnd = std::max<int64_t>(0, a.D - a.mainBegin);
    // first tap: LinearIterator(D,S).setX(S-D, 2D) advanced by d (src/math.hpp:96-112)
    const int64_t y0 = floorDiv((a.S - a.D) * a.S, 2 * a.D * a.D);
    int64_t m = (a.S - a.D) % (2 * a.D);
    if (m < 0) m += 2 * a.D;
    for (int64_t d = 0; d < a.D; ++d) {
        if (d >= a.mainEnd) {          // trailing border loop runs last in the reference
            a.first[d] = int32_t(a.S - 1);
            a.row[d] = edgeRow;
        } else if (d < a.mainBegin) {
            a.first[d] = 0;
            a.row[d] = edgeRow;
        } else {
            a.first[d] = int32_t(y0 + (m + 2 * d * a.S) / (2 * a.D));
            a.row[d] = int32_t(d % a.rD);
        }
    }
    finishStats(a);
}

}  // namespace

// src/IQOLanczosResizerImpl_Generic.cpp:32-96
int lanczosNumCoefs(int degree, uint64_t rS, uint64_t rD, uint64_t pxScale)
{
    if (rS <= rD) return 2 * degree;
    const uint64_t eff = std::max<uint64_t>(1, uint64_t(degree) / pxScale);
    return 2 * int(ceil(double(eff * rS) / double(rD)));
}

// src/IQOAreaResizerImpl_Generic.cpp:11-65
int areaNumCoefs(uint64_t rS, uint64_t rD)
{
    if (rS < rD) return 1;
    const uint64_t whole = (rS / rD) * rD;
    uint64_t n = (rS + rD - 1) / rD;
    const uint64_t l = rS / gcdU64(rS, whole) * whole;
    if (l > rS) ++n;
    return int(n);
}

int buildPlan(Plan &p, int kind, unsigned degree, size_t srcW, size_t srcH, size_t dstW, size_t dstH, size_t pxScale)
{
    p.error.clear();
    if (kind < 0 || kind > 2) {
        p.error = "unknown resizer kind";
        return kPlanBadArg;
    }
    if (!srcW || !srcH || !dstW || !dstH) {
        p.error = "image sizes must be non-zero";
        return kPlanBadArg;
    }
    if (kind == kLanczos && (degree == 0 || pxScale == 0)) {
        p.error = "Lanczos degree and pxScale must be non-zero";
        return kPlanBadArg;
    }
    const size_t kMax = size_t(1) << 30;
    if (srcW >= kMax || srcH >= kMax || dstW >= kMax || dstH >= kMax || degree > 64 || pxScale > 64) {
        p.error = "image size beyond the 2^30 limit of the index maps";
        return kPlanTooLarge;
    }
    p.kind = Kind(kind);
    p.degree = degree;
    p.pxScale = pxScale;
    int rc = kPlanOk;
    if (kind == kLanczos) {
        p.shift = 20;
        p.workSigned = true;
        initAxis(p.x, srcW, dstW, 1 << 14);
        initAxis(p.y, srcH, dstH, 1 << 6);
        rc = lanczosAxis(p.x, int(degree), pxScale, p.error, "X");
        if (rc == kPlanOk) rc = lanczosAxis(p.y, int(degree), pxScale, p.error, "Y");
    } else {
        p.shift = 23;
        p.workSigned = false;
        initAxis(p.x, srcW, dstW, 1 << 15);
        initAxis(p.y, srcH, dstH, 1 << 8);
        if (kind == kArea) {
            areaAxis(p.x);
            areaAxis(p.y);
        } else {
            linearAxis(p.x);
            linearAxis(p.y);
        }
    }
    return rc;
}

}  // namespace iqo_b200
