// Host planner (see plan.hpp).  Compile WITHOUT -ffast-math / FMA contraction: the float
// rounding steps below decide the integer coefficients bit-for-bit.
#include "plan.hpp"

#include <string.h>

#include <math.h>
#include <stdio.h>

#include <stdlib.h>

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
        // Y: the reference's row loops share their iterators and desynchronise (:390-453).  X: resizeXborder
        // re-seeds its iterator (:547-549), so an empty main range is well defined while the first border
        // loop stays inside the row (mainBegin <= D): every column takes the border formula.
        if (name[0] != 'X' || a.mainBegin > a.D) {
            err = std::string("Lanczos ") + name + " axis: source shorter than the kernel (reference unsupported)";
            return kPlanUnsupported;
        }
        a.mainBegin = a.mainEnd = 0;
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

// ---------------------------------------------------------------------------------------------
// Specialised plan for 2:1 / 2:1 single-phase Lanczos (kernels.cu: resizeHalfKernel)
// ---------------------------------------------------------------------------------------------
namespace iqo_b200 {

namespace {

int floorDivI(int a, int b)
{
    int q = a / b;
    return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q;
}

// coefficient c (|c| <= 16384) as unsigned low byte + signed high byte: c = hi*256 + lo
void splitPlanes(int c, uint32_t &lo, uint32_t &hi)
{
    int h = floorDivI(c, 256);
    lo = uint32_t(c - h * 256) & 0xffu;
    hi = uint32_t(h) & 0xffu;
}

uint32_t pairWord(int cLowHalf, int cHighHalf)
{
    uint32_t l0, h0, l1, h1;
    splitPlanes(cLowHalf, l0, h0);
    splitPlanes(cHighHalf, l1, h1);
    return l0 | (l1 << 8) | (h0 << 16) | (h1 << 24);
}

}  // namespace

void buildHalfPlan(const Plan &p, HalfPlan &h)
{
    h.eligible = false;
    h.sEligible = false;
    h.why.clear();
    const AxisPlan &X = p.x, &Y = p.y;
    if (p.kind != kLanczos) { h.why = "not Lanczos"; return; }
    if (!X.identity && X.mainBegin >= X.mainEnd) { h.why = "no main columns (source narrower than the kernel)"; return; }
    if (X.rD != 1 || X.rS != 2 || Y.rD != 1 || Y.rS != 2) { h.why = "not 2:1 on both axes"; return; }
    if (X.S % 4 != 0) { h.why = "source width not a multiple of 4"; return; }
    const int NY = Y.N, NX = X.N;
    if (NX > 12 || (NX & 1) || ((NX / 2) & 1)) { h.why = "horizontal kernel longer than 12 taps or of odd half-length"; return; }
    if (Y.coefMin < -128 || Y.coefMax > 127) { h.why = "vertical coefficients do not fit int8"; return; }

    // ---- vertical: trim zero taps of the main phase, place taps into 4-row groups ----
    const int32_t *cy = &Y.coef[0];
    int lead = 0, trail = 0;
    while (lead < NY - 1 && cy[lead] == 0) ++lead;
    while (trail < NY - 1 - lead && cy[NY - 1 - trail] == 0) ++trail;
    const int cy0 = 1 - NY / 2 + lead;        // first non-zero tap of row y sits on source row 2y + cy0
    const int NYt = NY - lead - trail;
    int qlo = 1 << 30, qhi = -(1 << 30);
    for (int par = 0; par < 2; ++par) {
        qlo = std::min(qlo, floorDivI(2 * par + cy0, 4));
        qhi = std::max(qhi, floorDivI(2 * par + cy0 + NYt - 1, 4));
    }
    h.qmin = qlo;
    h.NG = qhi - qlo + 1;
    if (h.NG > 3) { h.why = "vertical kernel spans more than three 4-row groups"; return; }
    // packed words of every coefficient row; a row r >= 1 belongs to one destination row y
    std::vector<int> rowOwner(size_t(Y.numRows), -1);
    for (int64_t y = 0; y < Y.D; ++y)
        if (Y.row[y] != 0) rowOwner[size_t(Y.row[y])] = int(y);
    h.borderY.assign(size_t(Y.numRows) * 3, 0);
    for (int par = 0; par < 2; ++par)
        for (int g = 0; g < 3; ++g) h.cwY[par][g] = 0;
    for (int r = 0; r < Y.numRows; ++r) {
        for (int par = 0; par < 2; ++par) {
            if (r > 0 && (rowOwner[size_t(r)] & 1) != par) continue;
            uint32_t w[3] = {0, 0, 0};
            for (int i = 0; i < NY; ++i) {
                const int c = Y.coef[size_t(r) * NY + i];
                if (c == 0) continue;
                const int pos = 2 * par + (1 - NY / 2 + i) - 4 * h.qmin;  // byte position inside the NG groups
                if (pos < 0 || pos >= 4 * h.NG) { h.why = "internal: tap outside the group window"; return; }
                w[pos >> 2] |= (uint32_t(c) & 0xffu) << (8 * (pos & 3));
            }
            for (int g = 0; g < 3; ++g) {
                if (r == 0) h.cwY[par][g] = w[g];
                else h.borderY[size_t(r) * 3 + g] = w[g];
            }
        }
    }
    // range of the intermediate over all rows (border rows are rescaled by 64/deno)
    long long wmin = 0, wmax = 0;
    for (int r = 0; r < Y.numRows; ++r) {
        long long pos = 0, neg = 0;
        for (int i = 0; i < NY; ++i) {
            const int c = Y.coef[size_t(r) * NY + i];
            (c > 0 ? pos : neg) += c;
        }
        long long lo = 255 * neg, hi = 255 * pos;
        if (lo < -32768 || hi > 32767) { h.why = "vertical sum may wrap int16"; return; }
        const int den = Y.deno[size_t(r)];
        if (den != 0) {
            if (den < 0) { h.why = "negative border denominator"; return; }
            lo = lo * 64 / den - 1;
            hi = hi * 64 / den + 1;
            // the reference casts the rescaled border row to int16 (resizeYborder): content that would wrap it lies outside
            // the biased 16-bit window of the fast kernels, so such tables go to the generic kernel, which wraps like the reference
            if (lo < -32768 || hi > 32767) { h.why = "border row may wrap int16 after its division"; return; }
        }
        wmin = std::min(wmin, lo);
        wmax = std::max(wmax, hi);
    }
    h.workBias = int(-wmin);
    if (wmax + h.workBias > 32767) { h.why = "intermediate range too wide for pair sums"; return; }

    // ---- horizontal ----
    const int32_t *cx = &X.coef[0];
    const int cx0 = 1 - NX / 2;               // odd
    h.wa = (cx0 - 1) / 2;                     // exact: cx0 - 1 is even and negative
    h.NWX = NX / 2 + 1;
    for (int i = 0; i < 7; ++i) h.cwX[i] = 0;
    for (int i = 0; i < h.NWX; ++i) {
        const int ta = 2 * i - 1, tb = 2 * i;  // taps in the low / high half of pair word i
        h.cwX[i] = pairWord(ta >= 0 ? cx[ta] : 0, tb < NX ? cx[tb] : 0);
    }
    h.symmetric = true;
    for (int i = 0; i < NX; ++i)
        if (cx[i] != cx[NX - 1 - i]) h.symmetric = false;
    for (int i = 0; i < 4; ++i) h.cwXs[i] = 0;
    if (h.symmetric) {
        const int m = h.NWX / 2;               // centre pair word
        for (int j = 1; j < m; ++j) h.cwXs[j - 1] = pairWord(cx[2 * j - 1], cx[2 * j]);
        h.cwXs[m - 1] = pairWord(cx[2 * m - 1], cx[2 * m]);
        h.cwXs[m] = pairWord(cx[NX - 1], cx[0]);  // low half: last tap, high half: first tap
    }
    long long sumX = 0;
    for (int i = 0; i < NX; ++i) sumX += cx[i];
    h.accInit = int((1ll << (p.shift - 1)) - (long long)h.workBias * sumX);

    // ---- borders ----
    h.magicY.assign(size_t(Y.numRows), 0);
    for (int r = 0; r < Y.numRows; ++r)
        if (Y.deno[size_t(r)] > 1) h.magicY[size_t(r)] = uint32_t((1ull << 32) / uint64_t(Y.deno[size_t(r)]) + 1);
    h.borderX.clear();
    for (int64_t d = 0; d < X.D; ++d) {
        if (d >= X.mainBegin && d < X.mainEnd) continue;
        const int32_t *c = &X.coef[size_t(X.row[d]) * NX];
        long long sum = 0;
        for (int i = 0; i < h.NWX; ++i) {
            const int ta = 2 * i - 1, tb = 2 * i;
            h.borderX.push_back(int32_t(pairWord(ta >= 0 ? c[ta] : 0, tb < NX ? c[tb] : 0)));
        }
        for (int i = h.NWX; i < 7; ++i) h.borderX.push_back(0);
        for (int i = 0; i < NX; ++i) sum += c[i];
        h.borderX.push_back(int32_t(X.deno[size_t(X.row[d])] * 64));
        h.borderX.push_back(int32_t((1ll << (p.shift - 1)) - (long long)h.workBias * sum));
    }
    h.eligible = true;

    // ---- streaming variant ----
    h.sEligible = false;
    {
        // group offset: fewest non-zero main-phase words among the patterns the kernel is built for
        // (no zero word, or only the last word of parity 0 zero)
        int best = 1 << 30;
        for (int delta = 0; delta < 4; ++delta) {
            int lo = 1 << 30, hi = -(1 << 30);
            for (int par = 0; par < 2; ++par) {
                lo = std::min(lo, floorDivI(2 * par + cy0 - delta, 4));
                hi = std::max(hi, floorDivI(2 * par + cy0 + NYt - 1 - delta, 4));
            }
            const int ng = hi - lo + 1;
            if (ng < 2 || ng > 3) continue;  // group counts the streaming kernel is built for
            uint32_t w[2][3] = {{0, 0, 0}, {0, 0, 0}};
            for (int par = 0; par < 2; ++par)
                for (int i = 0; i < NY; ++i) {
                    const int c = cy[i];
                    if (c == 0) continue;
                    const int pos = 2 * par + (1 - NY / 2 + i) - delta - 4 * lo;
                    w[par][pos >> 2] |= (uint32_t(c) & 0xffu) << (8 * (pos & 3));
                }
            int z = 0, words = 0;
            for (int par = 0; par < 2; ++par)
                for (int g = 0; g < ng; ++g) {
                    if (w[par][g] == 0) z |= 1 << (par * 3 + g);
                    else ++words;
                }
            if (!(z == 0 || (z == 4 && ng == 3))) {  // pattern without a kernel: compute the zero words too
                z = 0;
                words = 2 * ng;
            }
            if (words < best) {
                best = words;
                h.sDelta = delta;
                h.sQmin = lo;
                h.sNG = ng;
                h.sZ = z;
                memcpy(h.sCwY, w, sizeof w);
            }
        }
        if (best < (1 << 30) && h.sNG >= 2) {
            h.sBorderY.assign(size_t(Y.numRows) * 3, 0);
            bool ok = true;
            for (int r = 1; r < Y.numRows && ok; ++r) {
                if (rowOwner[size_t(r)] < 0) continue;
                const int par = rowOwner[size_t(r)] & 1;
                for (int i = 0; i < NY; ++i) {
                    const int c = Y.coef[size_t(r) * NY + i];
                    if (c == 0) continue;
                    const int pos = 2 * par + (1 - NY / 2 + i) - h.sDelta - 4 * h.sQmin;
                    if (pos < 0 || pos >= 4 * h.sNG) { ok = false; break; }
                    h.sBorderY[size_t(r) * 3 + size_t(pos >> 2)] |= (uint32_t(c) & 0xffu) << (8 * (pos & 3));
                }
            }
            h.NXH = NX / 2;
            for (int i = 0; i < 6; ++i) h.cwXo[i] = 0;
            for (int i = 0; i < h.NXH; ++i) h.cwXo[i] = pairWord(cx[2 * i], cx[2 * i + 1]);
            h.skipHi0 = h.symmetric && (h.cwXo[0] >> 16) == 0;
            h.borderXo.clear();
            for (int64_t d = 0; d < X.D; ++d) {
                if (d >= X.mainBegin && d < X.mainEnd) continue;
                const int32_t *c = &X.coef[size_t(X.row[d]) * NX];
                long long sum = 0;
                for (int i = 0; i < 6; ++i) h.borderXo.push_back(i < h.NXH ? int32_t(pairWord(c[2 * i], c[2 * i + 1])) : 0);
                for (int i = 0; i < NX; ++i) sum += c[i];
                h.borderXo.push_back(int32_t(X.deno[size_t(X.row[d])] * 64));
                h.borderXo.push_back(int32_t((1ll << (p.shift - 1)) - (long long)h.workBias * sum));
            }
            h.sEligible = ok && X.D >= 32;  // narrower images: left and right border columns would share W chunks
        }
    }
}

}  // namespace iqo_b200

// ---------------------------------------------------------------------------------------------
// Plan of the general packed kernel
// ---------------------------------------------------------------------------------------------
namespace iqo_b200 {

void buildPackedPlan(const Plan &p, PackedPlan &q, int padNP)
{
    q.eligible = false;
    q.why.clear();
    const AxisPlan &X = p.x, &Y = p.y;
    const bool isSigned = p.workSigned;
    if (p.kind == kLanczos && !X.identity && X.mainBegin >= X.mainEnd) { q.why = "no main columns (source narrower than the kernel)"; return; }

    // ---- vertical: every out-of-image tap must be weightless; trim them off ----
    q.firstY.assign(size_t(Y.D), 0);
    q.ntapY.assign(size_t(Y.D), 0);
    q.coefOffY.assign(size_t(Y.D), 0);
    for (int64_t y = 0; y < Y.D; ++y) {
        const int32_t *c = &Y.coef[size_t(Y.row[y]) * Y.N];
        int lo = 0, hi = Y.N;  // taps [lo, hi) are inside the image
        const int64_t f = Y.first[y];
        if (f < 0) lo = int(std::min<int64_t>(-f, Y.N));
        if (f + Y.N > Y.S) hi = int(std::max<int64_t>(Y.S - f, lo));
        for (int i = 0; i < Y.N; ++i)
            if ((i < lo || i >= hi) && c[i] != 0) { q.why = "vertical tap outside the image with non-zero weight"; return; }
        while (lo < hi && c[lo] == 0) ++lo;            // leading / trailing zero taps cost nothing
        while (hi > lo && c[hi - 1] == 0) --hi;
        if (hi == lo) { lo = 0; hi = 0; }
        q.firstY[size_t(y)] = int32_t(f + lo);
        q.ntapY[size_t(y)] = hi - lo;
        q.coefOffY[size_t(y)] = int32_t(size_t(Y.row[y]) * Y.N + lo);
        if (hi == lo) q.firstY[size_t(y)] = 0;
    }
    // 16-bit lane range of the intermediate (border rows are rescaled by 64/deno)
    long long wmin = 0, wmax = 0, laneMin = 0, laneMax = 0;
    q.magicY.assign(size_t(Y.numRows), 0);
    for (int r = 0; r < Y.numRows; ++r) {
        long long pos = 0, neg = 0;
        for (int i = 0; i < Y.N; ++i) {
            const int c = Y.coef[size_t(r) * Y.N + i];
            (c > 0 ? pos : neg) += c;
        }
        long long lo = 255 * neg, hi = 255 * pos;
        laneMin = std::min(laneMin, lo);   // partial sums of the packed multiply-adds
        laneMax = std::max(laneMax, hi);
        if (isSigned ? (lo < -32768 || hi > 32767) : (hi > 65535)) { q.why = "vertical sum may wrap 16 bits"; return; }
        const int den = Y.deno[size_t(r)];
        if (den != 0) {
            if (den < 0 || den > 255) { q.why = "border denominator out of range"; return; }
            lo = lo * 64 / den - 1;
            hi = hi * 64 / den + 1;
            // the reference casts the rescaled border row to int16 (resizeYborder): content that would wrap it lies outside
            // the biased 16-bit window of the fast kernels, so such tables go to the generic kernel, which wraps like the reference
            if (lo < -32768 || hi > 32767) { q.why = "border row may wrap int16 after its division"; return; }
            if (den > 1) q.magicY[size_t(r)] = uint32_t((1ull << 32) / uint64_t(den) + 1);
        }
        wmin = std::min(wmin, lo);
        wmax = std::max(wmax, hi);
    }
    q.workBias = int(-std::min(wmin, laneMin));
    if (std::max(wmax, laneMax) + q.workBias > 65535) { q.why = "intermediate range wider than a 16-bit lane"; return; }

    // ---- horizontal ----
    q.firstX.assign(size_t(X.D), 0);
    if (padNP < X.N / 2 + 1) { q.why = "horizontal kernel longer than the packed kernel supports"; return; }
    q.NP = padNP;
    q.ntMax = 1;
    for (int64_t y = 0; y < Y.D; ++y) q.ntMax = std::max(q.ntMax, q.ntapY[size_t(y)]);
    if (q.ntMax > 64) { q.why = "vertical kernel longer than 64 taps"; return; }
    std::vector<int> rowShift(size_t(X.numRows), 0);  // leading taps dropped because they lie left of column 0
    for (int64_t d = 0; d < X.D; ++d) {
        const int r = X.row[d];
        const int32_t *c = &X.coef[size_t(r) * X.N];
        const int64_t f = X.first[d];
        const int shift = f < 0 ? int(std::min<int64_t>(-f, X.N)) : 0;
        for (int i = 0; i < X.N; ++i) {
            const bool outside = (f + i < 0) || (f + i >= X.S);
            if (outside && c[i] != 0) { q.why = "horizontal tap outside the image with non-zero weight"; return; }
        }
        if (shift != 0 && r < int(X.rD) && X.rD != X.D) {
            // a phase row shared by several columns cannot be re-based; only planner-made
            // per-column rows (Lanczos borders) or one-row-per-column tables get here
            bool shared = false;
            for (int64_t e = 0; e < X.D && !shared; ++e) shared = (e != d && X.row[e] == r);
            if (shared) { q.why = "shared phase row starts left of the image"; return; }
        }
        rowShift[size_t(r)] = shift;
        q.firstX[size_t(d)] = int32_t(f + shift);
    }
    q.cwX.assign(size_t(X.numRows) * 2 * q.NP, 0);
    std::vector<int32_t> accInit(size_t(X.numRows), 0), divisor(size_t(X.numRows), 0);
    for (int r = 0; r < X.numRows; ++r) {
        const int32_t *c = &X.coef[size_t(r) * X.N];
        const int shift = rowShift[size_t(r)];
        long long sum = 0;
        for (int i = 0; i < X.N; ++i) sum += c[i];
        for (int par = 0; par < 2; ++par) {
            // window element i (coefficient c[shift + i]) sits in half (par + i) & 1 of word (par + i) >> 1
            for (int w = 0; w < q.NP; ++w) {
                const int ia = 2 * w - par, ib = 2 * w + 1 - par;  // taps in the low / high half
                const int ca = (ia >= 0 && shift + ia < X.N) ? c[shift + ia] : 0;
                const int cb = (ib >= 0 && shift + ib < X.N) ? c[shift + ib] : 0;
                uint32_t la, ha, lb, hb;
                if (isSigned) {
                    splitPlanes(ca, la, ha);
                    splitPlanes(cb, lb, hb);
                } else {
                    la = uint32_t(ca) & 0xff; ha = (uint32_t(ca) >> 8) & 0xff;
                    lb = uint32_t(cb) & 0xff; hb = (uint32_t(cb) >> 8) & 0xff;
                }
                q.cwX[(size_t(r) * 2 + par) * q.NP + w] = la | (lb << 8) | (ha << 16) | (hb << 24);
            }
        }
        accInit[size_t(r)] = int32_t((1ll << (p.shift - 1)) - (long long)q.workBias * sum);
        divisor[size_t(r)] = X.deno[size_t(r)] * 64;
    }
    q.recX.assign(size_t(X.D) * 4, 0);
    for (int64_t d = 0; d < X.D; ++d) {
        const int r = X.row[d];
        const int f = q.firstX[size_t(d)];
        q.recX[size_t(d) * 4 + 0] = f;
        q.recX[size_t(d) * 4 + 1] = int32_t((size_t(r) * 2 + (f & 1)) * q.NP);
        q.recX[size_t(d) * 4 + 2] = accInit[size_t(r)];
        q.recX[size_t(d) * 4 + 3] = divisor[size_t(r)];
    }
    q.eligible = true;
}

}  // namespace iqo_b200

// ---------------------------------------------------------------------------------------------
// Plan of the streaming small-kernel 2:1 path
// ---------------------------------------------------------------------------------------------
namespace iqo_b200 {

void buildSmallPlan(const Plan &p, SmallPlan &s)
{
    s.eligible = false;
    s.why.clear();
    const AxisPlan &X = p.x, &Y = p.y;
    if (p.kind != kLanczos) { s.why = "not Lanczos"; return; }
    if (!X.identity && X.mainBegin >= X.mainEnd) { s.why = "no main columns (source narrower than the kernel)"; return; }
    if (X.rD != 1 || X.rS != 2 || Y.rD != 1 || Y.rS != 2) { s.why = "not 2:1 on both axes"; return; }
    if (X.S % 16 != 0) { s.why = "source width not a multiple of 16"; return; }
    // trimmed main-phase taps
    int ly = 0, ty = Y.N;
    while (ly < Y.N - 1 && Y.coef[ly] == 0) ++ly;
    while (ty > ly + 1 && Y.coef[ty - 1] == 0) --ty;
    int lx = 0, tx = X.N;
    while (lx < X.N - 1 && X.coef[lx] == 0) ++lx;
    while (tx > lx + 1 && X.coef[tx - 1] == 0) --tx;
    s.TY = ty - ly;
    s.cy0 = 1 - Y.N / 2 + ly;
    s.TX = tx - lx;
    s.cx0 = 1 - X.N / 2 + lx;
    if (s.TY > 4 || s.TX > 4) { s.why = "more than four non-zero taps"; return; }
    if (s.cx0 < -2 || s.cx0 + s.TX - 1 > 3) { s.why = "horizontal window outside the thread's column range"; return; }
    for (int t = 0; t < 4; ++t) s.cY[t] = (t < s.TY) ? Y.coef[ly + t] : 0;
    // every coefficient row at the main tap positions; a border row must not need other positions
    s.rowsY.assign(size_t(Y.numRows) * 4, 0);
    s.magicY.assign(size_t(Y.numRows), 0);
    long long wmin = 0, wmax = 0;
    for (int r = 0; r < Y.numRows; ++r) {
        long long pos = 0, neg = 0;
        for (int i = 0; i < Y.N; ++i) {
            const int c = Y.coef[size_t(r) * Y.N + i];
            if (c != 0 && (i < ly || i >= ty)) { s.why = "border row uses a tap outside the trimmed window"; return; }
            (c > 0 ? pos : neg) += c;
        }
        for (int t = 0; t < s.TY; ++t) s.rowsY[size_t(r) * 4 + t] = Y.coef[size_t(r) * Y.N + ly + t];
        long long lo = 255 * neg, hi = 255 * pos;
        if (lo < -32768 || hi > 32767) { s.why = "vertical sum may wrap int16"; return; }
        const int den = Y.deno[size_t(r)];
        if (den != 0) {
            if (den < 0 || den > 255) { s.why = "border denominator out of range"; return; }
            lo = lo * 64 / den - 1;
            hi = hi * 64 / den + 1;
            // the reference casts the rescaled border row to int16 (resizeYborder): content that would wrap it lies outside
            // the biased 16-bit window of the fast kernels, so such tables go to the generic kernel, which wraps like the reference
            if (lo < -32768 || hi > 32767) { s.why = "border row may wrap int16 after its division"; return; }
            if (den > 1) s.magicY[size_t(r)] = uint32_t((1ull << 32) / uint64_t(den) + 1);
        }
        wmin = std::min(wmin, lo);
        wmax = std::max(wmax, hi);
    }
    s.workBias = int(-wmin);
    if (wmax + s.workBias > 65535) { s.why = "intermediate range wider than a 16-bit lane"; return; }
    for (int64_t d = 0; d < X.D; ++d)
        if (X.deno[size_t(X.row[d])] < 0) { s.why = "negative border denominator"; return; }
    // horizontal pair words: column 2d + cx0 + i sits in half (par + i) & 1 of word d + wbase + ((par + i) >> 1)
    const int par = s.cx0 & 1;
    s.wbase = floorDivI(s.cx0, 2);
    s.NW = (par + s.TX + 1) / 2;
    long long sum = 0;
    for (int i = 0; i < X.N; ++i) sum += X.coef[i];
    for (int w = 0; w < 3; ++w) {
        const int ia = 2 * w - par, ib = 2 * w + 1 - par;
        const int ca = (ia >= 0 && ia < s.TX) ? X.coef[lx + ia] : 0;
        const int cb = (ib >= 0 && ib < s.TX) ? X.coef[lx + ib] : 0;
        s.cwX[w] = (w < s.NW) ? pairWord(ca, cb) : 0u;
    }
    s.accInit = int((1ll << (p.shift - 1)) - (long long)s.workBias * sum);
    s.eligible = true;
}

}  // namespace iqo_b200

// ---------------------------------------------------------------------------------------------
// Plan of the rational-ratio streaming kernel
// ---------------------------------------------------------------------------------------------
namespace iqo_b200 {

void buildRatioPlan(const Plan &p, RatioPlan &r)
{
    r.eligible = false;
    r.why.clear();
    const AxisPlan &X = p.x, &Y = p.y;
    if (p.kind != kLanczos) { r.why = "not Lanczos"; return; }
    if (!X.identity && X.mainBegin >= X.mainEnd) { r.why = "no main columns (source narrower than the kernel)"; return; }
    if (X.identity || Y.identity) { r.why = "pass-through axis"; return; }
    if (X.rD > 8 || 8 % X.rD != 0) { r.why = "horizontal period does not divide 8"; return; }
    if (X.D % 8 != 0 || X.S % 8 != 0) { r.why = "widths not multiples of 8"; return; }
    const int NX = X.N, NY = Y.N;
    if (NX > 12 || (NX & 1)) { r.why = "horizontal kernel longer than 12 taps"; return; }
    if (Y.coefMin < -128 || Y.coefMax > 127) { r.why = "vertical coefficients do not fit int8"; return; }
    r.RS = int(X.rS);
    r.RD = int(X.rD);
    r.NX = NX;
    r.GS = int(8 * X.rS / X.rD);
    if (r.GS & 1) { r.why = "odd source span per 8 pixels"; return; }
    r.c0 = X.first[0];
    r.odd = r.c0 & 1;
    // the compile-time tap pattern must hold for every destination column
    for (int64_t d = 0; d < X.D; ++d) {
        const int64_t G = d / 8, q = d % 8;
        if (X.first[size_t(d)] != r.GS * G + (q * X.rS) / X.rD + r.c0) { r.why = "first-tap pattern is not periodic in 8"; return; }
        if (d >= X.mainBegin && d < X.mainEnd && X.row[size_t(d)] != d % X.rD) { r.why = "unexpected coefficient row"; return; }
    }
    // the last tap of the strip's last pixel must lie inside the 256 columns (255 with the one-column shift) of a warp
    const int i0 = ((r.c0 % 8) + 8) % 8;
    const int off7 = int((7 * X.rS) / X.rD);
    r.groupsPerStrip = std::min(32, (255 - r.odd - i0 - off7 - (NX - 1)) / r.GS + 1);
    if (r.groupsPerStrip < 1) { r.why = "source window wider than a warp strip"; return; }
    while (r.groupsPerStrip > 1 && (r.GS * r.groupsPerStrip) % 8 != 0) --r.groupsPerStrip;  // strips start on whole 8-column words
    if ((r.GS * r.groupsPerStrip) % 8 != 0) { r.why = "no strip width keeps the source window 8-byte aligned"; return; }

    // ---- vertical records ----
    long long wmin = 0, wmax = 0;
    for (int rr = 0; rr < Y.numRows; ++rr) {
        long long pos = 0, neg = 0;
        for (int i = 0; i < NY; ++i) {
            const int c = Y.coef[size_t(rr) * NY + i];
            (c > 0 ? pos : neg) += c;
        }
        long long lo = 255 * neg, hi = 255 * pos;
        if (lo < -32768 || hi > 32767) { r.why = "vertical sum may wrap int16"; return; }
        const int den = Y.deno[size_t(rr)];
        if (den != 0) {
            if (den < 0 || den > 255) { r.why = "border denominator out of range"; return; }
            lo = lo * 64 / den - 1;
            hi = hi * 64 / den + 1;
            // the reference casts the rescaled border row to int16 (resizeYborder): content that would wrap it lies outside
            // the biased 16-bit window of the fast kernels, so such tables go to the generic kernel, which wraps like the reference
            if (lo < -32768 || hi > 32767) { r.why = "border row may wrap int16 after its division"; return; }
        }
        wmin = std::min(wmin, lo);
        wmax = std::max(wmax, hi);
    }
    r.workBias = int(-wmin);
    if (wmax + r.workBias > 65535) { r.why = "intermediate range too wide"; return; }
    // first / last source row with a non-zero weight, per destination row
    std::vector<int> rlo(size_t(Y.D), 0), rhi(size_t(Y.D), -1);
    for (int64_t y = 0; y < Y.D; ++y) {
        const int32_t *c = &Y.coef[size_t(Y.row[size_t(y)]) * NY];
        const int f = Y.first[size_t(y)];
        bool any = false;
        for (int i = 0; i < NY; ++i)
            if (c[i] != 0) {
                const int row = f + i;
                if (row < 0 || row >= Y.S) { r.why = "vertical tap outside the image with non-zero weight"; return; }
                if (!any) rlo[size_t(y)] = row;
                rhi[size_t(y)] = row;
                any = true;
            }
        if (!any) { r.why = "all-zero vertical row"; return; }
    }
    // the kernel parks source groups in order: the first group of a row may not lie before that of an
    // earlier row (phases whose leading coefficients are zero would otherwise step back)
    r.rowRec.assign(size_t(Y.D) * 8, 0);
    int gmin = 1 << 30;
    for (int64_t y = Y.D - 1; y >= 0; --y) {
        gmin = std::min(gmin, rlo[size_t(y)] / 4);
        const int g0 = gmin, g1 = rhi[size_t(y)] / 4;
        if (g1 - g0 + 1 > 4) { r.why = "vertical kernel spans more than four 4-row groups"; return; }
        const int32_t *c = &Y.coef[size_t(Y.row[size_t(y)]) * NY];
        const int f = Y.first[size_t(y)];
        int32_t *rec = &r.rowRec[size_t(y) * 8];
        rec[0] = g0;
        rec[1] = g1 - g0 + 1;
        for (int i = 0; i < NY; ++i)
            if (c[i] != 0) {
                const int pos = f + i - 4 * g0;
                rec[2 + (pos >> 2)] |= int32_t((uint32_t(c[i]) & 0xffu) << (8 * (pos & 3)));
            }
        const int den = Y.deno[size_t(Y.row[size_t(y)])];
        rec[6] = den;
        rec[7] = den > 1 ? int32_t(uint32_t((1ull << 32) / uint64_t(den) + 1)) : 0;
    }

    // ---- horizontal pair words: natural pairs (2m, 2m+1); parity = first tap on an odd W element ----
    r.cwX.assign(size_t(r.RD) * 2 * 7, 0);
    for (int ph = 0; ph < r.RD; ++ph) {
        const int32_t *c = &X.coef[size_t(ph) * NX];
        for (int par = 0; par < 2; ++par)
            for (int j = 0; j < 7; ++j) {
                const int ta = 2 * j - par, tb = 2 * j + 1 - par;  // taps in the low / high half of word j
                r.cwX[(size_t(ph) * 2 + par) * 7 + j] =
                    pairWord(ta >= 0 && ta < NX ? c[ta] : 0, tb >= 0 && tb < NX ? c[tb] : 0);
            }
    }
    long long sumX = 0;
    for (int i = 0; i < NX; ++i) sumX += X.coef[i];
    for (int ph = 1; ph < r.RD; ++ph) {
        long long s2 = 0;
        for (int i = 0; i < NX; ++i) s2 += X.coef[size_t(ph) * NX + i];
        if (s2 != sumX) { r.why = "phase tables with different sums"; return; }
    }
    r.accInit = int((1ll << (p.shift - 1)) - (long long)r.workBias * sumX);
    r.tailZeros = 1;
    for (int ph = 0; ph < r.RD; ++ph)
        if (X.coef[size_t(ph) * NX + NX - 1] != 0) r.tailZeros = 0;
    r.eligible = true;
}

}  // namespace iqo_b200

// ---------------------------------------------------------------------------------------------
// Plan of the general Lanczos streaming kernel
// ---------------------------------------------------------------------------------------------
namespace iqo_b200 {

void buildLStreamPlan(const Plan &p, const PackedPlan &q, LStreamPlan &g)
{
    g.eligible = false;
    g.why.clear();
    const AxisPlan &X = p.x, &Y = p.y;
    if (p.kind != kLanczos) { g.why = "not Lanczos"; return; }
    if (!q.eligible) { g.why = "packed tables unavailable: " + q.why; return; }
    if (X.identity || Y.identity) { g.why = "pass-through axis"; return; }
    if (Y.coefMin < -128 || Y.coefMax > 127) { g.why = "vertical coefficients do not fit int8"; return; }
    const int NX = X.N, NY = Y.N;
    // strip width: the source window [firstX[t0] & ~7, firstX[t1] + NX) of every strip must fit 256 columns
    int w = 256;
    for (;;) {
        bool ok = true;
        for (int64_t t0 = 0; t0 < X.D && ok; t0 += w) {
            const int64_t t1 = std::min<int64_t>(X.D, t0 + w) - 1;
            const int lo = q.firstX[size_t(t0)] & ~7, hi = q.firstX[size_t(t1)] + NX - 1;
            if (hi - lo > 255) ok = false;
        }
        if (ok) break;
        w -= 8;
        if (w < 8) { g.why = "horizontal kernel wider than a warp strip"; return; }
    }
    g.stripW = w;
    // vertical records (see buildRatioPlan)
    std::vector<int> rlo(size_t(Y.D), 0), rhi(size_t(Y.D), -1);
    for (int64_t y = 0; y < Y.D; ++y) {
        const int32_t *c = &Y.coef[size_t(Y.row[size_t(y)]) * NY];
        const int f = Y.first[size_t(y)];
        bool any = false;
        for (int i = 0; i < NY; ++i)
            if (c[i] != 0) {
                const int row = f + i;
                if (row < 0 || row >= Y.S) { g.why = "vertical tap outside the image with non-zero weight"; return; }
                if (!any) rlo[size_t(y)] = row;
                rhi[size_t(y)] = row;
                any = true;
            }
        if (!any) { g.why = "all-zero vertical row"; return; }
    }
    g.rowRec.assign(size_t(Y.D) * 16, 0);
    g.maxGroups = 1;
    int gmin = 1 << 30;
    for (int64_t y = Y.D - 1; y >= 0; --y) {
        gmin = std::min(gmin, rlo[size_t(y)] / 4);
        const int g0 = gmin, g1 = rhi[size_t(y)] / 4;
        if (g1 - g0 + 1 > 8) { g.why = "vertical kernel spans more than eight 4-row groups"; return; }
        g.maxGroups = std::max(g.maxGroups, g1 - g0 + 1);
        const int32_t *c = &Y.coef[size_t(Y.row[size_t(y)]) * NY];
        const int f = Y.first[size_t(y)];
        int32_t *rec = &g.rowRec[size_t(y) * 16];
        rec[0] = g0;
        rec[1] = g1 - g0 + 1;
        const int den = Y.deno[size_t(Y.row[size_t(y)])];
        if (den < 0 || den > 255) { g.why = "border denominator out of range"; return; }
        rec[2] = den;
        rec[3] = den > 1 ? int32_t(uint32_t((1ull << 32) / uint64_t(den) + 1)) : 0;
        for (int i = 0; i < NY; ++i)
            if (c[i] != 0) {
                const int pos = f + i - 4 * g0;
                rec[4 + (pos >> 2)] |= int32_t((uint32_t(c[i]) & 0xffu) << (8 * (pos & 3)));
            }
    }
    g.eligible = true;
}

}  // namespace iqo_b200

// ---------------------------------------------------------------------------------------------
// Plan of the tensor-path Lanczos kernel (both passes as banded integer matrix products)
// ---------------------------------------------------------------------------------------------
namespace iqo_b200 {

namespace {

int64_t floorTo(int64_t v, int64_t m)
{
    int64_t q = v / m;
    if (v % m != 0 && v < 0) --q;
    return q * m;
}

}  // namespace

void buildMmaPlan(const Plan &p, MmaPlan &m, int wcols)
{
    m.eligible = false;
    m.why.clear();
    m.workBias = 0;
    m.vKMax = m.hKMax = 1;
    m.nChunks = 2;
    m.stripTiles = 0;
    m.wcols = wcols;
    const AxisPlan &X = p.x, &Y = p.y;
    m.isSigned = p.workSigned;
    if (p.kind == kLanczos) {
        if (!X.identity && X.mainBegin >= X.mainEnd) { m.why = "no main columns (source narrower than the kernel)"; return; }
        if (Y.coefMin < -128 || Y.coefMax > 127) { m.why = "vertical coefficients do not fit int8"; return; }
        if (X.coefMin < -32768 || X.coefMax > 32767) { m.why = "horizontal coefficients do not fit two byte planes"; return; }
    } else {
        if (Y.coefMin < 0 || Y.coefMax > 256) { m.why = "vertical weights outside 0..256"; return; }
        if (X.coefMin < 0 || X.coefMax > 32768) { m.why = "horizontal weights outside 0..32768"; return; }
    }
    if (X.S % 2 != 0) { m.why = "odd source width (the tensor map views the rows as 16-bit pairs)"; return; }
    if (wcols % 16 != 0 || wcols < 64 || wcols > 512) { m.why = "bad strip width"; return; }
    const int NX = X.N, NY = Y.N;

    // ---- intermediate range (as buildPackedPlan): no 16-bit wrap, border rows rescaled by 64 / deno ----
    long long wmin = 0, wmax = 0;
    for (int r = 0; r < Y.numRows; ++r) {
        long long pos = 0, neg = 0;
        for (int i = 0; i < NY; ++i) {
            const int c = Y.coef[size_t(r) * NY + i];
            (c > 0 ? pos : neg) += c;
        }
        long long lo = 255 * neg, hi = 255 * pos;
        if (m.isSigned ? (lo < -32768 || hi > 32767) : (hi > 65535)) { m.why = "vertical sum may wrap 16 bits"; return; }
        const int den = Y.deno[size_t(r)];
        if (den != 0) {
            if (den < 0 || den > 255) { m.why = "border denominator out of range"; return; }
            lo = lo * 64 / den - 1;
            hi = hi * 64 / den + 1;
            // the reference casts the rescaled border row to int16 (resizeYborder): content that would wrap it lies outside
            // the biased 16-bit window of the fast kernels, so such tables go to the generic kernel, which wraps like the reference
            if (lo < -32768 || hi > 32767) { m.why = "border row may wrap int16 after its division"; return; }
        }
        wmin = std::min(wmin, lo);
        wmax = std::max(wmax, hi);
    }
    m.workBias = int(-wmin);
    if (wmax + m.workBias > 65535) { m.why = "intermediate range wider than 16 bits"; return; }

    // ---- vertical blocks ----
    const int64_t blocks = (Y.D + 15) / 16;
    m.vBlock.assign(size_t(blocks) * 2, 0);
    m.vRow.assign(size_t(blocks) * 32, 0);
    std::vector<int64_t> lo(size_t(blocks), 0);
    m.vKMax = 1;
    for (int64_t b = 0; b < blocks; ++b) {
        int64_t rlo = 1ll << 40, rhi = -(1ll << 40);
        for (int64_t y = 16 * b; y < std::min<int64_t>(Y.D, 16 * b + 16); ++y) {
            const int32_t *c = &Y.coef[size_t(Y.row[size_t(y)]) * NY];
            const int64_t f = Y.first[size_t(y)];
            for (int i = 0; i < NY; ++i)
                if (c[i] != 0) {
                    if (f + i < 0 || f + i >= Y.S) { m.why = "vertical tap outside the image with non-zero weight"; return; }
                    rlo = std::min(rlo, f + i);
                    rhi = std::max(rhi, f + i);
                }
            const int den = Y.deno[size_t(Y.row[size_t(y)])];
            m.vRow[size_t(y) * 2] = den;
            m.vRow[size_t(y) * 2 + 1] = den > 1 ? int32_t(uint32_t((1ull << 32) / uint64_t(den) + 1)) : 0;
        }
        if (rhi < rlo) { m.why = "all-zero vertical block"; return; }
        lo[size_t(b)] = rlo;
        m.vBlock[size_t(b) * 2 + 1] = int32_t(rhi);
    }
    // the kernel's source FIFO only moves forward: first rows must not decrease from block to block
    for (int64_t b = blocks - 2; b >= 0; --b) lo[size_t(b)] = std::min(lo[size_t(b)], lo[size_t(b) + 1]);
    // Area / Linear: rows that carry a weight of 256 somewhere in the block need a second k slot (255 + 1)
    std::vector<std::vector<int64_t> > dup(static_cast<size_t>(blocks));
    if (!m.isSigned)
        for (int64_t b = 0; b < blocks; ++b)
            for (int64_t y = 16 * b; y < std::min<int64_t>(Y.D, 16 * b + 16); ++y) {
                const int32_t *c = &Y.coef[size_t(Y.row[size_t(y)]) * NY];
                for (int i = 0; i < NY; ++i)
                    if (c[i] == 256) {
                        const int64_t row = Y.first[size_t(y)] + i;
                        if (std::find(dup[size_t(b)].begin(), dup[size_t(b)].end(), row) == dup[size_t(b)].end()) dup[size_t(b)].push_back(row);
                    }
            }
    int chunksNeeded = 1;
    for (int64_t b = 0; b < blocks; ++b) {
        const int64_t rlo = lo[size_t(b)], rhi = m.vBlock[size_t(b) * 2 + 1];
        const int ks = int((rhi - rlo + 1 + int64_t(dup[size_t(b)].size()) + 31) / 32);
        if (ks > kMmaMaxKSteps) { m.why = "vertical kernel spans more than three 32-row k-steps per 16 destination rows"; return; }
        m.vKMax = std::max(m.vKMax, ks);
        m.vBlock[size_t(b) * 2] = int32_t(rlo);
        m.vBlock[size_t(b) * 2 + 1] = int32_t(rhi - rlo + 1);   // source rows the block reads; k beyond them meets zero coefficients
        chunksNeeded = std::max(chunksNeeded, int(floorTo(rhi, kMmaChunkRows) / kMmaChunkRows - floorTo(rlo, kMmaChunkRows) / kMmaChunkRows + 1));
    }
    m.nChunks = std::max(2, chunksNeeded);
    m.vFrag.assign(size_t(blocks) * m.vKMax * 128, 0);
    m.vRowMap.assign(size_t(blocks) * m.vKMax * 32, 0);
    for (int64_t b = 0; b < blocks; ++b) {
        const int nrows = m.vBlock[size_t(b) * 2 + 1];
        const int nslots = nrows + int(dup[size_t(b)].size());
        const int ks = (nslots + 31) / 32;
        int32_t *rmap = &m.vRowMap[size_t(b) * m.vKMax * 32];
        std::vector<int64_t> slotRow(size_t(m.vKMax) * 32);
        for (int k = 0; k < m.vKMax * 32; ++k) {
            const int64_t row = k < nrows ? lo[size_t(b)] + k : k < nslots ? dup[size_t(b)][size_t(k - nrows)] : lo[size_t(b)];
            slotRow[size_t(k)] = row;
            // chunk c = floor(row / 8) lives in FIFO slot c mod nChunks; a FIFO row is wcols bytes
            const int64_t c = floorTo(row, kMmaChunkRows) / kMmaChunkRows;
            const int64_t slot = ((c % m.nChunks) + m.nChunks) % m.nChunks;
            rmap[k] = int32_t((slot * kMmaChunkRows + (row - c * kMmaChunkRows)) * wcols);
        }
        // A[row][k]: coefficient of destination row 16 b + row for the source row of slot k
        auto coefAt = [&](int row, int64_t k) -> uint32_t {
            const int64_t y = 16 * b + row;
            if (y >= Y.D || k >= nslots) return 0u;
            const int64_t i = slotRow[size_t(k)] - Y.first[size_t(y)];
            if (i < 0 || i >= NY) return 0u;
            const int c = Y.coef[size_t(Y.row[size_t(y)]) * NY + size_t(i)];
            if (c == 256) return k < nrows ? 255u : 1u;   // split over the row's own slot and its extra slot
            return k < nrows ? uint32_t(c) & 0xffu : 0u;
        };
        for (int s = 0; s < ks; ++s)
            for (int lane = 0; lane < 32; ++lane) {
                const int g = lane >> 2, t = lane & 3;
                uint32_t *f = &m.vFrag[((size_t(b) * m.vKMax + s) * 32 + lane) * 4];
                for (int i = 0; i < 4; ++i) {
                    f[0] |= coefAt(g, 32 * s + 4 * t + i) << (8 * i);
                    f[1] |= coefAt(g + 8, 32 * s + 4 * t + i) << (8 * i);
                    f[2] |= coefAt(g, 32 * s + 16 + 4 * t + i) << (8 * i);
                    f[3] |= coefAt(g + 8, 32 * s + 16 + 4 * t + i) << (8 * i);
                }
            }
    }

    // ---- horizontal tiles ----
    const int64_t tiles = (X.D + 7) / 8;
    m.hTile.assign(size_t(tiles) * 2, 0);
    m.hCol.assign(size_t(tiles) * 16, 0);
    std::vector<int64_t> c0(size_t(tiles), 0), cEnd(size_t(tiles), 0);
    m.hKMax = 1;
    for (int64_t T = 0; T < tiles; ++T) {
        int64_t clo = 1ll << 40, chi = -(1ll << 40);
        for (int64_t d = 8 * T; d < std::min<int64_t>(X.D, 8 * T + 8); ++d) {
            const int r = X.row[size_t(d)];
            const int32_t *c = &X.coef[size_t(r) * NX];
            const int64_t f = X.first[size_t(d)];
            long long sum = 0;
            for (int i = 0; i < NX; ++i) {
                sum += c[i];
                if (c[i] != 0) {
                    if (f + i < 0 || f + i >= X.S) { m.why = "horizontal tap outside the image with non-zero weight"; return; }
                    clo = std::min(clo, f + i);
                    chi = std::max(chi, f + i);
                }
            }
            m.hCol[size_t(d) * 2] = int32_t((1ll << (p.shift - 1)) - (long long)m.workBias * sum);
            m.hCol[size_t(d) * 2 + 1] = X.deno[size_t(r)] * 64;
        }
        if (chi < clo) { m.why = "all-zero horizontal tile"; return; }
        clo = floorTo(clo, 8);
        const int ks = int((chi - clo + 32) / 32);
        if (ks > kMmaMaxKSteps) { m.why = "horizontal kernel spans more than three 32-column k-steps per 8 destination columns"; return; }
        m.hKMax = std::max(m.hKMax, ks);
        c0[size_t(T)] = clo;
        cEnd[size_t(T)] = clo + 32 * ks;
        m.hTile[size_t(T) * 2] = int32_t(clo);
        m.hTile[size_t(T) * 2 + 1] = ks;
    }
    // strips: the largest even tile count whose source window [xs, xs + wcols) holds every tile's k range
    // (tuning knob IQO_CUDA_MMA_TILE_MULT: strips of a multiple of that many tiles, so that the warps of a CTA get equal shares)
    static const int tileMult = [] { const char *e = getenv("IQO_CUDA_MMA_TILE_MULT"); const int v = e ? atoi(e) : 2; return (v == 4 || v == 8) ? v : 2; }();
    int st = int(std::min<int64_t>(32, (tiles + 1) & ~1ll));
    if (tileMult > 2 && st >= 2 * tileMult) st = st / tileMult * tileMult;
    for (; st >= 2; st -= (st > tileMult && tileMult > 2 ? tileMult : 2)) {
        bool ok = true;
        for (int64_t T0 = 0; T0 < tiles && ok; T0 += st) {
            const int64_t xs = floorTo(c0[size_t(T0)], 16);
            for (int64_t T = T0; T < std::min<int64_t>(tiles, T0 + st) && ok; ++T)
                if (c0[size_t(T)] < xs || cEnd[size_t(T)] - xs > wcols) ok = false;
        }
        if (ok) break;
    }
    if (st < 2) { m.why = "horizontal kernel wider than a warp strip"; return; }
    m.stripTiles = st;
    m.wcols = wcols;
    const int64_t strips = (tiles + st - 1) / st;
    m.stripXs.assign(size_t(strips), 0);
    for (int64_t s = 0; s < strips; ++s) m.stripXs[size_t(s)] = int32_t(floorTo(c0[size_t(s * st)], 16));
    m.hFrag.assign(size_t(tiles) * m.hKMax * 128, 0);
    for (int64_t T = 0; T < tiles; ++T) {
        const int ks = m.hTile[size_t(T) * 2 + 1];
        // B[k][n]: coefficient of destination column 8 T + n for the source column the kernel's A fragment holds at k:
        // k = 16 h + 4 t + i  <->  column c0 + 32 s + 16 h + {2 t, 2 t + 1, 8 + 2 t, 9 + 2 t}[i]
        auto coefAt = [&](int n, int64_t col) -> int {
            const int64_t d = 8 * T + n;
            if (d >= X.D) return 0;
            const int64_t i = col - X.first[size_t(d)];
            if (i < 0 || i >= NX) return 0;
            return X.coef[size_t(X.row[size_t(d)]) * NX + size_t(i)];
        };
        for (int s = 0; s < ks; ++s)
            for (int lane = 0; lane < 32; ++lane) {
                const int g = lane >> 2, t = lane & 3;
                uint32_t *f = &m.hFrag[((size_t(T) * m.hKMax + s) * 32 + lane) * 4];
                for (int h = 0; h < 2; ++h)
                    for (int i = 0; i < 4; ++i) {
                        const int off = (i & 1) + 2 * t + ((i >> 1) ? 8 : 0);
                        uint32_t l, hi;
                        const int cv = coefAt(g, c0[size_t(T)] + 32 * s + 16 * h + off);
                        if (m.isSigned) {
                            splitPlanes(cv, l, hi);
                        } else {
                            l = uint32_t(cv) & 0xffu;
                            hi = uint32_t(cv) >> 8;   // 0 .. 128
                        }
                        f[h] |= l << (8 * i);
                        f[2 + h] |= hi << (8 * i);
                    }
            }
    }
    m.eligible = true;
}

}  // namespace iqo_b200

// ---------------------------------------------------------------------------------------------
// Float ("SIMD-semantics") tables, SURVEY 8f-4
// ---------------------------------------------------------------------------------------------
namespace iqo_b200 {

namespace {

void floatAxis(const AxisPlan &a, int degree, uint64_t px, FloatAxis &f)
{
    const int N = a.N;
    f.coef.assign(size_t(a.numRows) * N, 0.0f);
    f.deno.assign(size_t(a.numRows), 0.0f);
    if (a.identity) {
        f.coef[0] = 1.0f;   // pass-through axis: the value itself
        return;
    }
    std::vector<float> w(static_cast<size_t>(N));
    for (int64_t t = 0; t < a.rD; ++t) {
        const float sum = lanczosPhase(degree, uint64_t(a.rS), uint64_t(a.rD), uint64_t(t), px, w);
        for (int i = 0; i < N; ++i) f.coef[size_t(t) * N + i] = w[size_t(i)] / sum;   // table[i] /= sumCoefs (float division)
    }
    for (int64_t d = 0; d < a.D; ++d) {
        const int r = a.row[size_t(d)];
        if (r < a.rD) continue;
        // border index: taps outside the image dropped, denominator = float sum of the remaining normalised coefficients
        // accumulated in tap order (src/IQOLanczosResizerImpl_AVX512.cpp:337-355)
        const int64_t t = d % a.rD, first = a.first[size_t(d)];
        float den = 0.0f;
        for (int i = 0; i < N; ++i) {
            const bool inside = first + i >= 0 && first + i < a.S;
            const float c = inside ? f.coef[size_t(t) * N + i] : 0.0f;
            f.coef[size_t(r) * N + i] = c;
            if (inside) den += c;
        }
        f.deno[size_t(r)] = den;
    }
}

}  // namespace

void buildFloatPlan(const Plan &p, FloatPlan &f)
{
    f.eligible = false;
    f.why.clear();
    if (p.kind != kLanczos) { f.why = "the float mode exists for Lanczos only"; return; }
    floatAxis(p.x, int(p.degree), p.pxScale, f.x);
    floatAxis(p.y, int(p.degree), p.pxScale, f.y);
    for (size_t r = size_t(p.x.rD); r < f.x.deno.size(); ++r)
        if (!p.x.identity && !(f.x.deno[r] != 0.0f)) { f.why = "a border column's in-range coefficients sum to 0"; return; }
    for (size_t r = size_t(p.y.rD); r < f.y.deno.size(); ++r)
        if (!p.y.identity && !(f.y.deno[r] != 0.0f)) { f.why = "a border row's in-range coefficients sum to 0"; return; }
    f.eligible = true;
}

}  // namespace iqo_b200
