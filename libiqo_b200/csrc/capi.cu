// C ABI of the CUDA backend (include/iqo_cuda.h).  Host logic only; kernels live in kernels.cu.
#include "../../include/iqo_cuda.h"

#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <memory>
#include <mutex>
#include <new>
#include <stdexcept>
#include <string>
#include <thread>
#include <tuple>
#include <vector>

#include "kernels.cuh"
#include "plan.hpp"

using namespace iqo_b200;

namespace {

thread_local std::string t_lastError;

int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    t_lastError = buf;
    return code;
}

// C++ exceptions (std::bad_alloc from the planner's vectors, std::system_error from std::thread) must not
// unwind through the C ABI: IQO_GUARD_BEGIN / IQO_GUARD_END map them to status codes.
#define IQO_GUARD_BEGIN try {
#define IQO_GUARD_END                                                                          \
    }                                                                                          \
    catch (const std::bad_alloc &)                                                             \
    {                                                                                          \
        return fail(IQO_CUDA_E_NOMEM, "host allocation failed (std::bad_alloc)");              \
    }                                                                                          \
    catch (const std::exception &e_)                                                           \
    {                                                                                          \
        return fail(IQO_CUDA_E_CUDA, "unexpected C++ exception: %s", e_.what());               \
    }                                                                                          \
    catch (...)                                                                                \
    {                                                                                          \
        return fail(IQO_CUDA_E_CUDA, "unexpected C++ exception");                              \
    }

#define CUDA_TRY(expr)                                                                         \
    do {                                                                                       \
        cudaError_t e_ = (expr);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            return fail(IQO_CUDA_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                                   \
    } while (0)

struct DeviceGuard {
    int prev;
    bool ok;
    explicit DeviceGuard(int dev) : prev(-1), ok(false)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) return;
        ok = (prev == dev) || (cudaSetDevice(dev) == cudaSuccess);
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

struct AxisTables {
    int32_t *first, *row, *coef, *deno;
    AxisTables() : first(0), row(0), coef(0), deno(0) {}
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
EncodeTiledFn encodeTiled()
{
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void *p = 0;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            return 0;
        }
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

#ifndef IQO_TMA_PROMO_DEFAULT
#define IQO_TMA_PROMO_DEFAULT 1
#endif
#ifndef IQO_MMA_WARPS_DEFAULT
#define IQO_MMA_WARPS_DEFAULT 4
#endif
#ifndef IQO_MMA_AUTO_DEFAULT
#define IQO_MMA_AUTO_DEFAULT 1
#endif
#ifndef IQO_MMA_WCOLS_DEFAULT
#define IQO_MMA_WCOLS_DEFAULT 272
#endif
#ifndef IQO_STREAM_TMA_DEFAULT
#define IQO_STREAM_TMA_DEFAULT 1
#endif

// L2 promotion of the streaming kernels' tensor maps (tuning knob IQO_CUDA_TMA_PROMO: 0 none, 1 64 B, 2 128 B, 3 256 B)
CUtensorMapL2promotion streamPromotion()
{
    static const int v = [] { const char *e = getenv("IQO_CUDA_TMA_PROMO"); return e ? atoi(e) : IQO_TMA_PROMO_DEFAULT; }();
    return v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : v == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : v == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                                                                                 : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
}

size_t alignUp(size_t v, size_t a)
{
    return (v + a - 1) / a * a;
}

}  // namespace

// Everything that depends only on (device, kind, degree, sizes, pxScale): host plan and device
// tables.  Immutable after construction, shared by every handle of that shape and kept in a small
// cache, so that constructing a resizer again (the reference's benchmark builds its resizers
// inside the timed loop, benchmark/benchmark.cpp:214-227) costs a map lookup.
struct SharedPlan {
    Plan plan;
    HalfPlan half;
    SmallPlan small;
    RatioPlan ratio;
    LStreamPlan lstream;
    MmaPlan mma;
    FloatPlan flt;        // optional float (SIMD-semantics) mode, built on first use
    bool fltBuilt;
    float *fCoefX, *fCoefY, *fDenoX, *fDenoY;
    std::mutex fltMu;
    PackedPlan packed;
    GenericGeom geom;
    PackedGeom pgeom;
    int sms;              // multiprocessors of the device (launch-size heuristics; 148 on B200)
    int device;
    AxisTables tx, ty;
    uint32_t *dBorderY, *dMagicY, *dSBorderY;
    int32_t *dBorderX, *dBorderXo;
    // packed kernel tables
    int32_t *pFirstY, *pNtapY, *pCoefOffY, *pRecX;
    uint32_t *pMagicY, *pCwX;
    int32_t *rRowRec;     // rational-ratio streaming path
    int32_t *gRowRec;     // general Lanczos streaming path
    int32_t *sRowsY;      // small-kernel path
    uint32_t *sMagicY;
    // tensor-path kernel tables
    int32_t *mVBlock, *mVRow, *mVRowMap, *mStripXs, *mHTile, *mHCol;
    uint32_t *mVFrag, *mHFrag;
    SharedPlan()
        : device(0), dBorderY(0), dMagicY(0), dSBorderY(0), dBorderX(0), dBorderXo(0), pFirstY(0), pNtapY(0), pCoefOffY(0), pRecX(0), pMagicY(0), pCwX(0), rRowRec(0), gRowRec(0), sRowsY(0), sMagicY(0), fltBuilt(false), fCoefX(0), fCoefY(0), fDenoX(0), fDenoY(0), mVBlock(0), mVRow(0), mVRowMap(0), mStripXs(0), mHTile(0), mHCol(0), mVFrag(0), mHFrag(0)
    {
    }
    ~SharedPlan();
};

// Per-handle mutable resources (streams, staging buffers); recycled through a pool on destroy.
struct Workspace {
    int device;
    cudaStream_t stream[2];
    uint8_t *dSrc[2], *dDst[2];
    size_t srcCap, dstCap;  // bytes per slot
    // pinned host staging for callers that pass pageable memory (two slots, like the device staging)
    uint8_t *hSrc[2], *hDst[2];
    size_t hSrcCap, hDstCap;
    cudaEvent_t slotDone[2];
    Workspace() : device(0), srcCap(0), dstCap(0), hSrcCap(0), hDstCap(0)
    {
        for (int i = 0; i < 2; ++i) {
            stream[i] = 0;
            dSrc[i] = dDst[i] = 0;
            hSrc[i] = hDst[i] = 0;
            slotDone[i] = 0;
        }
    }
};

struct iqo_cuda_resizer {
    std::shared_ptr<SharedPlan> sp;
    Workspace *ws;
    Plan &plan;
    HalfPlan &half;
    AxisTables &tx, &ty;
    GenericGeom &geom;
    uint32_t *&dBorderY, *&dMagicY;
    int32_t *&dBorderX;
    cudaStream_t *stream;
    uint8_t **dSrc, **dDst;
    int device;
    bool useTma, useStream, forceStream, useMma, forceMma;
    int arithmetic;             // IQO_CUDA_ARITH_*
    int path;
    const char *lastKernel;
    size_t srcPitch, dstPitch;  // device pitches of the staging frames
    size_t slotFrames;          // frames each staging slot currently holds

    iqo_cuda_resizer(const std::shared_ptr<SharedPlan> &s, Workspace *w)
        : sp(s), ws(w), plan(s->plan), half(s->half), tx(s->tx), ty(s->ty), geom(s->geom), dBorderY(s->dBorderY),
          dMagicY(s->dMagicY), dBorderX(s->dBorderX), stream(w->stream), dSrc(w->dSrc), dDst(w->dDst), device(s->device),
          useTma(true), useStream(true), forceStream(false), useMma(true), forceMma(false), arithmetic(IQO_CUDA_ARITH_FIXED), path(IQO_CUDA_PATH_AUTO), lastKernel("none"), srcPitch(0), dstPitch(0), slotFrames(0)
    {
    }
};

namespace {

int uploadAxis(const AxisPlan &a, AxisTables &t)
{
    CUDA_TRY(cudaMalloc(&t.first, a.first.size() * 4));
    CUDA_TRY(cudaMalloc(&t.row, a.row.size() * 4));
    CUDA_TRY(cudaMalloc(&t.coef, a.coef.size() * 4));
    CUDA_TRY(cudaMalloc(&t.deno, a.deno.size() * 4));
    CUDA_TRY(cudaMemcpy(t.first, a.first.data(), a.first.size() * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(t.row, a.row.data(), a.row.size() * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(t.coef, a.coef.data(), a.coef.size() * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(t.deno, a.deno.data(), a.deno.size() * 4, cudaMemcpyHostToDevice));
    return IQO_CUDA_OK;
}

void freeAxis(AxisTables &t)
{
    cudaFree(t.first);
    cudaFree(t.row);
    cudaFree(t.coef);
    cudaFree(t.deno);
}

}  // namespace

SharedPlan::~SharedPlan()
{
    DeviceGuard guard(device);
    freeAxis(tx);
    freeAxis(ty);
    cudaFree(dBorderY);
    cudaFree(dMagicY);
    cudaFree(dBorderX);
    cudaFree(dSBorderY);
    cudaFree(dBorderXo);
    cudaFree(pFirstY);
    cudaFree(pNtapY);
    cudaFree(pCoefOffY);
    cudaFree(pRecX);
    cudaFree(sRowsY);
    cudaFree(rRowRec);
    cudaFree(gRowRec);
    cudaFree(sMagicY);
    cudaFree(pMagicY);
    cudaFree(pCwX);
    cudaFree(mVBlock);
    cudaFree(mVRow);
    cudaFree(mVRowMap);
    cudaFree(fCoefX);
    cudaFree(fCoefY);
    cudaFree(fDenoX);
    cudaFree(fDenoY);
    cudaFree(mStripXs);
    cudaFree(mHTile);
    cudaFree(mHCol);
    cudaFree(mVFrag);
    cudaFree(mHFrag);
    cudaGetLastError();
}

namespace {

template <typename T>
bool uploadVec(T *&dptr, const std::vector<T> &v)
{
    const size_t bytes = std::max<size_t>(sizeof(T), v.size() * sizeof(T));
    if (cudaMalloc(&dptr, bytes) != cudaSuccess) return false;
    return v.empty() || cudaMemcpy(dptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice) == cudaSuccess;
}

AxisDev axisDev(const AxisPlan &a, const AxisTables &t)
{
    AxisDev d;
    d.first = t.first;
    d.row = t.row;
    d.coef = t.coef;
    d.deno = t.deno;
    d.N = a.N;
    d.S = int(a.S);
    d.D = int(a.D);
    return d;
}

// true when p is device memory (or managed); host / unregistered otherwise
bool isDevicePointer(const void *p)
{
    cudaPointerAttributes attr;
    cudaError_t e = cudaPointerGetAttributes(&attr, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
}

static int64_t floorDivI64(int64_t a, int64_t b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }
// Linear ratios the streaming kernel is instantiated for (items of 4 rS source columns)
bool linearUpRatio(const AxisPlan &X)
{
    if (X.identity || X.N != 2) return false;
    const bool known = (X.rS == 1 && (X.rD == 2 || X.rD == 3 || X.rD == 4)) || (X.rS == 2 && (X.rD == 3 || X.rD == 5)) ||
                       (X.rS == 3 && (X.rD == 4 || X.rD == 2)) || (X.rS == 4 && (X.rD == 5 || X.rD == 3));
    if (!known || X.S % (4 * X.rS) != 0) return false;
    // the kernel's compile-time first tap floor(((2 d + 1) S - D) / (2 D)) must be the planner's (it is for these ratios;
    // from 2:1 on the reference's iterator starts elsewhere): checked on the first items
    for (int64_t d = 1; d < std::min<int64_t>(X.D - 1, 64); ++d)
        if (X.first[size_t(d)] != floorDivI64((2 * d + 1) * X.S - X.D, 2 * X.D)) return false;
    return true;
}
// Area reductions the streaming kernel is instantiated for; *nxEff = taps up to the last one that is non-zero in some phase
bool areaDownRatio(const AxisPlan &X, int *nxEff)
{
    if (X.identity || X.rS <= X.rD || X.rD > 3 || X.numRows < int(X.rD)) return false;
    int n = 0;
    for (int ph = 0; ph < int(X.rD); ++ph)
        for (int k = 0; k < X.N; ++k)
            if (X.coef[size_t(ph) * X.N + k] != 0) n = std::max(n, k + 1);
    if (nxEff) *nxEff = n;
    return areaDownHasKernel(int(X.rS), int(X.rD), n) && X.S % areaDownItemColumns(int(X.rS), int(X.rD)) == 0;
}
const char *linearUpName(const AxisPlan &X)
{
    return X.rS == 1   ? (X.rD == 2 ? "linear_up2" : X.rD == 3 ? "linear_up3" : "linear_up4")
           : X.rS == 2 ? (X.rD == 3 ? "linear_up_2_3" : "linear_up_2_5")
           : X.rS == 3 ? (X.rD == 4 ? "linear_up_3_4" : "linear_3_2")
                       : (X.rD == 5 ? "linear_up_4_5" : "linear_4_3");
}

// AUTO gives a launch to the tensor-path kernel when it has enough warps (strip x band) to fill the device
bool mmaWorthIt(const iqo_cuda_resizer *r, size_t nFrames, size_t dstRows)
{
    static const int envAuto = [] { const char *e = getenv("IQO_CUDA_MMA_AUTO"); return e ? atoi(e) : IQO_MMA_AUTO_DEFAULT; }();
    if (!envAuto) return false;
    const MmaPlan &mp = r->sp->mma;
    if (!mp.eligible || mp.stripTiles <= 0) return false;
    const long long strips = ((r->plan.x.D + 7) / 8 + mp.stripTiles - 1) / mp.stripTiles;
    return strips * (long long)((dstRows + 31) / 32) * (long long)nFrames >= 4ll * r->sp->sms;
}

int launch(iqo_cuda_resizer *r, size_t nFrames, size_t dstRow0, size_t dstRows, size_t srcRow0, size_t srcRows,
           size_t srcSt, size_t srcFrameStride, const uint8_t *src,
           size_t dstSt, size_t dstFrameStride, uint8_t *dst, cudaStream_t stream)
{
    ResizeArgs a;
    a.x = axisDev(r->plan.x, r->tx);
    a.y = axisDev(r->plan.y, r->ty);
    a.src = src;
    a.dst = dst;
    a.srcPitch = (long long)srcSt;
    a.dstPitch = (long long)dstSt;
    a.srcFrameStride = (long long)srcFrameStride;
    a.dstFrameStride = (long long)dstFrameStride;
    a.nFrames = int(nFrames);
    a.srcRow0 = int(srcRow0);
    a.srcRows = int(srcRows);
    a.dstRow0 = int(dstRow0);
    a.dstRows = int(dstRows);
    a.shift = r->plan.shift;
    a.lanczos = r->plan.kind == kLanczos;
    a.workSigned = r->plan.workSigned;
    const bool whole = dstRow0 == 0 && dstRows == size_t(r->plan.y.D) && srcRow0 == 0;
    if (r->arithmetic == IQO_CUDA_ARITH_SIMD_FLOAT) {
        // opt-in float mode: one kernel, outside the parity contract
        SharedPlan &fp = *r->sp;
        FloatArgs q;
        q.x = a.x;
        q.y = a.y;
        q.coefX = fp.fCoefX;
        q.coefY = fp.fCoefY;
        q.denoX = fp.fDenoX;
        q.denoY = fp.fDenoY;
        q.src = src;
        q.dst = dst;
        q.srcPitch = a.srcPitch;
        q.dstPitch = a.dstPitch;
        q.srcFrameStride = a.srcFrameStride;
        q.dstFrameStride = a.dstFrameStride;
        q.nFrames = a.nFrames;
        q.srcRow0 = a.srcRow0;
        q.srcRows = a.srcRows;
        q.dstRow0 = a.dstRow0;
        q.dstRows = a.dstRows;
        r->lastKernel = "float_simd_semantics";
        CUDA_TRY(launchFloat(q, r->geom, stream));
        return IQO_CUDA_OK;
    }
    // Lanczos, both passes on the integer tensor path (any ratio, row bands included)
    auto tryMma = [&]() -> int {
        if (!(r->useMma && r->sp->mma.eligible && ((uintptr_t)src % 16) == 0 && srcSt % 16 == 0 && (nFrames == 1 || srcFrameStride % 16 == 0))) return 0;
        {
        const SharedPlan &sp = *r->sp;
        const MmaPlan &mp = sp.mma;
        MmaArgs q;
        q.dstPitch = (long long)dstSt;
        q.dstFrameStride = (long long)dstFrameStride;
        q.DW = int(r->plan.x.D);
        q.srcRow0 = int(srcRow0);
        q.dstRow0 = int(dstRow0);
        q.dstRows = int(dstRows);
        q.stripTiles = mp.stripTiles;
        q.wcols = mp.wcols;
        q.vKMax = mp.vKMax;
        q.hKMax = mp.hKMax;
        q.nChunks = mp.nChunks;
        static const int envMmaWarps = [] { const char *e = getenv("IQO_CUDA_MMA_WARPS"); return e ? atoi(e) : IQO_MMA_WARPS_DEFAULT; }();
        q.warps = (envMmaWarps == 1 || envMmaWarps == 2 || envMmaWarps == 4) ? envMmaWarps : IQO_MMA_WARPS_DEFAULT;
        q.workBias = mp.workBias;
        q.mbY = int(r->plan.y.mainBegin);
        q.meY = int(r->plan.y.mainEnd);
        q.mbX = int(r->plan.x.mainBegin);
        q.meX = int(r->plan.x.mainEnd);
        q.dstVec = (((uintptr_t)dst % 16) == 0 && dstSt % 16 == 0 && dstFrameStride % 16 == 0) ? 2
                   : (((uintptr_t)dst % 8) == 0 && dstSt % 8 == 0 && dstFrameStride % 8 == 0) ? 1 : 0;
        q.vBlock = reinterpret_cast<const int2 *>(sp.mVBlock);
        q.vFrag = reinterpret_cast<const uint4 *>(sp.mVFrag);
        q.vRow = reinterpret_cast<const int2 *>(sp.mVRow);
        q.vRowMap = sp.mVRowMap;
        q.isSigned = mp.isSigned ? 1 : 0;
        q.stripXs = sp.mStripXs;
        q.hTile = reinterpret_cast<const int2 *>(sp.mHTile);
        q.hFrag = reinterpret_cast<const uint4 *>(sp.mHFrag);
        q.hCol = reinterpret_cast<const int2 *>(sp.mHCol);
        // bands of whole 16-row blocks: enough warps to fill the device several times over
        const long long strips = ((q.DW + 7) / 8 + q.stripTiles - 1) / q.stripTiles;
        const long long blocks = (long long)((dstRow0 + dstRows + 15) / 16 - dstRow0 / 16);
        int bandBlocks = 16;
        static const int envBand = [] { const char *e = getenv("IQO_CUDA_MMA_BAND_BLOCKS"); return e ? atoi(e) : 0; }();
        if (envBand > 0) bandBlocks = envBand;
        else
            while (bandBlocks > 4 && strips * ((blocks + bandBlocks - 1) / bandBlocks) * (long long)nFrames < 4ll * sp.sms * 4) bandBlocks /= 2;
        q.bandBlocks = bandBlocks;
        const size_t smem = mmaSmemBytes(q.wcols, q.stripTiles, q.nChunks, q.hKMax);
        if (smem <= 200 * 1024 && (blocks + bandBlocks - 1) / bandBlocks <= 65535) {
            bool ok = true;
            for (size_t f0 = 0; f0 < nFrames && ok; f0 += 65535) {
                const size_t nf = std::min<size_t>(65535, nFrames - f0);
                q.nFrames = int(nf);
                q.dst = dst + f0 * dstFrameStride;
                const uint8_t *fsrc = src + f0 * srcFrameStride;
                CUtensorMap tmap;
                // 16-bit view of the source rows present in the buffer: (x / 2, y, frame), box (wcols / 2) x 8 x 1
                const cuuint64_t dims[3] = {cuuint64_t(r->plan.x.S / 2), cuuint64_t(srcRows), cuuint64_t(nf)};
                const cuuint64_t strides[2] = {cuuint64_t(srcSt), cuuint64_t(nf > 1 ? srcFrameStride : srcSt * srcRows)};
                const cuuint32_t box[3] = {cuuint32_t(q.wcols / 2), cuuint32_t(kMmaChunkRows), 1};
                const cuuint32_t estr[3] = {1, 1, 1};
                CUresult cr = encodeTiled()(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, const_cast<uint8_t *>(fsrc), dims, strides,
                                            box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                            streamPromotion(), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (cr != CUDA_SUCCESS) {
                    ok = false;   // e.g. a stride the descriptor cannot express: other kernels take the launch
                    if (f0 != 0) return fail(IQO_CUDA_E_CUDA, "cuTensorMapEncodeTiled failed (%d)", int(cr));
                    break;
                }
                // reductions of the RD | 8 family whose 3:2-kernel pattern exists (cfg1): tensor-path vertical pass, the
                // compile-time dp2a horizontal pass of the 3:2 kernel (IQO_CUDA_MMA_DP2A=0: the all-mma kernel)
                static const int envMmaDp2a = [] { const char *e = getenv("IQO_CUDA_MMA_DP2A"); return e ? atoi(e) : 1; }();
                const RatioPlan &rp = sp.ratio;
                if (envMmaDp2a && r->plan.kind == kLanczos && rp.eligible && mmaRatioHasKernel(rp.RS, rp.RD, rp.NX, rp.odd) &&
                    (rp.c0 & 1) == rp.odd && q.DW % 8 == 0 && rp.RD <= 4) {
                    RatioArgs ra;
                    memset(&ra, 0, sizeof ra);
                    ra.RS = rp.RS, ra.RD = rp.RD, ra.NX = rp.NX, ra.odd = rp.odd;
                    ra.tailZeros = rp.tailZeros;
                    ra.c0 = rp.c0;
                    for (int ph = 0; ph < rp.RD && ph < 4; ++ph)
                        for (int par = 0; par < 2; ++par)
                            for (int j = 0; j < 7; ++j) ra.cwX[ph][par][j] = rp.cwX[(size_t(ph) * 2 + par) * 7 + j];
                    ra.gx = a.x;
                    r->lastKernel = "lanczos_mma_dp2a";
                    CUDA_TRY(launchMmaRatio(q, tmap, ra, stream));
                    continue;
                }
                r->lastKernel = r->plan.kind == kLanczos ? "lanczos_mma" : r->plan.kind == kArea ? "area_mma" : "linear_mma";
                CUDA_TRY(launchMma(q, tmap, stream));
            }
            if (ok) return 1;
        }
        }
        return 0;
    };
    if (r->forceMma) {
        const int rc = tryMma();
        if (rc != 0) return rc < 0 ? rc : IQO_CUDA_OK;
    }
    // 2:1 Lanczos with at most four non-zero taps per axis (YUV420 chroma planes): streaming kernel
    if (r->path == IQO_CUDA_PATH_AUTO && r->sp->small.eligible && whole && ((uintptr_t)src % 16) == 0 && srcSt % 16 == 0 &&
        srcFrameStride % 16 == 0 && ((uintptr_t)dst % 8) == 0 && dstSt % 8 == 0 && dstFrameStride % 8 == 0) {
        const SmallPlan &sm = r->sp->small;
        SmallArgs q;
        q.srcPitch = (long long)srcSt;
        q.dstPitch = (long long)dstSt;
        q.srcFrameStride = (long long)srcFrameStride;
        q.dstFrameStride = (long long)dstFrameStride;
        q.SW = int(r->plan.x.S);
        q.SH = int(r->plan.y.S);
        q.DW = int(r->plan.x.D);
        q.DH = int(r->plan.y.D);
        q.TY = sm.TY;
        q.cy0 = sm.cy0;
        q.NW = sm.NW;
        q.wbase = sm.wbase;
        memcpy(q.cY, sm.cY, sizeof q.cY);
        memcpy(q.cwX, sm.cwX, sizeof q.cwX);
        q.accInit = sm.accInit;
        q.workBias = sm.workBias;
        q.mbX = int(r->plan.x.mainBegin);
        q.meX = int(r->plan.x.mainEnd);
        q.mbY = int(r->plan.y.mainBegin);
        q.meY = int(r->plan.y.mainEnd);
        q.rowsY = r->sp->sRowsY;
        q.magicY = r->sp->sMagicY;
        q.gx = a.x;
        q.gy = a.y;
        r->lastKernel = "half_small";
        for (size_t f0 = 0; f0 < nFrames; f0 += 65535) {
            q.nFrames = int(std::min<size_t>(65535, nFrames - f0));
            q.src = src + f0 * srcFrameStride;
            q.dst = dst + f0 * dstFrameStride;
            CUDA_TRY(launchSmall(q, stream));
        }
        return IQO_CUDA_OK;
    }
    if (r->path == IQO_CUDA_PATH_AUTO && r->half.eligible && whole &&
        ((uintptr_t)src % 4) == 0 && srcSt % 4 == 0 && srcFrameStride % 4 == 0) {
        const HalfPlan &hp = r->half;
        HalfArgs h;
        h.src = src;
        h.dst = dst;
        h.srcPitch = (long long)srcSt;
        h.dstPitch = (long long)dstSt;
        h.srcFrameStride = (long long)srcFrameStride;
        h.dstFrameStride = (long long)dstFrameStride;
        h.SW = int(r->plan.x.S);
        h.SH = int(r->plan.y.S);
        h.DW = int(r->plan.x.D);
        h.DH = int(r->plan.y.D);
        h.nFrames = int(nFrames);
        // tile height: the image is cut into equal tiles of at most maxRows rows (64 -> 256-thread
        // CTAs with 4 strips, 32 -> 128-thread CTAs with 2 strips)
        int maxRows = 32;  // measured on B200: 3.24 ms vs 3.67 ms per 4096 1080p frames
        static const int envTileRows = [] { const char *e = getenv("IQO_CUDA_HALF_TILE_ROWS"); return e ? atoi(e) : 0; }();
        if (envTileRows > 0) maxRows = (envTileRows <= 32) ? 32 : 64;  // tuning knob, read once
        const int tiles = (h.DH + maxRows - 1) / maxRows;
        h.tileRows = 2 * ((h.DH + 2 * tiles - 1) / (2 * tiles));
        h.dstVec = ((uintptr_t)dst % 8) == 0 && dstSt % 8 == 0 && dstFrameStride % 8 == 0;
        h.qmin = hp.qmin;
        h.NG = hp.NG;
        memcpy(h.cwY, hp.cwY, sizeof h.cwY);
        h.borderY = r->dBorderY;
        h.rowY = r->ty.row;
        h.denoY = r->ty.deno;
        h.mbY = int(r->plan.y.mainBegin);
        h.meY = int(r->plan.y.mainEnd);
        h.workBias = hp.workBias;
        h.NWX = hp.NWX;
        h.symmetric = hp.symmetric;
        h.endsHi = hp.symmetric ? ((hp.cwXs[hp.NWX / 2] >> 16) != 0) : 1;
        memcpy(h.cwX, hp.cwX, sizeof h.cwX);
        memcpy(h.cwXs, hp.cwXs, sizeof h.cwXs);
        h.accInit = hp.accInit;
        h.mbX = int(r->plan.x.mainBegin);
        h.meX = int(r->plan.x.mainEnd);
        h.magicY = r->dMagicY;
        h.borderX = r->dBorderX;
        h.NX = r->plan.x.N;
        h.zero = 0;
        h.bandPairs = 0;
        h.delta = 0;
        h.zmask = 0;
        h.NXH = 0;
        h.skipHi0 = 0;
        memset(h.cwXo, 0, sizeof h.cwXo);
        const int boxRows = 4 * (h.tileRows / 2 + hp.NG - 1);
        const bool tmaOk = r->useTma && encodeTiled() != 0 && boxRows <= halfSourceRowsMax() &&
                           ((uintptr_t)src % 16) == 0 && srcSt % 16 == 0 && (nFrames == 1 || srcFrameStride % 16 == 0);
        // streaming variant (a warp per column strip and row band): source rows are copied as aligned 16-byte chunks
        // A warp per strip needs a big launch to fill the GPU (measured cross-over against the tiled kernel: about 8000
        // warps at the shortest band = 54 per SM, i.e. ~64 frames of 1080p); smaller launches start faster on the tiled kernel.
        const long long warpsMin = (long long)((h.DW + 119) / 120) * (((h.DH + 1) / 2 + 23) / 24) * (long long)nFrames;
        const bool streamOk = r->useStream && hp.sEligible && h.SW % 8 == 0 && ((uintptr_t)src % 16) == 0 && srcSt % 16 == 0 &&
                              (nFrames == 1 || srcFrameStride % 16 == 0) && (r->forceStream || warpsMin >= 54ll * r->sp->sms);
        if (streamOk) {
            // bands: enough warps to fill the device several times over, but long enough that the
            // ring refill and the re-read halo rows of a band stay small
            const long long strips = (h.DW + 119) / 120, pairs = (h.DH + 1) / 2;
            int bandPairs = 144;
            static const int envBandPairs = [] { const char *e = getenv("IQO_CUDA_STREAM_BAND_PAIRS"); return e ? atoi(e) : 0; }();
            if (envBandPairs > 0) bandPairs = envBandPairs;  // tuning knob, read once
            const long long wantWarps = 6ll * r->sp->sms * 20;
            while (bandPairs > 24 && strips * ((pairs + bandPairs - 1) / bandPairs) * (long long)nFrames < wantWarps) bandPairs /= 2;
            const long long bands = (pairs + bandPairs - 1) / bandPairs;
            h.bandPairs = int((pairs + bands - 1) / bands);
            h.tileShift = 0;
            h.qmin = hp.sQmin;
            h.NG = hp.sNG;
            memcpy(h.cwY, hp.sCwY, sizeof h.cwY);
            h.borderY = r->sp->dSBorderY;
            h.borderX = r->sp->dBorderXo;
            h.delta = hp.sDelta;
            h.zmask = hp.sZ;
            h.NXH = hp.NXH;
            h.skipHi0 = hp.skipHi0;
            memcpy(h.cwXo, hp.cwXo, sizeof h.cwXo);
            r->lastKernel = hp.symmetric ? "half_sym_stream" : "half_stream";
            // source FIFO fed by TMA (one box per turn) unless switched off or the descriptor cannot be built
            static const int envStreamTma = [] { const char *e = getenv("IQO_CUDA_STREAM_TMA"); return e ? atoi(e) : IQO_STREAM_TMA_DEFAULT; }();
            for (size_t f0 = 0; f0 < nFrames; f0 += 65535) {
                const size_t nf = std::min<size_t>(65535, nFrames - f0);
                h.src = src + f0 * srcFrameStride;
                h.dst = dst + f0 * dstFrameStride;
                h.nFrames = int(nf);
                CUtensorMap smap;
                bool tma = envStreamTma != 0 && r->useTma && encodeTiled() != 0;
                if (tma) {
                    // 16-bit view (a box dimension holds at most 256 elements): (x / 2, y, frame), box 136 x 4 NG x 1
                    const cuuint64_t dims[3] = {cuuint64_t(h.SW / 2), cuuint64_t(h.SH), cuuint64_t(nf)};
                    const cuuint64_t strides[2] = {cuuint64_t(srcSt), cuuint64_t(nf > 1 ? srcFrameStride : srcSt * size_t(h.SH))};
                    const cuuint32_t box[3] = {136, cuuint32_t(halfStreamBoxRows(h.NG)), 1};
                    const cuuint32_t estr[3] = {1, 1, 1};
                    CUresult cr = encodeTiled()(&smap, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, const_cast<uint8_t *>(h.src), dims, strides,
                                                box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                streamPromotion(), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                    if (cr != CUDA_SUCCESS) tma = false;
                }
                CUDA_TRY(launchHalfStream(h, tma ? &smap : 0, stream));
            }
            return IQO_CUDA_OK;
        }
        for (size_t f0 = 0; f0 < nFrames; f0 += 65535) {
            const size_t nf = std::min<size_t>(65535, nFrames - f0);
            h.src = src + f0 * srcFrameStride;
            h.dst = dst + f0 * dstFrameStride;
            h.nFrames = int(nf);
            CUtensorMap tmap;
            bool tma = tmaOk;
            if (tma) {
                // 3-D tensor (x, y, frame) of bytes; box = 256 columns x boxRows rows x 1 frame.
                // Out-of-image coordinates (negative too) are filled with zeros by the hardware.
                const cuuint64_t dims[3] = {cuuint64_t(h.SW), cuuint64_t(h.SH), cuuint64_t(nf)};
                const cuuint64_t strides[2] = {cuuint64_t(srcSt), cuuint64_t(nf > 1 ? srcFrameStride : srcSt * size_t(h.SH))};
                const cuuint32_t box[3] = {256, cuuint32_t(boxRows), 1};
                const cuuint32_t estr[3] = {1, 1, 1};
                CUresult cr = encodeTiled()(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t *>(h.src), dims, strides,
                                            box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (cr != CUDA_SUCCESS) tma = false;  // e.g. a stride the descriptor cannot express
            }
            h.tileShift = tma ? 4 : 0;
            r->lastKernel = tma ? (hp.symmetric ? "half_sym_tma" : "half_tma") : (hp.symmetric ? "half_sym" : "half");
            CUDA_TRY(launchHalf(h, tma ? &tmap : 0, boxRows, stream));
        }
        return IQO_CUDA_OK;
    }
    const SharedPlan &sp = *r->sp;
    // Area 2:1 on both axes: pure streaming kernel (needs whole 16-byte source chunks per 8 pixels)
    {
        const AxisPlan &X = r->plan.x, &Y = r->plan.y;
        if (r->path == IQO_CUDA_PATH_AUTO && r->plan.kind == kArea && whole && X.rD == 1 && X.rS == 2 && Y.rD == 1 && Y.rS == 2 &&
            X.N == 2 && Y.N == 2 && Y.coef[0] < 256 && Y.coef[1] < 256 && ((uintptr_t)src % 16) == 0 && srcSt % 16 == 0 &&
            srcFrameStride % 16 == 0 && ((uintptr_t)dst % 8) == 0 && dstSt % 8 == 0 && dstFrameStride % 8 == 0 &&
            X.S % 16 == 0) {
            r->lastKernel = "area2";
            for (size_t f0 = 0; f0 < nFrames; f0 += 65535) {
                const int nf = int(std::min<size_t>(65535, nFrames - f0));
                CUDA_TRY(launchArea2(src + f0 * srcFrameStride, dst + f0 * dstFrameStride, (long long)srcSt, (long long)dstSt,
                                     (long long)srcFrameStride, (long long)dstFrameStride, int(X.D), int(Y.D), nf, &Y.coef[0],
                                     &X.coef[0], stream));
            }
            return IQO_CUDA_OK;
        }
    }
    // Area reductions at 3:2, 4:3, 2:1, 5:2, 3:1 or 4:1 on X (any ratio on Y): streaming kernel
    {
        const AxisPlan &X = r->plan.x, &Y = r->plan.y;
        int nxEff = 0;
        if (r->path == IQO_CUDA_PATH_AUTO && r->plan.kind == kArea && whole && areaDownRatio(X, &nxEff) && !Y.identity && Y.N <= 16 &&
            Y.D <= 65535 * 8 && ((uintptr_t)src % 4) == 0 && srcSt % 4 == 0 && srcFrameStride % 4 == 0 &&
            ((uintptr_t)dst % 4) == 0 && dstSt % 4 == 0 && dstFrameStride % 4 == 0) {
            r->lastKernel = "area_down";
            int nyEff = 1;   // vertical taps up to the last one that is non-zero in some row (3:2: the third tap is zero in both phases)
            for (int row = 0; row < Y.numRows; ++row)
                for (int k = 0; k < Y.N; ++k)
                    if (Y.coef[size_t(row) * Y.N + k] != 0) nyEff = std::max(nyEff, k + 1);
            for (size_t f0 = 0; f0 < nFrames; f0 += 65535) {
                const int nf = int(std::min<size_t>(65535, nFrames - f0));
                CUDA_TRY(launchAreaDown(int(X.rS), int(X.rD), X.N, nxEff, src + f0 * srcFrameStride, dst + f0 * dstFrameStride,
                                        (long long)srcSt, (long long)dstSt, (long long)srcFrameStride, (long long)dstFrameStride,
                                        int(X.S), int(Y.S), int(X.D), int(Y.D), nf, Y.N, nyEff, sp.ty.first, sp.ty.row, sp.ty.coef,
                                        &X.coef[0], stream));
            }
            return IQO_CUDA_OK;
        }
    }
    // Linear at 1:2, 1:3, 1:4, 2:3, 2:5, 3:4, 4:5 (up-sampling), 3:2 or 4:3 on X (any ratio on Y): streaming kernel
    {
        const AxisPlan &X = r->plan.x, &Y = r->plan.y;
        if (r->path == IQO_CUDA_PATH_AUTO && r->plan.kind == kLinear && whole && linearUpRatio(X) && !Y.identity && Y.N == 2 &&
            Y.D <= 65535 && ((uintptr_t)src % 4) == 0 && srcSt % 4 == 0 && srcFrameStride % 4 == 0 &&
            ((uintptr_t)dst % 4) == 0 && dstSt % 4 == 0 && dstFrameStride % 4 == 0) {
            int q1[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int t = 0; t < int(X.rD); ++t) q1[t] = X.coef[size_t(t) * 2 + 1];   // weight of the right column of phase t
            r->lastKernel = linearUpName(X);
            for (size_t f0 = 0; f0 < nFrames; f0 += 65535) {
                const int nf = int(std::min<size_t>(65535, nFrames - f0));
                CUDA_TRY(launchLinearUp(int(X.rS), int(X.rD), src + f0 * srcFrameStride, dst + f0 * dstFrameStride, (long long)srcSt,
                                        (long long)dstSt, (long long)srcFrameStride, (long long)dstFrameStride, int(X.S), int(Y.S),
                                        int(X.D), int(Y.D), nf, sp.ty.first, sp.ty.row, sp.ty.coef, q1, stream));
            }
            return IQO_CUDA_OK;
        }
    }
    // Reductions with ten or more taps per axis pattern (Lanczos3 / 4 at 3:2, 2:1 on X only ...) are faster on the tensor-path
    // kernel once the launch fills the device (measured in one run on B200, 512 frames 1080p -> 720p: Lanczos3 0.61 vs
    // 0.67 ms, Lanczos4 0.69 vs 0.83 ms, 1080p -> 960x720 0.56 vs 0.72 ms; Lanczos2 ties, up-sampling is 8 % slower there)
    if (!r->forceStream && !r->forceMma && sp.ratio.eligible && sp.ratio.RS > sp.ratio.RD && sp.ratio.NX >= 10 && mmaWorthIt(r, nFrames, dstRows)) {
        const int rc = tryMma();
        if (rc != 0) return rc < 0 ? rc : IQO_CUDA_OK;
    }
    // Lanczos at 3:2, 1:2, 3:4 ...: rational-ratio streaming kernel
    // (cross-over against the packed kernel: about 600 warps at the shortest band, three 1080p -> 720p frames)
    if (r->useStream && sp.ratio.eligible && whole && ((uintptr_t)src % 8) == 0 && srcSt % 8 == 0 && srcFrameStride % 8 == 0 &&
        (r->forceStream || (long long)((r->plan.x.D + 8 * sp.ratio.groupsPerStrip - 1) / (8 * sp.ratio.groupsPerStrip)) *
                                   ((r->plan.y.D + 31) / 32) * (long long)nFrames >= 4ll * sp.sms)) {
        const RatioPlan &rp = sp.ratio;
        RatioArgs q;
        q.srcPitch = (long long)srcSt;
        q.dstPitch = (long long)dstSt;
        q.srcFrameStride = (long long)srcFrameStride;
        q.dstFrameStride = (long long)dstFrameStride;
        q.SW = int(r->plan.x.S);
        q.SH = int(r->plan.y.S);
        q.DW = int(r->plan.x.D);
        q.DH = int(r->plan.y.D);
        q.RS = rp.RS;
        q.RD = rp.RD;
        q.NX = rp.NX;
        q.tailZeros = rp.tailZeros;
        q.groupsPerStrip = rp.groupsPerStrip;
        q.c0 = rp.c0;
        q.workBias = rp.workBias;
        q.accInit = rp.accInit;
        q.dstVec = ((uintptr_t)dst % 8) == 0 && dstSt % 8 == 0 && dstFrameStride % 8 == 0;
        q.rowRec = sp.rRowRec;
        memset(q.cwX, 0, sizeof q.cwX);
        for (int ph = 0; ph < rp.RD && ph < 4; ++ph)
            for (int par = 0; par < 2; ++par)
                for (int j = 0; j < 7; ++j) q.cwX[ph][par][j] = rp.cwX[(size_t(ph) * 2 + par) * 7 + j];
        q.odd = rp.odd;
        q.mbX = int(r->plan.x.mainBegin);
        q.meX = int(r->plan.x.mainEnd);
        q.gx = a.x;
        q.gy = a.y;
        // bands of whole 8-row turns: enough warps to fill the device several times over
        const long long strips = (q.DW + 8 * q.groupsPerStrip - 1) / (8 * q.groupsPerStrip);
        int bandRows = 256;
        while (bandRows > 32 && strips * ((q.DH + bandRows - 1) / bandRows) * (long long)nFrames < 6ll * sp.sms * 16) bandRows /= 2;
        q.bandRows = bandRows;
        if ((q.DH + bandRows - 1) / bandRows <= 65535) {
            r->lastKernel = "ratio_stream";
            for (size_t f0 = 0; f0 < nFrames; f0 += 65535) {
                q.nFrames = int(std::min<size_t>(65535, nFrames - f0));
                q.src = src + f0 * srcFrameStride;
                q.dst = dst + f0 * dstFrameStride;
                CUDA_TRY(launchRatio(q, stream));
            }
            return IQO_CUDA_OK;
        }
    }
    // Lanczos at any other ratio, row bands included: the tensor-path kernel when the launch is big enough
    // (measured on B200, cfg5's ratio: 0.72 ms against 1.13 ms for the general streaming kernel below)
    if (!r->forceMma && mmaWorthIt(r, nFrames, dstRows)) {
        const int rc = tryMma();
        if (rc != 0) return rc < 0 ? rc : IQO_CUDA_OK;
    }
    // ... else the general streaming kernel
    // (cross-over against the packed kernel, warm back-to-back launches of cfg5's row bands, tools/gigapixel.py --bands:
    //  42 warps per SM at the shortest band 0.154 vs 0.280 ms, 11 per SM 48 vs 61 us, 5.4 per SM 35 vs 31 us)
    if (r->useStream && sp.lstream.eligible && ((uintptr_t)src % 8) == 0 && srcSt % 8 == 0 && srcFrameStride % 8 == 0 &&
        (r->forceStream || (long long)((r->plan.x.D + sp.lstream.stripW - 1) / sp.lstream.stripW) * ((dstRows + 31) / 32) *
                                   (long long)nFrames >= 12ll * sp.sms)) {
        LStreamArgs q;
        q.srcPitch = (long long)srcSt;
        q.dstPitch = (long long)dstSt;
        q.srcFrameStride = (long long)srcFrameStride;
        q.dstFrameStride = (long long)dstFrameStride;
        q.SW = int(r->plan.x.S);
        q.DW = int(r->plan.x.D);
        q.srcRow0 = int(srcRow0);
        q.srcRows = int(srcRows);
        q.dstRow0 = int(dstRow0);
        q.dstRows = int(dstRows);
        q.stripW = sp.lstream.stripW;
        q.workBias = sp.packed.workBias;
        q.rowRec = sp.gRowRec;
        q.recX = reinterpret_cast<const int4 *>(sp.pRecX);
        q.cwX = sp.pCwX;
        q.NP = sp.packed.NP;
        const long long strips = (q.DW + q.stripW - 1) / q.stripW;
        int bandRows = 256;
        while (bandRows > 32 && strips * ((q.dstRows + bandRows - 1) / bandRows) * (long long)nFrames < 6ll * sp.sms * 16) bandRows /= 2;
        q.bandRows = bandRows;
        if ((q.dstRows + bandRows - 1) / bandRows <= 65535) {
            r->lastKernel = "lanczos_stream";
            for (size_t f0 = 0; f0 < nFrames; f0 += 65535) {
                q.nFrames = int(std::min<size_t>(65535, nFrames - f0));
                q.src = src + f0 * srcFrameStride;
                q.dst = dst + f0 * dstFrameStride;
                CUDA_TRY(launchLStream(q, stream));
            }
            return IQO_CUDA_OK;
        }
    }
    if (r->path == IQO_CUDA_PATH_AUTO && sp.packed.eligible && ((uintptr_t)src % 4) == 0 && srcSt % 4 == 0 &&
        srcFrameStride % 4 == 0 && dstRows <= size_t(65535) * sp.pgeom.tileH) {
        PackedArgs q;
        q.srcPitch = (long long)srcSt;
        q.dstPitch = (long long)dstSt;
        q.srcFrameStride = (long long)srcFrameStride;
        q.dstFrameStride = (long long)dstFrameStride;
        q.SW = int(r->plan.x.S);
        q.SH = int(r->plan.y.S);
        q.DW = int(r->plan.x.D);
        q.DH = int(r->plan.y.D);
        q.srcRow0 = int(srcRow0);
        q.dstRow0 = int(dstRow0);
        q.dstRows = int(dstRows);
        q.tileW = sp.pgeom.tileW;
        q.tileH = sp.pgeom.tileH;
        q.wordsPerRow = sp.pgeom.wordsPerRow;
        q.shift = r->plan.shift;
        q.isSigned = r->plan.workSigned;
        q.workBias = sp.packed.workBias;
        q.firstY = sp.pFirstY;
        q.ntapY = sp.pNtapY;
        q.coefOffY = sp.pCoefOffY;
        q.coefY = sp.ty.coef;
        q.rowY = sp.ty.row;
        q.denoY = sp.ty.deno;
        q.magicY = sp.pMagicY;
        q.ntMax = sp.packed.ntMax;
        q.recX = reinterpret_cast<const int4 *>(sp.pRecX);
        q.cwX = sp.pCwX;
        q.NX = r->plan.x.N;
        q.NP = sp.packed.NP;
        r->lastKernel = "packed";
        for (size_t f0 = 0; f0 < nFrames; f0 += 65535) {
            q.nFrames = int(std::min<size_t>(65535, nFrames - f0));
            q.src = src + f0 * srcFrameStride;
            q.dst = dst + f0 * dstFrameStride;
            CUDA_TRY(launchPacked(q, stream));
        }
        return IQO_CUDA_OK;
    }
    r->lastKernel = "generic";
    CUDA_TRY(launchGeneric(a, r->geom, stream));
    return IQO_CUDA_OK;
}

int checkStrides(const iqo_cuda_resizer *r, size_t srcSt, const void *src, size_t dstSt, const void *dst)
{
    if (!r) return fail(IQO_CUDA_E_ARG, "NULL resizer");
    if (!src || !dst) return fail(IQO_CUDA_E_ARG, "NULL image pointer");
    if (srcSt < size_t(r->plan.x.S)) return fail(IQO_CUDA_E_ARG, "srcSt (%zu) < srcW (%lld)", srcSt, (long long)r->plan.x.S);
    if (dstSt < size_t(r->plan.x.D)) return fail(IQO_CUDA_E_ARG, "dstSt (%zu) < dstW (%lld)", dstSt, (long long)r->plan.x.D);
    return IQO_CUDA_OK;
}

// make sure the two staging slots of the handle's workspace hold `frames` frames each
int ensureStaging(iqo_cuda_resizer *r, size_t frames)
{
    r->srcPitch = alignUp(size_t(r->plan.x.S), 16);
    r->dstPitch = alignUp(size_t(r->plan.x.D), 16);
    const size_t sBytes = r->srcPitch * size_t(r->plan.y.S) * frames;
    const size_t dBytes = r->dstPitch * size_t(r->plan.y.D) * frames;
    Workspace *w = r->ws;
    if (w->srcCap < sBytes) {
        for (int i = 0; i < 2; ++i) {
            cudaFree(w->dSrc[i]);
            w->dSrc[i] = 0;
        }
        w->srcCap = 0;
        for (int i = 0; i < 2; ++i)
            if (cudaMalloc(&w->dSrc[i], sBytes) != cudaSuccess) {
                cudaGetLastError();
                return fail(IQO_CUDA_E_NOMEM, "cannot allocate %zu bytes of device staging", sBytes);
            }
        w->srcCap = sBytes;
    }
    if (w->dstCap < dBytes) {
        for (int i = 0; i < 2; ++i) {
            cudaFree(w->dDst[i]);
            w->dDst[i] = 0;
        }
        w->dstCap = 0;
        for (int i = 0; i < 2; ++i)
            if (cudaMalloc(&w->dDst[i], dBytes) != cudaSuccess) {
                cudaGetLastError();
                return fail(IQO_CUDA_E_NOMEM, "cannot allocate %zu bytes of device staging", dBytes);
            }
        w->dstCap = dBytes;
    }
    r->slotFrames = frames;
    return IQO_CUDA_OK;
}

// true when p is ordinary pageable host memory (not pinned / registered, not device)
bool isPageable(const void *p)
{
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return attr.type == cudaMemoryTypeUnregistered;
}

// rows x widthBytes from src to dst (different pitches allowed), split over a few host threads: one thread moves
// ~10 GB/s, the PCIe link wants five times that
void parallelCopy2D(uint8_t *dst, size_t dstPitch, const uint8_t *src, size_t srcPitch, size_t widthBytes, size_t rows)
{
    static const int envThreads = [] { const char *e = getenv("IQO_CUDA_COPY_THREADS"); return e ? atoi(e) : 0; }();
    const size_t total = widthBytes * rows;
    int nt = envThreads > 0 ? envThreads : int(std::min<size_t>(8, std::max<unsigned>(1u, std::thread::hardware_concurrency() / 2)));
    if (total < (size_t(4) << 20)) nt = 1;
    auto part = [=](size_t r0, size_t r1) {
        if (dstPitch == widthBytes && srcPitch == widthBytes) {
            memcpy(dst + r0 * widthBytes, src + r0 * widthBytes, (r1 - r0) * widthBytes);
        } else {
            for (size_t r = r0; r < r1; ++r) memcpy(dst + r * dstPitch, src + r * srcPitch, widthBytes);
        }
    };
    if (nt <= 1) {
        part(0, rows);
        return;
    }
    std::vector<std::thread> th;
    const size_t per = (rows + nt - 1) / nt;
    for (int i = 1; i < nt; ++i) {
        const size_t r0 = std::min(rows, per * i), r1 = std::min(rows, r0 + per);
        if (r1 > r0) th.push_back(std::thread(part, r0, r1));
    }
    part(0, std::min(rows, per));
    for (size_t i = 0; i < th.size(); ++i) th[i].join();
}

// pinned host staging of the handle's workspace: `sBytes` / `dBytes` per slot (0: not needed)
int ensureHostStaging(iqo_cuda_resizer *r, size_t sBytes, size_t dBytes)
{
    Workspace *w = r->ws;
    for (int i = 0; i < 2; ++i)
        if (!w->slotDone[i] && cudaEventCreateWithFlags(&w->slotDone[i], cudaEventDisableTiming) != cudaSuccess)
            return fail(IQO_CUDA_E_CUDA, "event creation failed: %s", cudaGetErrorString(cudaGetLastError()));
    if (w->hSrcCap < sBytes) {
        for (int i = 0; i < 2; ++i) {
            cudaFreeHost(w->hSrc[i]);
            w->hSrc[i] = 0;
        }
        w->hSrcCap = 0;
        for (int i = 0; i < 2; ++i)
            if (cudaHostAlloc(reinterpret_cast<void **>(&w->hSrc[i]), sBytes, cudaHostAllocPortable) != cudaSuccess) {
                cudaGetLastError();
                return fail(IQO_CUDA_E_NOMEM, "cannot allocate %zu bytes of pinned host staging", sBytes);
            }
        w->hSrcCap = sBytes;
    }
    if (w->hDstCap < dBytes) {
        for (int i = 0; i < 2; ++i) {
            cudaFreeHost(w->hDst[i]);
            w->hDst[i] = 0;
        }
        w->hDstCap = 0;
        for (int i = 0; i < 2; ++i)
            if (cudaHostAlloc(reinterpret_cast<void **>(&w->hDst[i]), dBytes, cudaHostAllocPortable) != cudaSuccess) {
                cudaGetLastError();
                return fail(IQO_CUDA_E_NOMEM, "cannot allocate %zu bytes of pinned host staging", dBytes);
            }
        w->hDstCap = dBytes;
    }
    return IQO_CUDA_OK;
}

// ---- plan cache and workspace pool (process wide, never destroyed: no static-destruction
// order problems with the CUDA runtime) ----
struct PlanKey {
    int device, kind;
    unsigned degree;
    size_t sw, sh, dw, dh, px;
    bool operator<(const PlanKey &o) const
    {
        return std::tie(device, kind, degree, sw, sh, dw, dh, px) < std::tie(o.device, o.kind, o.degree, o.sw, o.sh, o.dw, o.dh, o.px);
    }
};

struct Registry {
    std::mutex mu;
    std::map<PlanKey, std::shared_ptr<SharedPlan> > plans;
    std::vector<PlanKey> order;  // insertion order, for eviction
    std::vector<Workspace *> freeWs;
};

Registry &registry()
{
    static Registry *g = new Registry();
    return *g;
}

const size_t kMaxCachedPlans = 32, kMaxPooledWorkspaces = 16;

int buildSharedPlan(std::shared_ptr<SharedPlan> &out, int device, int kind, unsigned degree, size_t srcW, size_t srcH,
                    size_t dstW, size_t dstH, size_t pxScale)
{
    std::shared_ptr<SharedPlan> sp(new SharedPlan());
    sp->device = device;
    sp->sms = 148;
    if (cudaDeviceGetAttribute(&sp->sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || sp->sms <= 0) {
        cudaGetLastError();
        sp->sms = 148;
    }
    int rc = buildPlan(sp->plan, kind, degree, srcW, srcH, dstW, dstH, pxScale);
    if (rc != kPlanOk) return fail(rc, "%s", sp->plan.error.c_str());  // PlanError values equal the IQO_CUDA_E_* codes
    {
        cudaError_t e = initKernels();
        if (e != cudaSuccess) return fail(IQO_CUDA_E_CUDA, "kernel setup failed: %s", cudaGetErrorString(e));
    }
    rc = uploadAxis(sp->plan.x, sp->tx);
    if (rc == IQO_CUDA_OK) rc = uploadAxis(sp->plan.y, sp->ty);
    if (rc != IQO_CUDA_OK) return rc;
    buildHalfPlan(sp->plan, sp->half);
    if (sp->half.eligible) {
        const HalfPlan &hp = sp->half;
        const size_t b0 = hp.borderY.size() * 4, b1 = hp.magicY.size() * 4, b2 = std::max<size_t>(4, hp.borderX.size() * 4);
        if (cudaMalloc(&sp->dBorderY, b0) != cudaSuccess || cudaMalloc(&sp->dMagicY, b1) != cudaSuccess ||
            cudaMalloc(&sp->dBorderX, b2) != cudaSuccess ||
            cudaMemcpy(sp->dBorderY, hp.borderY.data(), b0, cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(sp->dMagicY, hp.magicY.data(), b1, cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(sp->dBorderX, hp.borderX.data(), hp.borderX.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
            cudaGetLastError();
            sp->half.eligible = false;
        }
        if (sp->half.eligible && sp->half.sEligible &&
            (!halfStreamHasKernel(hp.sNG, hp.NXH) || !uploadVec(sp->dSBorderY, hp.sBorderY) || !uploadVec(sp->dBorderXo, hp.borderXo))) {
            cudaGetLastError();
            sp->half.sEligible = false;
        }
    }
    sp->geom = chooseGenericGeom(sp->plan.x.first.data(), sp->plan.x.N, int(sp->plan.x.S), int(sp->plan.x.D));
    if (sp->geom.smemBytes > 200 * 1024)  // the last-resort kernel cannot hold one row of the source window
        return fail(IQO_CUDA_E_TOO_LARGE, "horizontal source window of %d columns per 8 destination pixels exceeds the kernels' shared memory",
                    sp->geom.workW);
    buildSmallPlan(sp->plan, sp->small);
    if (sp->small.eligible && (!uploadVec(sp->sRowsY, sp->small.rowsY) || !uploadVec(sp->sMagicY, sp->small.magicY))) {
        cudaGetLastError();
        sp->small.eligible = false;
    }
    buildRatioPlan(sp->plan, sp->ratio);
    if (sp->ratio.eligible && (!ratioHasKernel(sp->ratio.RS, sp->ratio.RD, sp->ratio.NX, sp->ratio.odd) || !uploadVec(sp->rRowRec, sp->ratio.rowRec))) {
        cudaGetLastError();
        sp->ratio.eligible = false;
    }
    buildPackedPlan(sp->plan, sp->packed, packedPadNP(sp->plan.x.N / 2 + 1));
    if (sp->packed.eligible) {
        const PackedPlan &q = sp->packed;
        if (!uploadVec(sp->pFirstY, q.firstY) || !uploadVec(sp->pNtapY, q.ntapY) || !uploadVec(sp->pCoefOffY, q.coefOffY) ||
            !uploadVec(sp->pMagicY, q.magicY) || !uploadVec(sp->pCwX, q.cwX) || !uploadVec(sp->pRecX, q.recX)) {
            cudaGetLastError();
            sp->packed.eligible = false;
        } else {
            sp->pgeom = choosePackedGeom(q.firstX.data(), sp->plan.x.N, int(sp->plan.x.S), int(sp->plan.x.D), q.NP, q.ntMax);
            if (sp->pgeom.smemBytes > 150 * 1024) sp->packed.eligible = false;  // extreme down-sampling: generic kernel
        }
    }
    sp->lstream.eligible = false;
    if (sp->packed.eligible) {
        buildLStreamPlan(sp->plan, sp->packed, sp->lstream);
        if (sp->lstream.eligible && (!lstreamHasKernel(sp->packed.NP) || !uploadVec(sp->gRowRec, sp->lstream.rowRec))) {
            cudaGetLastError();
            sp->lstream.eligible = false;
        }
    }
    static const int envMmaWcols = [] { const char *e = getenv("IQO_CUDA_MMA_WCOLS"); return e ? atoi(e) : 0; }();
    // strips of 272 source columns for the all-mma kernel; the variant with the dp2a horizontal pass keeps no tables and
    // no output tile in shared memory and is 4 % faster with 400 (measured 272 / 336 / 400 / 448 / 512: 2.22 / 2.14 / 2.14 / 2.13 / 2.76 ms on cfg1)
    const bool dp2aVariant = sp->plan.kind == kLanczos && sp->ratio.eligible &&
                             mmaRatioHasKernel(sp->ratio.RS, sp->ratio.RD, sp->ratio.NX, sp->ratio.odd) && (sp->ratio.c0 & 1) == sp->ratio.odd &&
                             sp->plan.x.D % 8 == 0;
    buildMmaPlan(sp->plan, sp->mma, envMmaWcols > 0 ? envMmaWcols : dp2aVariant ? 400 : IQO_MMA_WCOLS_DEFAULT);
    if (sp->mma.eligible) {
        const MmaPlan &q = sp->mma;
        if (encodeTiled() == 0 || !uploadVec(sp->mVBlock, q.vBlock) || !uploadVec(sp->mVRow, q.vRow) || !uploadVec(sp->mVRowMap, q.vRowMap) || !uploadVec(sp->mStripXs, q.stripXs) ||
            !uploadVec(sp->mHTile, q.hTile) || !uploadVec(sp->mHCol, q.hCol) || !uploadVec(sp->mVFrag, q.vFrag) || !uploadVec(sp->mHFrag, q.hFrag)) {
            cudaGetLastError();
            sp->mma.eligible = false;
        }
    }
    out = sp;
    return IQO_CUDA_OK;
}

int acquireWorkspace(Workspace *&out, int device)
{
    Registry &g = registry();
    {
        std::lock_guard<std::mutex> lock(g.mu);
        for (size_t i = 0; i < g.freeWs.size(); ++i)
            if (g.freeWs[i]->device == device) {
                out = g.freeWs[i];
                g.freeWs.erase(g.freeWs.begin() + i);
                return IQO_CUDA_OK;
            }
    }
    Workspace *w = new Workspace();
    w->device = device;
    for (int i = 0; i < 2; ++i)
        if (cudaStreamCreateWithFlags(&w->stream[i], cudaStreamNonBlocking) != cudaSuccess) {
            const char *msg = cudaGetErrorString(cudaGetLastError());
            for (int j = 0; j < i; ++j) cudaStreamDestroy(w->stream[j]);
            delete w;
            return fail(IQO_CUDA_E_CUDA, "stream creation failed: %s", msg);
        }
    out = w;
    return IQO_CUDA_OK;
}

void destroyWorkspace(Workspace *w)
{
    for (int i = 0; i < 2; ++i) {
        if (w->stream[i]) cudaStreamDestroy(w->stream[i]);
        if (w->slotDone[i]) cudaEventDestroy(w->slotDone[i]);
        cudaFree(w->dSrc[i]);
        cudaFree(w->dDst[i]);
        cudaFreeHost(w->hSrc[i]);
        cudaFreeHost(w->hDst[i]);
    }
    cudaGetLastError();
    delete w;
}

}  // namespace

// ------------------------------------------------------------------------------------------

extern "C" {

int iqo_cuda_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

const char *iqo_cuda_version(void)
{
    return "libiqo_b200 0.1 (sm_100a)";
}

const char *iqo_cuda_last_error(void)
{
    return t_lastError.c_str();
}

unsigned long long iqo_cuda_launch_count(void)
{
    return launchCount();
}

void *iqo_cuda_host_alloc(size_t bytes)
{
    void *p = 0;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        fail(IQO_CUDA_E_NOMEM, "cudaHostAlloc(%zu) failed", bytes);
        return 0;
    }
    return p;
}

void iqo_cuda_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}

int iqo_cuda_create_on(iqo_cuda_resizer **out, int device, int kind, unsigned degree,
                       size_t srcW, size_t srcH, size_t dstW, size_t dstH, size_t pxScale)
{
    IQO_GUARD_BEGIN
    if (!out) return fail(IQO_CUDA_E_ARG, "NULL output handle");
    *out = 0;
    if (kind != IQO_CUDA_LANCZOS) {
        degree = 0;  // ignored by Area / Linear: do not let it split the plan cache
        pxScale = 1;
    }
    {
        // argument errors first, so that they are reported as such even without a device
        Plan probe;
        if (kind < 0 || kind > 2 || !srcW || !srcH || !dstW || !dstH || (kind == IQO_CUDA_LANCZOS && (!degree || !pxScale))) {
            int rc = buildPlan(probe, kind, degree, srcW, srcH, dstW, dstH, pxScale);
            return fail(rc, "%s", probe.error.c_str());
        }
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        Plan probe;
        int rc = buildPlan(probe, kind, degree, srcW, srcH, dstW, dstH, pxScale);
        if (rc != kPlanOk) return fail(rc, "%s", probe.error.c_str());
        return fail(IQO_CUDA_E_CUDA, "no CUDA device available (this library has no CPU fallback)");
    }
    if (device < 0 || device >= ndev) return fail(IQO_CUDA_E_ARG, "device %d out of range (0..%d)", device, ndev - 1);
    DeviceGuard guard(device);
    if (!guard.ok) return fail(IQO_CUDA_E_CUDA, "cannot select device %d", device);

    Registry &g = registry();
    const PlanKey key = {device, kind, degree, srcW, srcH, dstW, dstH, pxScale};
    std::shared_ptr<SharedPlan> sp;
    {
        std::lock_guard<std::mutex> lock(g.mu);
        std::map<PlanKey, std::shared_ptr<SharedPlan> >::iterator it = g.plans.find(key);
        if (it != g.plans.end()) sp = it->second;
    }
    if (!sp) {
        int rc = buildSharedPlan(sp, device, kind, degree, srcW, srcH, dstW, dstH, pxScale);
        if (rc != IQO_CUDA_OK) return rc;
        std::lock_guard<std::mutex> lock(g.mu);
        if (g.plans.find(key) == g.plans.end()) {
            // evict the oldest plans that no handle uses any more
            for (size_t i = 0; g.plans.size() >= kMaxCachedPlans && i < g.order.size();) {
                std::map<PlanKey, std::shared_ptr<SharedPlan> >::iterator old = g.plans.find(g.order[i]);
                if (old != g.plans.end() && old->second.use_count() == 1) {
                    g.plans.erase(old);
                    g.order.erase(g.order.begin() + i);
                } else {
                    ++i;
                }
            }
            g.plans[key] = sp;
            g.order.push_back(key);
        } else {
            sp = g.plans[key];  // another thread built it meanwhile
        }
    }
    Workspace *ws = 0;
    int rc = acquireWorkspace(ws, device);
    if (rc != IQO_CUDA_OK) return rc;
    *out = new iqo_cuda_resizer(sp, ws);
    return IQO_CUDA_OK;
    IQO_GUARD_END
}

int iqo_cuda_create(iqo_cuda_resizer **out, int kind, unsigned degree,
                    size_t srcW, size_t srcH, size_t dstW, size_t dstH, size_t pxScale)
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        cudaGetLastError();
        dev = 0;  // create_on reports the missing device (after argument errors)
    }
    return iqo_cuda_create_on(out, dev, kind, degree, srcW, srcH, dstW, dstH, pxScale);
}

void iqo_cuda_destroy(iqo_cuda_resizer *r)
{
    if (!r) return;
    DeviceGuard guard(r->device);
    Workspace *w = r->ws;
    for (int i = 0; i < 2; ++i) cudaStreamSynchronize(w->stream[i]);
    cudaGetLastError();
    Registry &g = registry();
    bool pooled = false;
    {
        std::lock_guard<std::mutex> lock(g.mu);
        // huge staging buffers are not worth keeping around
        if (g.freeWs.size() < kMaxPooledWorkspaces && w->srcCap + w->dstCap + w->hSrcCap + w->hDstCap <= (size_t(512) << 20)) {
            g.freeWs.push_back(w);
            pooled = true;
        }
    }
    if (!pooled) destroyWorkspace(w);
    delete r;
}

// Drop every cached plan and pooled workspace (tests; before cudaDeviceReset).
void iqo_cuda_clear_cache(void)
{
    Registry &g = registry();
    std::vector<Workspace *> ws;
    {
        std::lock_guard<std::mutex> lock(g.mu);
        ws.swap(g.freeWs);
        g.plans.clear();
        g.order.clear();
    }
    for (size_t i = 0; i < ws.size(); ++i) {
        DeviceGuard guard(ws[i]->device);
        destroyWorkspace(ws[i]);
    }
}

int iqo_cuda_sync(iqo_cuda_resizer *r)
{
    if (!r) return fail(IQO_CUDA_E_ARG, "NULL resizer");
    DeviceGuard guard(r->device);
    CUDA_TRY(cudaStreamSynchronize(r->stream[0]));
    CUDA_TRY(cudaStreamSynchronize(r->stream[1]));
    return IQO_CUDA_OK;
}

int iqo_cuda_set_path(iqo_cuda_resizer *r, int path)
{
    if (!r || path < 0 || path > IQO_CUDA_PATH_NO_MMA) return fail(IQO_CUDA_E_ARG, "bad path");
    r->useTma = (path != IQO_CUDA_PATH_NO_TMA);
    r->useStream = (path == IQO_CUDA_PATH_AUTO || path == IQO_CUDA_PATH_STREAM || path == IQO_CUDA_PATH_NO_MMA);
    r->forceStream = (path == IQO_CUDA_PATH_STREAM);
    r->useMma = (path == IQO_CUDA_PATH_AUTO || path == IQO_CUDA_PATH_MMA);
    r->forceMma = (path == IQO_CUDA_PATH_MMA);
    r->path = (path == IQO_CUDA_PATH_GENERIC) ? IQO_CUDA_PATH_GENERIC : IQO_CUDA_PATH_AUTO;
    return IQO_CUDA_OK;
}

int iqo_cuda_set_arithmetic(iqo_cuda_resizer *r, int arithmetic)
{
    IQO_GUARD_BEGIN
    if (!r || (arithmetic != IQO_CUDA_ARITH_FIXED && arithmetic != IQO_CUDA_ARITH_SIMD_FLOAT)) return fail(IQO_CUDA_E_ARG, "bad arithmetic mode");
    if (arithmetic == IQO_CUDA_ARITH_SIMD_FLOAT) {
        SharedPlan &sp = *r->sp;
        std::lock_guard<std::mutex> lock(sp.fltMu);
        if (!sp.fltBuilt) {
            buildFloatPlan(sp.plan, sp.flt);
            if (!sp.flt.eligible) return fail(IQO_CUDA_E_UNSUPPORTED, "float mode: %s", sp.flt.why.c_str());
            DeviceGuard guard(r->device);
            if (!uploadVec(sp.fCoefX, sp.flt.x.coef) || !uploadVec(sp.fCoefY, sp.flt.y.coef) || !uploadVec(sp.fDenoX, sp.flt.x.deno) ||
                !uploadVec(sp.fDenoY, sp.flt.y.deno)) {
                cudaGetLastError();
                return fail(IQO_CUDA_E_NOMEM, "float mode: table upload failed");
            }
            sp.fltBuilt = true;
        }
    }
    r->arithmetic = arithmetic;
    return IQO_CUDA_OK;
    IQO_GUARD_END
}

const char *iqo_cuda_last_kernel(const iqo_cuda_resizer *r)
{
    return r ? r->lastKernel : "none";
}

int iqo_cuda_get_table(const iqo_cuda_resizer *r, int axis, int *numCoefs, int *numTables, int32_t *out, size_t cap)
{
    IQO_GUARD_BEGIN
    if (!r || axis < 0 || axis > 1) return fail(IQO_CUDA_E_ARG, "bad argument");
    const AxisPlan &a = axis ? r->plan.y : r->plan.x;
    if (numCoefs) *numCoefs = a.N;
    if (numTables) *numTables = a.identity ? 1 : int(a.rD);
    if (out) {
        size_t n = size_t(a.N) * size_t(a.identity ? 1 : a.rD);
        for (size_t i = 0; i < n && i < cap; ++i) out[i] = a.coef[i];
    }
    return IQO_CUDA_OK;
    IQO_GUARD_END
}

int iqo_cuda_plan_query(int kind, unsigned degree, size_t srcW, size_t srcH, size_t dstW, size_t dstH, size_t pxScale,
                        int axis, int *numCoefs, int *numTables, int *numRows, long long *mainBegin, long long *mainEnd,
                        int32_t *coefs, size_t coefCap, int32_t *first, int32_t *row, size_t indexCap)
{
    IQO_GUARD_BEGIN
    if (axis < 0 || axis > 1) return fail(IQO_CUDA_E_ARG, "bad axis");
    Plan p;
    int rc = buildPlan(p, kind, degree, srcW, srcH, dstW, dstH, pxScale);
    if (rc != kPlanOk) return fail(rc, "%s", p.error.c_str());
    const AxisPlan &a = axis ? p.y : p.x;
    if (numCoefs) *numCoefs = a.N;
    if (numTables) *numTables = a.identity ? 1 : int(a.rD);
    if (numRows) *numRows = a.numRows;
    if (mainBegin) *mainBegin = a.mainBegin;
    if (mainEnd) *mainEnd = a.mainEnd;
    if (coefs)
        for (size_t i = 0; i < a.coef.size() && i < coefCap; ++i) coefs[i] = a.coef[i];
    for (size_t i = 0; i < a.first.size() && i < indexCap; ++i) {
        if (first) first[i] = a.first[i];
        if (row) row[i] = a.row[i];
    }
    return IQO_CUDA_OK;
    IQO_GUARD_END
}

int iqo_cuda_plan_kernel(int kind, unsigned degree, size_t srcW, size_t srcH, size_t dstW, size_t dstH, size_t pxScale,
                         char *kernel, size_t kernelCap, char *why, size_t whyCap)
{
    IQO_GUARD_BEGIN
    Plan p;
    int rc = buildPlan(p, kind, degree, srcW, srcH, dstW, dstH, pxScale);
    if (rc != kPlanOk) return fail(rc, "%s", p.error.c_str());
    HalfPlan h;
    buildHalfPlan(p, h);
    PackedPlan q;
    buildPackedPlan(p, q, packedPadNP(p.x.N / 2 + 1));
    SmallPlan sm;
    buildSmallPlan(p, sm);
    RatioPlan rt;
    buildRatioPlan(p, rt);
    const bool ratio = rt.eligible && ratioHasKernel(rt.RS, rt.RD, rt.NX, rt.odd);
    LStreamPlan ls;
    ls.eligible = false;
    if (q.eligible) buildLStreamPlan(p, q, ls);
    const bool lstream = ls.eligible && lstreamHasKernel(q.NP);
    const bool area2 = p.kind == kArea && p.x.rD == 1 && p.x.rS == 2 && p.y.rD == 1 && p.y.rS == 2 && p.x.N == 2 && p.y.N == 2 && p.x.S % 16 == 0;
    const bool areadown = p.kind == kArea && areaDownRatio(p.x, 0) && !p.y.identity;
    const long long kx = (p.x.D % p.x.S == 0) ? p.x.D / p.x.S : 0;
    const bool linup = p.kind == kLinear && linearUpRatio(p.x) && !p.y.identity;
    if (kernel && kernelCap)
        snprintf(kernel, kernelCap, "%s", sm.eligible ? "half_small" : h.eligible ? (h.symmetric ? "half_sym" : "half") : area2 ? "area2"
                                          : areadown ? "area_down"
                                          : linup ? linearUpName(p.x) : ratio ? "ratio_stream" : lstream ? "lanczos_stream" : q.eligible ? "packed" : "generic");
    if (why && whyCap)
        snprintf(why, whyCap, "%s%s%s%s%s", h.why.c_str(), rt.eligible ? "" : "; ratio: ", rt.eligible ? "" : rt.why.c_str(),
                 q.eligible ? "" : "; packed: ", q.eligible ? "" : q.why.c_str());
    return IQO_CUDA_OK;
    IQO_GUARD_END
}

int iqo_cuda_resize_batch(iqo_cuda_resizer *r, size_t nFrames,
                          size_t srcSt, size_t srcFrameStride, const uint8_t *src,
                          size_t dstSt, size_t dstFrameStride, uint8_t *dst, void *stream)
{
    IQO_GUARD_BEGIN
    int rc = checkStrides(r, srcSt, src, dstSt, dst);
    if (rc) return rc;
    if (nFrames == 0) return IQO_CUDA_OK;
    if (nFrames > (size_t(1) << 30)) return fail(IQO_CUDA_E_TOO_LARGE, "too many frames");
    DeviceGuard guard(r->device);
    if (!isDevicePointer(src) || !isDevicePointer(dst))
        return fail(IQO_CUDA_E_ARG, "iqo_cuda_resize_batch needs device pointers (use iqo_cuda_resize_batch_host)");
    cudaStream_t s = (cudaStream_t)stream;  // NULL = the legacy default stream, as everywhere in CUDA
    return launch(r, nFrames, 0, size_t(r->plan.y.D), 0, size_t(r->plan.y.S),
                  srcSt, srcFrameStride, src, dstSt, dstFrameStride, dst, s);
    IQO_GUARD_END
}

int iqo_cuda_band_src_rows(const iqo_cuda_resizer *r, size_t dstRow0, size_t dstRows, size_t *srcRow0, size_t *srcRows)
{
    if (!r || !srcRow0 || !srcRows) return fail(IQO_CUDA_E_ARG, "NULL argument");
    const AxisPlan &y = r->plan.y;
    if (dstRows == 0 || dstRow0 + dstRows > size_t(y.D)) return fail(IQO_CUDA_E_ARG, "band outside the image");
    int64_t lo = y.S - 1, hi = 0;
    // first[] is non-decreasing; the band is spanned by its first and last row
    lo = std::min<int64_t>(std::max<int64_t>(y.first[dstRow0], 0), y.S - 1);
    hi = std::min<int64_t>(std::max<int64_t>(int64_t(y.first[dstRow0 + dstRows - 1]) + y.N - 1, 0), y.S - 1);
    *srcRow0 = size_t(lo);
    *srcRows = size_t(hi - lo + 1);
    return IQO_CUDA_OK;
}

int iqo_cuda_resize_band(iqo_cuda_resizer *r, size_t dstRow0, size_t dstRows, size_t srcRow0, size_t srcRows,
                         size_t srcSt, const uint8_t *src, size_t dstSt, uint8_t *dst, void *stream)
{
    IQO_GUARD_BEGIN
    int rc = checkStrides(r, srcSt, src, dstSt, dst);
    if (rc) return rc;
    size_t need0, needN;
    rc = iqo_cuda_band_src_rows(r, dstRow0, dstRows, &need0, &needN);
    if (rc) return rc;
    if (srcRow0 > need0 || srcRow0 + srcRows < need0 + needN)
        return fail(IQO_CUDA_E_ARG, "band needs source rows [%zu,%zu) but the buffer holds [%zu,%zu)",
                    need0, need0 + needN, srcRow0, srcRow0 + srcRows);
    DeviceGuard guard(r->device);
    if (!isDevicePointer(src) || !isDevicePointer(dst))
        return fail(IQO_CUDA_E_ARG, "iqo_cuda_resize_band needs device pointers");
    cudaStream_t s = (cudaStream_t)stream;  // NULL = the legacy default stream, as everywhere in CUDA
    return launch(r, 1, dstRow0, dstRows, srcRow0, srcRows, srcSt, 0, src, dstSt, 0, dst, s);
    IQO_GUARD_END
}

int iqo_cuda_resize_batch_host(iqo_cuda_resizer *r, size_t nFrames,
                               size_t srcSt, size_t srcFrameStride, const uint8_t *src,
                               size_t dstSt, size_t dstFrameStride, uint8_t *dst)
{
    IQO_GUARD_BEGIN
    int rc = checkStrides(r, srcSt, src, dstSt, dst);
    if (rc) return rc;
    if (nFrames == 0) return IQO_CUDA_OK;
    DeviceGuard guard(r->device);
    const size_t SW = size_t(r->plan.x.S), SH = size_t(r->plan.y.S);
    const size_t DW = size_t(r->plan.x.D), DH = size_t(r->plan.y.D);
    // chunk: about 64 MiB of source per slot, at least one frame
    const size_t frameBytes = alignUp(SW, 16) * SH + alignUp(DW, 16) * DH;
    size_t chunk = std::max<size_t>(1, (size_t(96) << 20) / frameBytes);
    chunk = std::min(chunk, nFrames);
    rc = ensureStaging(r, chunk);
    if (rc) return rc;
    const size_t sFrame = r->srcPitch * SH, dFrame = r->dstPitch * DH;
    // frames whose rows follow each other at the row stride can be moved by one 2-D copy per chunk
    const bool srcRegular = (srcFrameStride == srcSt * SH);
    const bool dstRegular = (dstFrameStride == dstSt * DH);

    // Pageable caller memory: the driver would stage such copies itself, synchronously and at ~10 GB/s.  Instead the
    // frames go through the handle's pinned staging slots: a few host threads fill / drain them while the copies and
    // the kernel of the other slot are in flight.
    const bool srcPg = isPageable(src), dstPg = isPageable(dst);
    if ((srcPg || dstPg) && nFrames * frameBytes >= (size_t(8) << 20)) {
        const size_t hsPitch = r->srcPitch, hdPitch = r->dstPitch;   // the pinned slots use the device layout: one flat copy each way
        rc = ensureHostStaging(r, srcPg ? chunk * sFrame : 0, dstPg ? chunk * dFrame : 0);
        if (rc) return rc;
        Workspace *w = r->ws;
        struct Pending {
            size_t first, n;
            bool live;
        } pend[2] = {{0, 0, false}, {0, 0, false}};
        auto retire = [&](int sl) -> int {   // wait for the slot's download, then hand the frames to the caller's buffer
            if (!pend[sl].live) return IQO_CUDA_OK;
            CUDA_TRY(cudaEventSynchronize(w->slotDone[sl]));
            if (dstPg) {
                if (dstRegular && dstSt == hdPitch && DW == hdPitch)
                    parallelCopy2D(dst + pend[sl].first * dstFrameStride, dstSt, w->hDst[sl], hdPitch, DW, DH * pend[sl].n);
                else
                    for (size_t f = 0; f < pend[sl].n; ++f)
                        parallelCopy2D(dst + (pend[sl].first + f) * dstFrameStride, dstSt, w->hDst[sl] + f * dFrame, hdPitch, DW, DH);
            }
            pend[sl].live = false;
            return IQO_CUDA_OK;
        };
        size_t donePg = 0;
        int sl = 0;
        while (donePg < nFrames) {
            const size_t n = std::min(chunk, nFrames - donePg);
            cudaStream_t st = r->stream[sl];
            rc = retire(sl);   // the slot's previous round is complete: its staging may be refilled
            if (rc) return rc;
            const uint8_t *hs = src + donePg * srcFrameStride;
            if (srcPg) {
                if (srcRegular && srcSt == hsPitch && SW == hsPitch)   // rows without padding: one flat copy
                    parallelCopy2D(w->hSrc[sl], hsPitch, hs, srcSt, hsPitch, SH * n);
                else
                    for (size_t f = 0; f < n; ++f) parallelCopy2D(w->hSrc[sl] + f * sFrame, hsPitch, hs + f * srcFrameStride, srcSt, SW, SH);
                CUDA_TRY(cudaMemcpyAsync(r->dSrc[sl], w->hSrc[sl], n * sFrame, cudaMemcpyHostToDevice, st));
            } else if (srcRegular) {
                CUDA_TRY(cudaMemcpy2DAsync(r->dSrc[sl], r->srcPitch, hs, srcSt, SW, SH * n, cudaMemcpyHostToDevice, st));
            } else {
                for (size_t f = 0; f < n; ++f)
                    CUDA_TRY(cudaMemcpy2DAsync(r->dSrc[sl] + f * sFrame, r->srcPitch, hs + f * srcFrameStride, srcSt, SW, SH, cudaMemcpyHostToDevice, st));
            }
            rc = launch(r, n, 0, DH, 0, SH, r->srcPitch, sFrame, r->dSrc[sl], r->dstPitch, dFrame, r->dDst[sl], st);
            if (rc) return rc;
            uint8_t *hd = dst + donePg * dstFrameStride;
            if (dstPg) {
                CUDA_TRY(cudaMemcpyAsync(w->hDst[sl], r->dDst[sl], n * dFrame, cudaMemcpyDeviceToHost, st));
            } else if (dstRegular) {
                CUDA_TRY(cudaMemcpy2DAsync(hd, dstSt, r->dDst[sl], r->dstPitch, DW, DH * n, cudaMemcpyDeviceToHost, st));
            } else {
                for (size_t f = 0; f < n; ++f)
                    CUDA_TRY(cudaMemcpy2DAsync(hd + f * dstFrameStride, dstSt, r->dDst[sl] + f * dFrame, r->dstPitch, DW, DH, cudaMemcpyDeviceToHost, st));
            }
            CUDA_TRY(cudaEventRecord(w->slotDone[sl], st));
            pend[sl].first = donePg;
            pend[sl].n = n;
            pend[sl].live = true;
            donePg += n;
            sl ^= 1;
        }
        rc = retire(sl);
        if (rc == IQO_CUDA_OK) rc = retire(sl ^ 1);
        if (rc) return rc;
        CUDA_TRY(cudaStreamSynchronize(r->stream[0]));
        CUDA_TRY(cudaStreamSynchronize(r->stream[1]));
        return IQO_CUDA_OK;
    }

    size_t done = 0;
    int slot = 0;
    while (done < nFrames) {
        const size_t n = std::min(chunk, nFrames - done);
        cudaStream_t s = r->stream[slot];
        // the slot's previous download must have finished before its buffers are reused:
        // same stream, so ordering is implicit.
        const uint8_t *hs = src + done * srcFrameStride;
        uint8_t *hd = dst + done * dstFrameStride;
        if (srcRegular) {
            CUDA_TRY(cudaMemcpy2DAsync(r->dSrc[slot], r->srcPitch, hs, srcSt, SW, SH * n, cudaMemcpyHostToDevice, s));
        } else {
            for (size_t f = 0; f < n; ++f)
                CUDA_TRY(cudaMemcpy2DAsync(r->dSrc[slot] + f * sFrame, r->srcPitch, hs + f * srcFrameStride, srcSt,
                                           SW, SH, cudaMemcpyHostToDevice, s));
        }
        rc = launch(r, n, 0, DH, 0, SH, r->srcPitch, sFrame, r->dSrc[slot], r->dstPitch, dFrame, r->dDst[slot], s);
        if (rc) return rc;
        if (dstRegular) {
            CUDA_TRY(cudaMemcpy2DAsync(hd, dstSt, r->dDst[slot], r->dstPitch, DW, DH * n, cudaMemcpyDeviceToHost, s));
        } else {
            for (size_t f = 0; f < n; ++f)
                CUDA_TRY(cudaMemcpy2DAsync(hd + f * dstFrameStride, dstSt, r->dDst[slot] + f * dFrame, r->dstPitch,
                                           DW, DH, cudaMemcpyDeviceToHost, s));
        }
        done += n;
        slot ^= 1;
    }
    CUDA_TRY(cudaStreamSynchronize(r->stream[0]));
    CUDA_TRY(cudaStreamSynchronize(r->stream[1]));
    return IQO_CUDA_OK;
    IQO_GUARD_END
}

int iqo_cuda_resize(iqo_cuda_resizer *r, size_t srcSt, const uint8_t *src, size_t dstSt, uint8_t *dst)
{
    IQO_GUARD_BEGIN
    int rc = checkStrides(r, srcSt, src, dstSt, dst);
    if (rc) return rc;
    DeviceGuard guard(r->device);
    const size_t SW = size_t(r->plan.x.S), SH = size_t(r->plan.y.S);
    const size_t DW = size_t(r->plan.x.D), DH = size_t(r->plan.y.D);
    const bool srcDev = isDevicePointer(src), dstDev = isDevicePointer(dst);
    // Stream ordering (see iqo_cuda.h): host-to-host calls run on the handle's private non-blocking stream.  As soon
    // as a device pointer is involved the work is enqueued on the legacy default stream instead, which is ordered
    // after everything already enqueued on the default stream and on every blocking stream -- so a source that
    // a previous kernel / copy is still producing (torch's current stream is the default stream unless changed)
    // is complete before it is read, and a destination still being read is not overwritten early.
    cudaStream_t s = (srcDev || dstDev) ? (cudaStream_t)0 : r->stream[0];
    const uint8_t *ksrc = src;
    uint8_t *kdst = dst;
    size_t kSrcSt = srcSt, kDstSt = dstSt;
    if (!srcDev || !dstDev) {
        rc = ensureStaging(r, 1);
        if (rc) return rc;
    }
    if (!srcDev) {
        CUDA_TRY(cudaMemcpy2DAsync(r->dSrc[0], r->srcPitch, src, srcSt, SW, SH, cudaMemcpyHostToDevice, s));
        ksrc = r->dSrc[0];
        kSrcSt = r->srcPitch;
    }
    if (!dstDev) {
        kdst = r->dDst[0];
        kDstSt = r->dstPitch;
    }
    rc = launch(r, 1, 0, DH, 0, SH, kSrcSt, 0, ksrc, kDstSt, 0, kdst, s);
    if (rc) return rc;
    if (!dstDev) CUDA_TRY(cudaMemcpy2DAsync(dst, dstSt, r->dDst[0], r->dstPitch, DW, DH, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return IQO_CUDA_OK;
    IQO_GUARD_END
}

// ---- planar YUV420 ----

}  // extern "C" (the struct below is C++)

struct iqo_cuda_yuv420 {
    iqo_cuda_resizer *luma, *chroma;
    size_t srcStX, srcStY, dstStX, dstStY;
    size_t srcSizeY, srcSizeU, srcSize, dstSizeY, dstSizeU, dstSize;
    uint8_t *dIn[2], *dOut[2];
    size_t slotFrames;
    // the chroma planes run on side streams forked from / joined to the caller's stream, so that
    // their CTAs fill the SMs the tail of the luma kernel leaves idle
    cudaStream_t side[2];
    cudaEvent_t fork, join[2];
};

namespace {

int yuvLaunch(iqo_cuda_yuv420 *h, size_t n, const uint8_t *src, uint8_t *dst, cudaStream_t s)
{
    iqo_cuda_resizer *y = h->luma, *c = h->chroma;
    CUDA_TRY(cudaEventRecord(h->fork, s));
    int rc = launch(y, n, 0, size_t(y->plan.y.D), 0, size_t(y->plan.y.S), h->srcStX, h->srcSize, src, h->dstStX, h->dstSize, dst, s);
    for (int p = 0; p < 2 && rc == IQO_CUDA_OK; ++p) {
        CUDA_TRY(cudaStreamWaitEvent(h->side[p], h->fork, 0));
        rc = launch(c, n, 0, size_t(c->plan.y.D), 0, size_t(c->plan.y.S), h->srcStX / 2, h->srcSize,
                    src + h->srcSizeY + p * h->srcSizeU, h->dstStX / 2, h->dstSize, dst + h->dstSizeY + p * h->dstSizeU, h->side[p]);
        // join even after a failed launch: the caller's stream must not run ahead of the side stream
        CUDA_TRY(cudaEventRecord(h->join[p], h->side[p]));
        CUDA_TRY(cudaStreamWaitEvent(s, h->join[p], 0));
    }
    return rc;
}

}  // namespace

extern "C" {

int iqo_cuda_yuv420_create(iqo_cuda_yuv420 **out, int kind, unsigned degree, size_t srcW, size_t srcH, size_t dstW, size_t dstH)
{
    IQO_GUARD_BEGIN
    if (!out) return fail(IQO_CUDA_E_ARG, "NULL output handle");
    *out = 0;
    if (!srcW || !srcH || !dstW || !dstH) return fail(IQO_CUDA_E_ARG, "image sizes must be non-zero");
    iqo_cuda_yuv420 *h = new iqo_cuda_yuv420();
    memset(h, 0, sizeof *h);
    h->srcStX = srcW + srcW % 2;
    h->srcStY = srcH + srcH % 2;
    h->dstStX = dstW + dstW % 2;
    h->dstStY = dstH + dstH % 2;
    h->srcSizeY = h->srcStX * h->srcStY;
    h->srcSizeU = h->srcSizeY / 4;
    h->srcSize = h->srcSizeY + 2 * h->srcSizeU;
    h->dstSizeY = h->dstStX * h->dstStY;
    h->dstSizeU = h->dstSizeY / 4;
    h->dstSize = h->dstSizeY + 2 * h->dstSizeU;
    int rc = iqo_cuda_create(&h->luma, kind, degree, srcW, srcH, dstW, dstH, 1);
    if (rc == IQO_CUDA_OK)
        rc = iqo_cuda_create(&h->chroma, kind, degree, h->srcStX / 2, h->srcStY / 2, h->dstStX / 2, h->dstStY / 2, 2);
    if (rc == IQO_CUDA_OK) {
        DeviceGuard guard(h->luma->device);
        cudaError_t e = cudaEventCreateWithFlags(&h->fork, cudaEventDisableTiming);
        for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
            e = cudaStreamCreateWithFlags(&h->side[i], cudaStreamNonBlocking);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->join[i], cudaEventDisableTiming);
        }
        if (e != cudaSuccess) rc = fail(IQO_CUDA_E_CUDA, "cannot create the chroma streams: %s", cudaGetErrorString(e));
    }
    if (rc != IQO_CUDA_OK) {
        std::string keep = t_lastError;
        iqo_cuda_yuv420_destroy(h);
        t_lastError = keep;
        return rc;
    }
    *out = h;
    return IQO_CUDA_OK;
    IQO_GUARD_END
}

void iqo_cuda_yuv420_destroy(iqo_cuda_yuv420 *h)
{
    if (!h) return;
    if (h->luma) {
        DeviceGuard guard(h->luma->device);
        cudaStreamSynchronize(h->luma->stream[0]);
        cudaStreamSynchronize(h->luma->stream[1]);
        for (int i = 0; i < 2; ++i) {
            if (h->side[i]) {
                cudaStreamSynchronize(h->side[i]);
                cudaStreamDestroy(h->side[i]);
            }
            if (h->join[i]) cudaEventDestroy(h->join[i]);
            cudaFree(h->dIn[i]);
            cudaFree(h->dOut[i]);
        }
        if (h->fork) cudaEventDestroy(h->fork);
        cudaGetLastError();
    }
    iqo_cuda_destroy(h->luma);
    iqo_cuda_destroy(h->chroma);
    delete h;
}

int iqo_cuda_yuv420_frame_bytes(const iqo_cuda_yuv420 *h, size_t *srcBytes, size_t *dstBytes)
{
    if (!h) return fail(IQO_CUDA_E_ARG, "NULL handle");
    if (srcBytes) *srcBytes = h->srcSize;
    if (dstBytes) *dstBytes = h->dstSize;
    return IQO_CUDA_OK;
}

int iqo_cuda_yuv420_resize(iqo_cuda_yuv420 *h, size_t nFrames, const uint8_t *src, uint8_t *dst, void *stream)
{
    IQO_GUARD_BEGIN
    if (!h) return fail(IQO_CUDA_E_ARG, "NULL handle");
    if (!src || !dst) return fail(IQO_CUDA_E_ARG, "NULL image pointer");
    if (nFrames == 0) return IQO_CUDA_OK;
    DeviceGuard guard(h->luma->device);
    const bool srcDev = isDevicePointer(src), dstDev = isDevicePointer(dst);
    if (srcDev && dstDev) return yuvLaunch(h, nFrames, src, dst, (cudaStream_t)stream);
    if (srcDev != dstDev) return fail(IQO_CUDA_E_ARG, "src and dst must both be host or both be device memory");
    // host frames: double-buffered chunks on the luma handle's two streams
    size_t chunk = std::max<size_t>(1, (size_t(96) << 20) / (h->srcSize + h->dstSize));
    chunk = std::min(chunk, nFrames);
    if (h->slotFrames < chunk) {
        for (int i = 0; i < 2; ++i) {
            cudaFree(h->dIn[i]);
            cudaFree(h->dOut[i]);
            h->dIn[i] = h->dOut[i] = 0;
        }
        h->slotFrames = 0;
        for (int i = 0; i < 2; ++i)
            if (cudaMalloc(&h->dIn[i], chunk * h->srcSize) != cudaSuccess || cudaMalloc(&h->dOut[i], chunk * h->dstSize) != cudaSuccess) {
                cudaGetLastError();
                return fail(IQO_CUDA_E_NOMEM, "cannot allocate YUV staging for %zu frames", chunk);
            }
        h->slotFrames = chunk;
    }
    int slot = 0;
    for (size_t done = 0; done < nFrames; slot ^= 1) {
        const size_t n = std::min(chunk, nFrames - done);
        cudaStream_t s = h->luma->stream[slot];
        CUDA_TRY(cudaMemcpyAsync(h->dIn[slot], src + done * h->srcSize, n * h->srcSize, cudaMemcpyHostToDevice, s));
        int rc = yuvLaunch(h, n, h->dIn[slot], h->dOut[slot], s);
        if (rc) return rc;
        CUDA_TRY(cudaMemcpyAsync(dst + done * h->dstSize, h->dOut[slot], n * h->dstSize, cudaMemcpyDeviceToHost, s));
        done += n;
    }
    CUDA_TRY(cudaStreamSynchronize(h->luma->stream[0]));
    CUDA_TRY(cudaStreamSynchronize(h->luma->stream[1]));
    return IQO_CUDA_OK;
    IQO_GUARD_END
}

// ---- multi-device drivers: one host thread + one handle per device, no collectives ----

int iqo_cuda_resize_batch_multi(int kind, unsigned degree,
                                size_t srcW, size_t srcH, size_t dstW, size_t dstH, size_t pxScale,
                                size_t nFrames,
                                size_t srcSt, size_t srcFrameStride, const uint8_t *src,
                                size_t dstSt, size_t dstFrameStride, uint8_t *dst,
                                int nDevices, const int *devices)
{
    IQO_GUARD_BEGIN
    if (nDevices <= 0) return fail(IQO_CUDA_E_ARG, "nDevices must be positive");
    if (!src || !dst) return fail(IQO_CUDA_E_ARG, "NULL image pointer");
    std::vector<int> rcs(nDevices, IQO_CUDA_OK);
    std::vector<std::string> msgs(nDevices);
    std::vector<std::thread> threads;
    const size_t per = (nFrames + nDevices - 1) / nDevices;
    for (int i = 0; i < nDevices; ++i) {
        threads.push_back(std::thread([=, &rcs, &msgs]() {
            const size_t f0 = std::min(nFrames, per * i), f1 = std::min(nFrames, f0 + per);
            const int dev = devices ? devices[i] : i;
            iqo_cuda_resizer *r = 0;
            int rc = iqo_cuda_create_on(&r, dev, kind, degree, srcW, srcH, dstW, dstH, pxScale);
            if (rc == IQO_CUDA_OK && f1 > f0)
                rc = iqo_cuda_resize_batch_host(r, f1 - f0, srcSt, srcFrameStride, src + f0 * srcFrameStride,
                                                dstSt, dstFrameStride, dst + f0 * dstFrameStride);
            if (rc != IQO_CUDA_OK) msgs[i] = iqo_cuda_last_error();
            iqo_cuda_destroy(r);
            rcs[i] = rc;
        }));
    }
    for (size_t i = 0; i < threads.size(); ++i) threads[i].join();
    for (int i = 0; i < nDevices; ++i)
        if (rcs[i] != IQO_CUDA_OK) return fail(rcs[i], "device %d: %s", devices ? devices[i] : i, msgs[i].c_str());
    return IQO_CUDA_OK;
    IQO_GUARD_END
}

int iqo_cuda_resize_bands_multi(int kind, unsigned degree,
                                size_t srcW, size_t srcH, size_t dstW, size_t dstH, size_t pxScale,
                                size_t srcSt, const uint8_t *src, size_t dstSt, uint8_t *dst,
                                int nDevices, const int *devices)
{
    IQO_GUARD_BEGIN
    if (nDevices <= 0) return fail(IQO_CUDA_E_ARG, "nDevices must be positive");
    if (!src || !dst) return fail(IQO_CUDA_E_ARG, "NULL image pointer");
    if (srcSt < srcW || dstSt < dstW) return fail(IQO_CUDA_E_ARG, "stride smaller than width");
    std::vector<int> rcs(nDevices, IQO_CUDA_OK);
    std::vector<std::string> msgs(nDevices);
    std::vector<std::thread> threads;
    const size_t per = (dstH + nDevices - 1) / nDevices;
    for (int i = 0; i < nDevices; ++i) {
        threads.push_back(std::thread([=, &rcs, &msgs]() {
            const size_t y0 = std::min(dstH, per * i), y1 = std::min(dstH, y0 + per);
            const int dev = devices ? devices[i] : i;
            if (y1 <= y0) return;
            iqo_cuda_resizer *r = 0;
            uint8_t *dS = 0, *dD = 0;
            int rc = iqo_cuda_create_on(&r, dev, kind, degree, srcW, srcH, dstW, dstH, pxScale);
            if (rc == IQO_CUDA_OK) {
                cudaSetDevice(dev);
                size_t s0 = 0, sn = 0;
                rc = iqo_cuda_band_src_rows(r, y0, y1 - y0, &s0, &sn);
                const size_t sp = alignUp(srcW, 16), dp = alignUp(dstW, 16);
                cudaStream_t st = r->stream[0];
                if (rc == IQO_CUDA_OK &&
                    (cudaMalloc(&dS, sp * sn) != cudaSuccess || cudaMalloc(&dD, dp * (y1 - y0)) != cudaSuccess))
                    rc = fail(IQO_CUDA_E_NOMEM, "band buffers: %s", cudaGetErrorString(cudaGetLastError()));
                // the band's halo rows are part of its own upload: no device-to-device exchange
                if (rc == IQO_CUDA_OK &&
                    cudaMemcpy2DAsync(dS, sp, src + s0 * srcSt, srcSt, srcW, sn, cudaMemcpyHostToDevice, st) != cudaSuccess)
                    rc = fail(IQO_CUDA_E_CUDA, "band upload: %s", cudaGetErrorString(cudaGetLastError()));
                if (rc == IQO_CUDA_OK) rc = iqo_cuda_resize_band(r, y0, y1 - y0, s0, sn, sp, dS, dp, dD, st);
                if (rc == IQO_CUDA_OK &&
                    (cudaMemcpy2DAsync(dst + y0 * dstSt, dstSt, dD, dp, dstW, y1 - y0, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
                     cudaStreamSynchronize(st) != cudaSuccess))
                    rc = fail(IQO_CUDA_E_CUDA, "band download: %s", cudaGetErrorString(cudaGetLastError()));
            }
            if (rc != IQO_CUDA_OK) msgs[i] = iqo_cuda_last_error();
            cudaFree(dS);
            cudaFree(dD);
            iqo_cuda_destroy(r);
            rcs[i] = rc;
        }));
    }
    for (size_t i = 0; i < threads.size(); ++i) threads[i].join();
    for (int i = 0; i < nDevices; ++i)
        if (rcs[i] != IQO_CUDA_OK) return fail(rcs[i], "device %d: %s", devices ? devices[i] : i, msgs[i].c_str());
    return IQO_CUDA_OK;
    IQO_GUARD_END
}

}  // extern "C"
