// sm_100a kernels for the libiqo resize hot path.
//
// Arithmetic contract (bit-exact with the reference's Generic implementation, SURVEY.md
// Appendix A): integer coefficient tables, 16-bit wrapped intermediate ("work") after the
// vertical pass, 32-bit horizontal sums, one rounding shift, the truncating divisions of the
// Lanczos border rows/columns.  Integer sums are order independent, so any tap order / packing
// is exact.
#include "kernels.cuh"

#include <stdlib.h>

#include <algorithm>
#include <atomic>
#include <type_traits>

namespace iqo_b200 {

namespace {

std::atomic<unsigned long long> g_launches(0);

// cudaFuncSetAttribute is per device: remember, per kernel instantiation, where it was done
struct PerDeviceOnce {
    std::atomic<unsigned long long> mask[2];
    PerDeviceOnce() { mask[0] = 0; mask[1] = 0; }
    bool done(int dev) const { return dev >= 0 && dev < 128 && ((mask[dev >> 6].load() >> (dev & 63)) & 1ull); }
    void set(int dev) { if (dev >= 0 && dev < 128) mask[dev >> 6].fetch_or(1ull << (dev & 63)); }
};

int currentDevice()
{
    int dev = 0;
    cudaGetDevice(&dev);
    return dev;
}

__device__ __forceinline__ int clampi(int v, int lo, int hi)
{
    return min(max(v, lo), hi);
}

// Final conversion of a horizontal sum to a pixel:
//   main columns  : clampU8(int16((sum + half) >> shift))        (convertToInt, ..._Generic.cpp:223-227)
//   Lanczos border: clampU8(int16((sum + half) / (deno * 64)))    (roundedDiv, :216-220,572) -- C division
__device__ __forceinline__ uint8_t finishPixel(int sum, int deno, int shift)
{
    const int half = 1 << (shift - 1);
    int v;
    if (deno != 0) {
        v = (sum + half) / (deno * 64);
    } else {
        v = (sum + half) >> shift;
    }
    v = (int)(short)v;
    return (uint8_t)clampi(v, 0, 255);
}

// ---------------------------------------------------------------------------------------
// Generic fused tile kernel: any kind, ratio, stride and phase count.
//
// One CTA produces a tileH x tileW destination tile of one frame:
//   1. vertical pass: for the tile's source-column window [x0, x0+ww) and each of its rows,
//      work[r][c] = wrap16( sum_i coefY[row(y)][i] * src[clamp(first(y)+i)][x0+c] ) (+ border division)
//      kept in shared memory only (the intermediate never reaches HBM);
//   2. horizontal pass from shared memory, rounding, clamp, byte store.
// Warps own rows; lanes run along columns so global loads/stores of a warp are contiguous.
// ---------------------------------------------------------------------------------------
template <bool kSigned>
__global__ void __launch_bounds__(256) resizeGenericKernel(ResizeArgs a, int tileW, int tileH, int workW)
{
    extern __shared__ __align__(16) unsigned char smemRaw[];
    typedef typename std::conditional<kSigned, short, unsigned short>::type work_t;
    work_t *work = reinterpret_cast<work_t *>(smemRaw);

    const int tx0 = blockIdx.x * tileW;
    const int ty0 = blockIdx.y * tileH;
    const int tw = min(tileW, a.x.D - tx0);
    const int th = min(tileH, a.dstRows - ty0);
    const uint8_t *__restrict__ src = a.src + (long long)blockIdx.z * a.srcFrameStride;
    uint8_t *__restrict__ dst = a.dst + (long long)blockIdx.z * a.dstFrameStride;

    const int SW = a.x.S, SH = a.y.S;
    const int Nx = a.x.N, Ny = a.y.N;
    // first[] is non-decreasing, so the tile's window is spanned by its first and last column
    const int x0 = clampi(__ldg(a.x.first + tx0), 0, SW - 1);
    const int x1 = clampi(__ldg(a.x.first + tx0 + tw - 1) + Nx - 1, 0, SW - 1);
    const int ww = x1 - x0 + 1;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarps = blockDim.x >> 5;

    // ---- vertical pass ----
    for (int r = warp; r < th; r += nwarps) {
        const int y = a.dstRow0 + ty0 + r;
        const int fy = __ldg(a.y.first + y);
        const int ry = __ldg(a.y.row + y);
        const int32_t *__restrict__ cy = a.y.coef + (long long)ry * Ny;
        const int deno = __ldg(a.y.deno + ry);
        for (int c = lane; c < ww; c += 32) {
            const uint8_t *col = src + x0 + c;
            int acc = 0;
            for (int i = 0; i < Ny; ++i) {
                const int sy = clampi(fy + i, 0, SH - 1) - a.srcRow0;
                acc += __ldg(cy + i) * (int)__ldg(col + (long long)sy * a.srcPitch);
            }
            if (deno != 0) {
                // resizeYborder: int16 numerator, * kBias, C division by the int16 denominator
                acc = ((int)(short)acc * 64) / deno;
            }
            work[r * workW + c] = (work_t)acc;
        }
    }
    __syncthreads();

    // ---- horizontal pass ----
    for (int r = warp; r < th; r += nwarps) {
        const work_t *wrow = work + r * workW;
        uint8_t *out = dst + (long long)(ty0 + r) * a.dstPitch + tx0;
        for (int dx = lane; dx < tw; dx += 32) {
            const int x = tx0 + dx;
            const int fx = __ldg(a.x.first + x);
            const int rx = __ldg(a.x.row + x);
            const int32_t *__restrict__ cx = a.x.coef + (long long)rx * Nx;
            int acc = 0;
            for (int i = 0; i < Nx; ++i) {
                const int sx = clampi(fx + i, 0, SW - 1) - x0;
                acc += __ldg(cx + i) * (int)wrow[sx];
            }
            out[dx] = finishPixel(acc, a.lanczos ? __ldg(a.x.deno + rx) : 0, a.shift);
        }
    }
}


// ---------------------------------------------------------------------------------------
// Float ("SIMD-semantics") mode (plan.hpp FloatPlan): the tile organisation of the generic kernel with
// the arithmetic of the reference's AVX-512 path -- float FMA accumulation in tap order over normalised float tables
// (src/IQOLanczosResizerImpl_AVX512.cpp:385-431 vertical, :547-590 horizontal), division by the in-range coefficient
// sum on border rows and columns, round-to-nearest-even, saturation (:47-60).  Opt-in, outside the parity contract.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) resizeFloatKernel(FloatArgs a, int tileW, int tileH, int workW)
{
    extern __shared__ __align__(16) unsigned char smemRawF[];
    float *work = reinterpret_cast<float *>(smemRawF);
    const int tx0 = blockIdx.x * tileW;
    const int ty0 = blockIdx.y * tileH;
    const int tw = min(tileW, a.x.D - tx0);
    const int th = min(tileH, a.dstRows - ty0);
    const uint8_t *__restrict__ src = a.src + (long long)blockIdx.z * a.srcFrameStride;
    uint8_t *__restrict__ dst = a.dst + (long long)blockIdx.z * a.dstFrameStride;
    const int SW = a.x.S, SH = a.y.S;
    const int Nx = a.x.N, Ny = a.y.N;
    const int x0 = clampi(__ldg(a.x.first + tx0), 0, SW - 1);
    const int x1 = clampi(__ldg(a.x.first + tx0 + tw - 1) + Nx - 1, 0, SW - 1);
    const int ww = x1 - x0 + 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarps = blockDim.x >> 5;
    for (int r = warp; r < th; r += nwarps) {
        const int y = a.dstRow0 + ty0 + r;
        const int fy = __ldg(a.y.first + y);
        const int ry = __ldg(a.y.row + y);
        const float *__restrict__ cy = a.coefY + (long long)ry * Ny;
        const float deno = __ldg(a.denoY + ry);
        for (int c = lane; c < ww; c += 32) {
            const uint8_t *col = src + x0 + c;
            float acc = 0.0f;
            for (int i = 0; i < Ny; ++i) {
                const int sy = clampi(fy + i, 0, SH - 1) - a.srcRow0;
                acc = fmaf((float)__ldg(col + (long long)sy * a.srcPitch), __ldg(cy + i), acc);
            }
            if (deno != 0.0f) acc = acc / deno;
            work[r * workW + c] = acc;
        }
    }
    __syncthreads();
    for (int r = warp; r < th; r += nwarps) {
        const float *wrow = work + r * workW;
        uint8_t *out = dst + (long long)(ty0 + r) * a.dstPitch + tx0;
        for (int dx = lane; dx < tw; dx += 32) {
            const int x = tx0 + dx;
            const int fx = __ldg(a.x.first + x);
            const int rx = __ldg(a.x.row + x);
            const float *__restrict__ cx = a.coefX + (long long)rx * Nx;
            float acc = 0.0f;
            for (int i = 0; i < Nx; ++i) {
                const int sx = clampi(fx + i, 0, SW - 1) - x0;
                acc = fmaf(wrow[sx], __ldg(cx + i), acc);
            }
            const float deno = __ldg(a.denoX + rx);
            if (deno != 0.0f) acc = acc / deno;
            out[dx] = (uint8_t)clampi(__float2int_rn(acc), 0, 255);
        }
    }
}

// ---------------------------------------------------------------------------------------
// Specialised kernel: Lanczos 2:1 down-sampling on both axes, single-phase tables
// (BASELINE configs 3 and 4).  One CTA = 256 threads = one tile of 120 x tileRows
// destination pixels of one frame.
//
//  vertical pass   thread = 4 adjacent source columns (one aligned 32-bit word per source row)
//                  x one strip of destination rows.  Source rows are read straight from global
//                  memory (a warp reads 128 contiguous bytes per row), four rows at a time;
//                  the 4x4 byte block is transposed in registers (8 PRMT) so that one register
//                  holds four vertically adjacent bytes of a column, and each destination row is
//                  NG dp4a(u8 x s8) per column against packed coefficient words.  A sliding
//                  window of NG groups lives in registers, so every source row is loaded and
//                  transposed once per strip.  Results (+bias, so they are non-negative u16) go
//                  to shared memory as 16-bit pairs; the intermediate never reaches HBM.
//  horizontal pass thread = 8 adjacent destination pixels of one row: four conflict-free 16-byte
//                  shared loads (XOR-swizzled chunks), then per pixel dp2a(u16 x u8 / u16 x s8)
//                  over "natural" pairs (source columns 2m, 2m+1) with the 14-bit coefficients
//                  split into an unsigned low and a signed high byte plane.  For palindromic
//                  tables mirrored pairs are first added as packed 16-bit halves (one IADD for
//                  two taps), which halves the dp2a count.  One rounding shift, saturating pack,
//                  8-byte coalesced store.
//  borders         Lanczos border rows use masked coefficient words + the reference's truncating
//                  division in the same dp4a structure; border columns (<= 6 per row) are
//                  recomputed by a scalar loop from the generic tables.
// All arithmetic is integer and order independent => bit-exact with the reference.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

__device__ __forceinline__ int dp4a_us(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

__device__ __forceinline__ int dp2a_lo_uu(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

__device__ __forceinline__ int dp2a_hi_us(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp2a.hi.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// d = (sat_u8(hi) << 8) | sat_u8(lo) | (upper << 16)
__device__ __forceinline__ uint32_t packSatU8(int hi, int lo, uint32_t upper)
{
    uint32_t d;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(hi), "r"(lo), "r"(upper));
    return d;
}

// physical 32-bit word index inside a 128-word W row: 16-byte chunks are XOR-swizzled so that
// both the 8-byte column writes and the 32-byte-strided 16-byte reads are bank-conflict free
__device__ __forceinline__ int swzWord(int word)
{
    const int chunk = word >> 2;
    return ((chunk ^ ((chunk >> 3) & 1)) << 2) | (word & 3);
}

constexpr int kMmaMaxKStepsDev = 3;  // plan.hpp kMmaMaxKSteps
constexpr int kHalfTileW = 120;     // destination columns per tile
constexpr int kHalfRowWords = 128;  // W row: 256 u16 = 128 words
constexpr int kHalfMaxRows = 64;    // destination rows per tile (<=)
constexpr int kHalfSrcRowBytes = 256;                                   // source window of a tile, bytes per row
constexpr int kHalfSrcMaxRows = 4 * (kHalfMaxRows / 2 + 3);             // 140: groups of the tile + prefetch slack
constexpr int kHalfTileBytes = kHalfSrcMaxRows * kHalfSrcRowBytes;      // 35840 (multiple of 128)
constexpr int kHalfWBytes = kHalfMaxRows * kHalfRowWords * 4;           // 32768

// Vertical pass of one strip (see resizeHalfKernel).
//   SMEM  = the tile's source window was staged in shared memory by TMA (out-of-image rows and
//           columns are zero filled by the hardware); otherwise rows are read from global memory.
//   EDGE  = the strip may contain Lanczos border rows (masked coefficient words + truncating
//           division) and, when reading global memory, needs its row indices clamped.
// `base` points at this thread's 4-byte column word in row 0 of the tile (SMEM) or of the frame.
template <int NG, bool EDGE, bool SMEM>
__device__ __forceinline__ void halfVerticalStrip(const HalfArgs &a, const uint8_t *__restrict__ base,
                                                  uint32_t *__restrict__ wout, int ty0, int k0, int k1)
{
    const int B = a.workBias;
    const int SHm1 = a.SH - 1;
    const long long pitch = SMEM ? (long long)kHalfSrcRowBytes : a.srcPitch;
    uint32_t raw[4];
    // SMEM: tile group j holds source rows 4(ty0/2 + qmin + j)...; global: absolute group index
    int g = SMEM ? k0 : (ty0 >> 1) + k0 + a.qmin;
    const uint8_t *gp = base + (long long)(4 * g) * pitch;

    auto fetch = [&]() {  // issue the loads of the four rows of group g, then advance to the next group
        if (SMEM) {
#pragma unroll
            for (int j = 0; j < 4; ++j) raw[j] = *reinterpret_cast<const uint32_t *>(gp + j * kHalfSrcRowBytes);
            gp += 4 * kHalfSrcRowBytes;
        } else if (EDGE) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int row = min(max(4 * g + j, 0), SHm1);
                raw[j] = __ldg(reinterpret_cast<const uint32_t *>(base + (long long)row * pitch));
            }
        } else {
            const uint8_t *p1 = gp + pitch;
            raw[0] = __ldg(reinterpret_cast<const uint32_t *>(gp));
            raw[1] = __ldg(reinterpret_cast<const uint32_t *>(p1));
            raw[2] = __ldg(reinterpret_cast<const uint32_t *>(gp + 2 * pitch));
            raw[3] = __ldg(reinterpret_cast<const uint32_t *>(p1 + 2 * pitch));
            gp += 4 * pitch;
        }
        ++g;
    };
    auto transposed = [&]() -> uint4 {  // .x/.y/.z/.w = four vertical bytes of column 0/1/2/3
        const uint32_t t0 = prmt(raw[0], raw[1], 0x5140), t1 = prmt(raw[0], raw[1], 0x7362);
        const uint32_t t2 = prmt(raw[2], raw[3], 0x5140), t3 = prmt(raw[2], raw[3], 0x7362);
        uint4 c;
        c.x = prmt(t0, t2, 0x5410);
        c.y = prmt(t0, t2, 0x7632);
        c.z = prmt(t1, t3, 0x5410);
        c.w = prmt(t1, t3, 0x7632);
        return c;
    };

    uint4 win[NG];
#pragma unroll
    for (int j = 0; j < NG - 1; ++j) {
        fetch();
        win[j] = transposed();
    }
    fetch();

    // one destination row pair: `s` (compile-time after unrolling) is the ring position of the window
    auto step = [&](const int s, const int kk) {
        win[(s + NG - 1) % NG] = transposed();
        fetch();  // prefetch the next pair's newest group (overshoots by one group at the strip end)
#pragma unroll
        for (int par = 0; par < 2; ++par) {
            const int rl = 2 * kk + par;  // local destination row
            uint32_t c0 = a.cwY[par][0], c1 = NG > 1 ? a.cwY[par][1] : 0u, c2 = NG > 2 ? a.cwY[par][2] : 0u;
            int deno = 0;
            uint32_t magic = 0;
            if (EDGE) {
                const int y = ty0 + rl;
                if (y < a.DH && (y < a.mbY || y >= a.meY)) {
                    const int row = __ldg(a.rowY + y);
                    deno = __ldg(a.denoY + row);
                    magic = __ldg(a.magicY + row);
                    c0 = __ldg(a.borderY + row * 3);
                    c1 = __ldg(a.borderY + row * 3 + 1);
                    c2 = __ldg(a.borderY + row * 3 + 2);
                }
            }
            const int init = (EDGE && deno) ? 0 : B;
            int v0 = init, v1 = init, v2 = init, v3 = init;
            {
                const uint4 q = win[s % NG];
                v0 = dp4a_us(q.x, c0, v0);
                v1 = dp4a_us(q.y, c0, v1);
                v2 = dp4a_us(q.z, c0, v2);
                v3 = dp4a_us(q.w, c0, v3);
            }
            if (NG > 1) {
                const uint4 q = win[(s + 1) % NG];
                v0 = dp4a_us(q.x, c1, v0);
                v1 = dp4a_us(q.y, c1, v1);
                v2 = dp4a_us(q.z, c1, v2);
                v3 = dp4a_us(q.w, c1, v3);
            }
            if (NG > 2) {
                const uint4 q = win[(s + 2) % NG];
                v0 = dp4a_us(q.x, c2, v0);
                v1 = dp4a_us(q.y, c2, v1);
                v2 = dp4a_us(q.z, c2, v2);
                v3 = dp4a_us(q.w, c2, v3);
            }
            if (EDGE && deno) {
                // resizeYborder: int16 numerator * 64 / denominator, C (truncating) division;
                // |numerator * 64| <= 2^21 and 1 <= deno <= 127, so floor(n / deno) is the
                // multiply-high by floor(2^32 / deno) + 1 (deno == 1: magic == 0, identity)
                auto bdiv = [&](int v) -> int {
                    const int n = (int)(short)v * 64;
                    const uint32_t m = (uint32_t)abs(n);
                    const int q = magic ? (int)__umulhi(m, magic) : (int)m;
                    return (int)(short)(n < 0 ? -q : q) + B;
                };
                v0 = bdiv(v0);
                v1 = bdiv(v1);
                v2 = bdiv(v2);
                v3 = bdiv(v3);
            }
            uint2 o;
            o.x = prmt((uint32_t)v0, (uint32_t)v1, 0x5410);
            o.y = prmt((uint32_t)v2, (uint32_t)v3, 0x5410);
            *reinterpret_cast<uint2 *>(wout + rl * kHalfRowWords) = o;
        }
    };

    int k = k0;
    for (; k + NG <= k1; k += NG) {  // whole turns of the register ring, no per-step checks
#pragma unroll
        for (int s = 0; s < NG; ++s) step(s, k + s);
    }
#pragma unroll
    for (int s = 0; s < NG - 1; ++s)  // at most NG-1 remaining pairs
        if (k + s < k1) step(s, k + s);
}

// Vertical pass of a tile: warps = 2 column halves x (2 or 4) strips of destination row pairs.
template <int NG, bool SMEM>
__device__ __forceinline__ void halfVertical(const HalfArgs &a, const uint8_t *__restrict__ srcOrTile, uint32_t *W,
                                             int xs0, int tx0, int ty0, int th)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cw = ((warp & 1) << 5) | lane;  // column word 0..63
    const int strip = warp >> 1;              // 0..strips-1
    const int stripShift = blockDim.x >> 8;   // log2(strips): 4 strips (256 threads) or 2 (128 threads)
    const int pairs = (th + 1) >> 1;
    const int k0 = (strip * pairs) >> (1 + stripShift), k1 = ((strip + 1) * pairs) >> (1 + stripShift);
    if (k0 >= k1) return;
    // column words right of the last source column any pixel of this tile needs do no work
    // (matters for the nearly empty last tile of a shifted grid)
    const int lastCol = 2 * (min(tx0 + kHalfTileW, a.DW) - 1) + a.NX / 2;
    if (xs0 + 4 * cw > lastCol) return;
    uint32_t *wout = W + swzWord(2 * cw);
    const bool borderRows = (ty0 + 2 * k0 < a.mbY) || (ty0 + 2 * k1 > a.meY);
    if (SMEM) {
        const uint8_t *base = srcOrTile + 4 * cw;
        if (borderRows)
            halfVerticalStrip<NG, true, true>(a, base, wout, ty0, k0, k1);
        else
            halfVerticalStrip<NG, false, true>(a, base, wout, ty0, k0, k1);
    } else {
        // columns outside the image only ever meet zero coefficients: read column 0 instead
        const int col = xs0 + 4 * cw;
        const uint8_t *base = srcOrTile + ((col >= 0 && col < a.SW) ? col : 0);
        // source rows this strip touches, including the one-group prefetch overshoot
        const int gFirst = (ty0 >> 1) + k0 + a.qmin;
        const int gLast = (ty0 >> 1) + k1 + a.qmin + NG - 1;
        if (borderRows || gFirst < 0 || 4 * gLast + 3 >= a.SH)
            halfVerticalStrip<NG, true, false>(a, base, wout, ty0, k0, k1);
        else
            halfVerticalStrip<NG, false, false>(a, base, wout, ty0, k0, k1);
    }
}

// Border columns of a tile (Lanczos: at most N/2 per side), recomputed after the main pass by
// the whole CTA, one (row, column) item per thread so that no warp diverges: the masked taps of
// resizeXborder (reference ..._Generic.cpp:539-574) as planner-made pair words through the same
// dp2a planes, then its truncating division.  [c0, c1) are the tile's border columns, `bx` the
// table entry of column c0.
template <int NWX>
__device__ __noinline__ void halfBorderColumns(const HalfArgs &a, const uint32_t *W, uint8_t *dstTile, const int32_t *bx,
                                               int tx0, int c0, int c1, int th, int tid, int nthr)
{
    constexpr int kBase = 4 - (NWX - 1) / 2;
    const int nb = c1 - c0;
    for (int item = tid; item < nb * th; item += nthr) {
        const int r = item / nb;
        const int j = item - r * nb;
        const int d = c0 + j;
        const uint32_t *wr = W + r * kHalfRowWords;
        const int32_t *e = bx + j * 9;
        int lo = __ldg(e + 8), hi = 0;
#pragma unroll
        for (int i = 0; i < NWX; ++i) {
            const uint32_t word = wr[swzWord(d - tx0 + kBase + i)];
            const uint32_t cw = (uint32_t)__ldg(e + i);
            lo = dp2a_lo_uu(word, cw, lo);
            hi = dp2a_hi_us(word, cw, hi);
        }
        const int v = (int)(short)((lo + (hi << 8)) / __ldg(e + 7));
        dstTile[(long long)r * a.dstPitch + (d - tx0)] = (uint8_t)min(max(v, 0), 255);
    }
}

__device__ __noinline__ void halfStoreBytes(uint8_t *out, uint2 o, int lo, int hi)
{
    for (int p = lo; p < hi; ++p) out[p] = (uint8_t)(((p < 4 ? o.x : o.y) >> (8 * (p & 3))) & 0xffu);
}

// Horizontal pass of eight adjacent destination pixels from the 16 pair words n[] that start at
// the group's first 16-byte chunk; returns the eight saturated bytes.
template <int NWX, bool SYM, bool ENDHI>
__device__ __forceinline__ uint2 halfGroupPixels(const HalfArgs &a, const uint32_t (&n)[16])
{
    constexpr int kBase = 4 - (NWX - 1) / 2;  // pair word of tap pair 0 for pixel 0 (wa + 4)
    int v[8];
    if (SYM) {
        constexpr int m = NWX / 2;
        // swapped halves of the words that serve as mirror partners
        uint32_t sw[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) sw[i] = (i >= kBase + m + 1 && i <= kBase + NWX - 2 + 7) ? prmt(n[i], n[i], 0x1032) : 0u;
        // mirrored pairs are added as packed u16 halves (no carry can cross: every
        // half-sum fits 16 bits).  The third addend is a kernel argument that is always
        // 0: a three-input add can only be an IADD3, which keeps these adds off the
        // multiplier pipe that the dp2a/dp4a instructions saturate.
        uint32_t sum[8][m > 1 ? m - 1 : 1];
#pragma unroll
        for (int j = 1; j < m; ++j)
#pragma unroll
            for (int p = 0; p < 8; ++p) sum[p][j - 1] = n[kBase + p + j] + sw[kBase + p + NWX - 1 - j] + a.zero;
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            const uint32_t ctr = n[kBase + p + m];
            const uint32_t ends = prmt(n[kBase + p], n[kBase + p + NWX - 1], 0x3254);
            // low byte plane first; the high plane continues from (low >> 8):
            // floor((lo + 256 hi) / 2^20) == floor((floor(lo / 256) + hi) / 2^12)
            int acc = a.accInit;
#pragma unroll
            for (int j = 1; j < m; ++j) acc = dp2a_lo_uu(sum[p][j - 1], a.cwXs[j - 1], acc);
            acc = dp2a_lo_uu(ctr, a.cwXs[m - 1], acc);
            acc = dp2a_lo_uu(ends, a.cwXs[m], acc);
            acc >>= 8;
#pragma unroll
            for (int j = 1; j < m; ++j) acc = dp2a_hi_us(sum[p][j - 1], a.cwXs[j - 1], acc);
            acc = dp2a_hi_us(ctr, a.cwXs[m - 1], acc);
            if (ENDHI) acc = dp2a_hi_us(ends, a.cwXs[m], acc);  // skipped when both end taps fit the low byte plane
            v[p] = acc >> 12;
        }
    } else {
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            int acc = a.accInit;
#pragma unroll
            for (int j = 0; j < NWX; ++j) acc = dp2a_lo_uu(n[kBase + p + j], a.cwX[j], acc);
            acc >>= 8;
#pragma unroll
            for (int j = 0; j < NWX; ++j) acc = dp2a_hi_us(n[kBase + p + j], a.cwX[j], acc);
            v[p] = acc >> 12;
        }
    }
    uint2 o;
    o.x = packSatU8(v[1], v[0], packSatU8(v[3], v[2], 0u));
    o.y = packSatU8(v[5], v[4], packSatU8(v[7], v[6], 0u));
    return o;
}

// Group l of a W row `wr` (tiled variants: swizzled chunks), first pixel d0.
template <int NWX, bool SYM, bool ENDHI>
__device__ __forceinline__ void halfGroup(const HalfArgs &a, const uint32_t *wr, uint8_t *__restrict__ out, int d0, int l)
{
    int pc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) pc[j] = swzWord(4 * (2 * l + j));
    // 8-byte store when the pixel group is 8-aligned, two 4-byte stores when the tile grid is
    // shifted by 4 (TMA variant), bytes for partial groups or unaligned destinations
    const bool vecStore = a.dstVec && d0 >= 0 && (d0 + 8 <= a.DW);
    const bool vec8 = (d0 & 7) == 0;
    // a group cut in half by the image edge (shifted grid): one aligned 4-byte store
    const bool halfLo = a.dstVec && !vecStore && d0 >= 0 && d0 + 4 == a.DW;
    const bool halfHi = a.dstVec && !vecStore && d0 == -4 && a.DW >= 4;
    uint32_t n[16];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint4 q = *reinterpret_cast<const uint4 *>(wr + pc[j]);
        n[4 * j] = q.x;
        n[4 * j + 1] = q.y;
        n[4 * j + 2] = q.z;
        n[4 * j + 3] = q.w;
    }
    const uint2 o = halfGroupPixels<NWX, SYM, ENDHI>(a, n);
    if (vecStore) {
        if (vec8) {
            *reinterpret_cast<uint2 *>(out) = o;
        } else {
            *reinterpret_cast<uint32_t *>(out) = o.x;
            *reinterpret_cast<uint32_t *>(out + 4) = o.y;
        }
    } else if (halfLo || halfHi) {
        if (halfLo) *reinterpret_cast<uint32_t *>(out) = o.x;
        if (halfHi) *reinterpret_cast<uint32_t *>(out + 4) = o.y;
    } else {
        halfStoreBytes(out, o, max(0, -d0), min(8, a.DW - d0));
    }
}

// Horizontal pass of a tile + border columns.  The caller has synchronised after the vertical pass.
template <int NWX, bool SYM, bool ENDHI>
__device__ __forceinline__ void halfHorizontal(const HalfArgs &a, const uint32_t *W, uint8_t *__restrict__ dst,
                                               int xs0, int tx0, int ty0, int th)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int txEnd = min(tx0 + kHalfTileW, a.DW);
    const int groups = (txEnd - tx0 + 7) >> 3;  // 8-pixel groups holding at least one pixel of the image

    auto doGroup = [&](const uint32_t *wr, uint8_t *out, int l) { halfGroup<NWX, SYM, ENDHI>(a, wr, out, tx0 + 8 * l, l); };

    if (groups == 15) {
        // full tile: a half-warp per row (15 of 16 lanes busy), two rows per warp and iteration
        const int half = lane >> 4, l = lane & 15;
        if (l < 15) {
            const int rstep = blockDim.x >> 4;  // rows per iteration: two per warp
            const int r0 = 2 * warp + half;
            const uint32_t *wr = W + r0 * kHalfRowWords;
            uint8_t *out = dst + (long long)(ty0 + r0) * a.dstPitch + (tx0 + 8 * l);
            const long long ostep = (long long)rstep * a.dstPitch;
            for (int r = r0; r < th; r += rstep, wr += rstep * kHalfRowWords, out += ostep) doGroup(wr, out, l);
        }
    } else if (groups > 0) {
        // partial tile (image edge): pack (row, group) items densely over the CTA
        const uint32_t rcp = (65536u + groups - 1) / groups;
        for (int item = threadIdx.x; item < groups * th; item += blockDim.x) {
            const int r = (int)(((uint32_t)item * rcp) >> 16);
            const int l = item - r * groups;
            doGroup(W + r * kHalfRowWords, dst + (long long)(ty0 + r) * a.dstPitch + (tx0 + 8 * l), l);
        }
    }
    // border columns
    const bool left = tx0 < a.mbX, right = txEnd > a.meX;
    if (left || right) {
        __syncthreads();  // the main stores of these pixels come first (block-scope ordering)
        uint8_t *dstTile = dst + (long long)ty0 * a.dstPitch + tx0;
        if (left)
            halfBorderColumns<NWX>(a, W, dstTile, a.borderX + 9 * max(tx0, 0), tx0, max(tx0, 0), min(a.mbX, txEnd), th, threadIdx.x,
                                   blockDim.x);
        if (right) {
            const int c0 = max(a.meX, max(tx0, a.mbX));
            halfBorderColumns<NWX>(a, W, dstTile, a.borderX + 9 * (a.mbX + c0 - a.meX), tx0, c0, txEnd, th, threadIdx.x, blockDim.x);
        }
    }
}

// ---- variant 1: source rows read with ordinary global loads (any 4-byte aligned pitch) ----
template <int NG, int NWX, bool SYM, bool ENDHI>
__global__ void __launch_bounds__(256, 3) resizeHalfKernel(const __grid_constant__ HalfArgs a)
{
    __shared__ __align__(16) uint32_t W[kHalfMaxRows * kHalfRowWords];
    const int tx0 = blockIdx.x * kHalfTileW - a.tileShift;
    const int ty0 = blockIdx.y * a.tileRows;
    const int th = min(a.tileRows, a.DH - ty0);
    const uint8_t *__restrict__ src = a.src + (long long)blockIdx.z * a.srcFrameStride;
    uint8_t *__restrict__ dst = a.dst + (long long)blockIdx.z * a.dstFrameStride;
    const int xs0 = 2 * tx0 - 8;  // source column of W element 0 (multiple of 4)
    halfVertical<NG, false>(a, src, W, xs0, tx0, ty0, th);
    __syncthreads();
    halfHorizontal<NWX, SYM, ENDHI>(a, W, dst, xs0, tx0, ty0, th);
}

// ---- variant 2: the tile's source window (256 bytes x boxRows rows) is staged by one TMA
// (cp.async.bulk.tensor) per tile; rows/columns outside the image arrive as zeros ----
struct HalfTmaArgs {
    alignas(64) CUtensorMap tmap;  // 3-D: (x, y, frame), u8, box 256 x boxRows x 1
    HalfArgs h;
    int boxRows;
    int tileBytes;  // shared bytes reserved for the source window (boxRows + one slack group, 128-aligned)
};

__device__ __forceinline__ uint32_t smemAddr(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

// One tile per CTA.  (A persistent variant that prefetched the next tile's window during the
// horizontal pass was measured slower on B200: 4.03 ms vs 3.15 ms per 4096 1080p frames -- the
// hardware CTA scheduler overlaps tile start-up and balances the tail better.)
template <int NG, int NWX, bool SYM, bool ENDHI>
__global__ void __launch_bounds__(256, 3) resizeHalfTmaKernel(const __grid_constant__ HalfTmaArgs p)
{
    extern __shared__ __align__(128) uint8_t smemDyn[];
    uint8_t *tile = smemDyn;
    uint32_t *W = reinterpret_cast<uint32_t *>(smemDyn + p.tileBytes);
    __shared__ __align__(8) unsigned long long mbar;
    const HalfArgs &a = p.h;
    const uint32_t mbarAddr = smemAddr(&mbar);

    // the tile grid is shifted by 4 destination pixels so that the box starts on a 16-byte
    // boundary (2*tx0 - 8 = 240*i - 16): TMA requires that of the innermost coordinate
    const int tx0 = blockIdx.x * kHalfTileW - a.tileShift;
    const int ty0 = blockIdx.y * a.tileRows;
    const int th = min(a.tileRows, a.DH - ty0);
    uint8_t *__restrict__ dst = a.dst + (long long)blockIdx.z * a.dstFrameStride;
    const int xs0 = 2 * tx0 - 8;

    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbarAddr));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const uint32_t bytes = (uint32_t)p.boxRows * kHalfSrcRowBytes;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbarAddr), "r"(bytes) : "memory");
        const int y0 = 4 * ((ty0 >> 1) + a.qmin);  // first source row of the tile's first group
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
            ::"r"(smemAddr(tile)), "l"(reinterpret_cast<unsigned long long>(&p.tmap)), "r"(xs0), "r"(y0), "r"((int)blockIdx.z),
              "r"(mbarAddr)
            : "memory");
    }
    __syncthreads();  // the barrier is initialised (and armed) before anyone polls it
    {
        // every thread waits for the transaction bytes to land (phase 0)
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0;\n\tselp.u32 %0, 1, 0, q;\n\t}"
                : "=r"(done)
                : "r"(mbarAddr)
                : "memory");
        }
    }
    halfVertical<NG, true>(a, tile, W, xs0, tx0, ty0, th);
    __syncthreads();
    halfHorizontal<NWX, SYM, ENDHI>(a, W, dst, xs0, tx0, ty0, th);
}

template <int NG, int NWX, bool SYM, bool ENDHI>
cudaError_t launchHalfT(const HalfArgs &a, const CUtensorMap *tmap, int boxRows, cudaStream_t stream)
{
    const int tilesX = (a.DW + a.tileShift + kHalfTileW - 1) / kHalfTileW;
    const int tilesY = (a.DH + a.tileRows - 1) / a.tileRows;
    dim3 grid(tilesX, tilesY, a.nFrames);
    const int threads = a.tileRows > 32 ? 256 : 128;
    if (tmap) {
        static PerDeviceOnce attrSet;  // per instantiation; a benign race sets it twice at worst
        const int dev = currentDevice();
        if (!attrSet.done(dev)) {
            cudaError_t e = cudaFuncSetAttribute(resizeHalfTmaKernel<NG, NWX, SYM, ENDHI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 kHalfTileBytes + kHalfWBytes);
            if (e != cudaSuccess) return e;
            attrSet.set(dev);
        }
        HalfTmaArgs p;
        p.tmap = *tmap;
        p.h = a;
        p.boxRows = boxRows;
        p.tileBytes = ((boxRows + 4) * kHalfSrcRowBytes + 127) & ~127;
        const int smem = p.tileBytes + a.tileRows * kHalfRowWords * 4;
        resizeHalfTmaKernel<NG, NWX, SYM, ENDHI><<<grid, threads, smem, stream>>>(p);
    } else {
        resizeHalfKernel<NG, NWX, SYM, ENDHI><<<grid, threads, 0, stream>>>(a);
    }
    g_launches.fetch_add(1);
    return cudaGetLastError();
}

// Horizontal pass of eight adjacent destination pixels for the streaming variant: the pair words
// hold columns (2m-1, 2m), so pixel p reads exactly the NXH words n[kBo + p .. kBo + p + NXH - 1]
// and a symmetric table folds them into NXH / 2 pre-added words.
template <int NXH, bool SYM, bool SKIP0>
__device__ __forceinline__ uint2 halfGroupPixelsOdd(const HalfArgs &a, const uint32_t (&n)[16])
{
    constexpr int kBo = 5 - NXH / 2;
    int v[8];
    if (SYM) {
        constexpr int m = NXH / 2;
        uint32_t sw[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) sw[i] = (i >= kBo + m && i <= kBo + NXH - 1 + 7) ? prmt(n[i], n[i], 0x1032) : 0u;
        uint32_t sum[8][m];
#pragma unroll
        for (int j = 0; j < m; ++j)
#pragma unroll
            for (int p = 0; p < 8; ++p) sum[p][j] = n[kBo + p + j] + sw[kBo + p + NXH - 1 - j] + a.zero;  // see halfGroupPixels
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            int acc = a.accInit;
#pragma unroll
            for (int j = 0; j < m; ++j) acc = dp2a_lo_uu(sum[p][j], a.cwXo[j], acc);
            acc >>= 8;
#pragma unroll
            for (int j = 0; j < m; ++j)
                if (!(SKIP0 && j == 0)) acc = dp2a_hi_us(sum[p][j], a.cwXo[j], acc);
            v[p] = acc >> 12;
        }
    } else {
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            int acc = a.accInit;
#pragma unroll
            for (int j = 0; j < NXH; ++j) acc = dp2a_lo_uu(n[kBo + p + j], a.cwXo[j], acc);
            acc >>= 8;
#pragma unroll
            for (int j = 0; j < NXH; ++j) acc = dp2a_hi_us(n[kBo + p + j], a.cwXo[j], acc);
            v[p] = acc >> 12;
        }
    }
    uint2 o;
    o.x = packSatU8(v[1], v[0], packSatU8(v[3], v[2], 0u));
    o.y = packSatU8(v[5], v[4], packSatU8(v[7], v[6], 0u));
    return o;
}

// ---- variant 3: warp-autonomous streaming ----
// A warp (one per CTA) owns one 120-pixel column strip of one frame and walks down a band of
// destination row pairs on its own: no CTA-wide barrier, no per-tile prologue, and the register
// ring of transposed source groups lives for the whole band instead of being refilled per tile.
//   source      the 272 bytes of a source row that cover the strip's 256-byte window travel
//               global -> shared as 17 aligned 16-byte cp.async.cg chunks (lanes 0..16; L1 is bypassed,
//               chunks outside the image are zero-filled without a read).  Four rows form a group;
//               the FIFO holds two turns of groups, so the loads of the next row pairs are always
//               in flight and cost no registers.  (A variant fed by one cp.async.bulk.tensor per
//               group was measured 5 % slower -- 2.64 vs 2.52 ms for cfg4 -- and dropped.)
//   vertical    a lane owns 8 adjacent source columns: 2 x (4x4 byte transposes) per group,
//               8 x NG dp4a per destination row, one 16-byte store per row into the warp's W rows;
//               the loop runs in turns of NG row pairs (compile-time ring, FIFO and W positions);
//   horizontal  after one __syncwarp: lane = (row of the pair, 8-pixel group), 30 of 32 lanes
//               busy, pre-added dp2a planes on odd-aligned pair words (halfGroupPixelsOdd).
//   borders     border rows: masked coefficient words + truncating division, in the turns that
//               touch them.  Border columns: the few W chunks they read are parked in a side
//               buffer and recomputed for up to 16 rows at a time, a lane per row.
// Needs 16-byte aligned source rows and a source width that is a multiple of 8.
template <bool B>
struct BoolTag {
    static constexpr bool value = B;
};

// Launch shape of the streaming kernel: one warp per CTA, 16 CTAs per SM.
#ifndef IQO_STREAM_WARPS
#define IQO_STREAM_WARPS 1
#endif
#ifndef IQO_STREAM_MINB
#define IQO_STREAM_MINB 16
#endif
// (measured: the 12-tap instantiations gain 2-3 % with 18 CTAs per SM, the 8-tap ones lose 1-2 %)
#define IQO_STREAM_BOUNDS __launch_bounds__(32 * IQO_STREAM_WARPS, NXH == 6 ? IQO_STREAM_MINB + 2 : IQO_STREAM_MINB)
constexpr int kStreamWarps = IQO_STREAM_WARPS;  // strips (warps) per CTA
#ifndef IQO_STREAM_SIDE_ROWS
#define IQO_STREAM_SIDE_ROWS 16
#endif
constexpr int kStreamSideRows = IQO_STREAM_SIDE_ROWS;   // destination rows parked before the border columns are flushed
constexpr int kStreamSideWords = 32;  // per parked row: W chunks 0..3 (left) and rc0..rc0+3 (right)

// Border columns [c0, c1) of `nrows` (<= 32) parked rows: a lane per row, the columns in a loop so
// that their coefficient words are uniform loads; `wordOff` maps (column + pair word) to a side word.
template <int NXH>
__device__ __noinline__ void streamBorderColumns(const HalfArgs &a, const uint32_t *side, uint8_t *dstRow0, const int32_t *bx,
                                                 int wordOff, int c0, int c1, int nrows, int lane)
{
    if (lane >= nrows) return;
    const uint32_t *wr = side + lane * kStreamSideWords + wordOff;
    uint8_t *out = dstRow0 + (long long)lane * a.dstPitch;
    for (int d = c0; d < c1; ++d, bx += 8) {
        int lo = __ldg(bx + 7), hi = 0;
#pragma unroll
        for (int i = 0; i < NXH; ++i) {
            const uint32_t word = wr[d + i];
            const uint32_t cw = (uint32_t)__ldg(bx + i);
            lo = dp2a_lo_uu(word, cw, lo);
            hi = dp2a_hi_us(word, cw, hi);
        }
        const int v = (int)(short)((lo + (hi << 8)) / __ldg(bx + 6));
        out[d] = (uint8_t)min(max(v, 0), 255);
    }
}

// shared-memory accesses by 32-bit shared address (keeps the address arithmetic to one add per step)
template <int OFF>
__device__ __forceinline__ uint4 ldsV4(uint32_t addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4+%5];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr), "n"(OFF) : "memory");
    return v;
}

template <int OFF>
__device__ __forceinline__ uint2 ldsV2(uint32_t addr)
{
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2+%3];" : "=r"(v.x), "=r"(v.y) : "r"(addr), "n"(OFF) : "memory");
    return v;
}

template <int OFF>
__device__ __forceinline__ void stsV4(uint32_t addr, uint4 v)
{
    asm volatile("st.shared.v4.u32 [%0+%1], {%2, %3, %4, %5};" ::"r"(addr), "n"(OFF), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

constexpr int kStreamRowBytes = 4 * kHalfRowWords + 16;  // W row stride: the 16-byte skew keeps the two rows of a pair on different banks
constexpr int kStreamSrcRowBytes = 272;                  // staged bytes of a source row: 17 aligned 16-byte chunks
constexpr int kStreamSlotBytes = 4 * kStreamSrcRowBytes; // one source group of a warp: 4 rows
#ifndef IQO_STREAM_FIFO_TURNS
#define IQO_STREAM_FIFO_TURNS 2
#endif
// shared bytes of one warp: W rows of one turn, the source FIFO, the parked border chunks
__host__ __device__ constexpr int streamWarpBytes(int NG)
{
    return 2 * NG * kStreamRowBytes + IQO_STREAM_FIFO_TURNS * NG * kStreamSlotBytes + kStreamSideRows * kStreamSideWords * 4;
}

// TMA-fed variant: the source FIFO is two boxes of one turn each (4 NG rows x 272 bytes, one
// cp.async.bulk.tensor per turn issued by lane 0, completion on an mbarrier per box) instead of
// 4 cp.async per lane and group.  The tensor map views the frames as 16-bit pairs (a box dimension
// holds at most 256 elements): (x / 2, y, frame), box 136 x 4 NG x 1; rows and columns outside the
// image arrive as zeros, which only ever meet zero coefficients.
struct HalfStreamTmaArgs {
    alignas(64) CUtensorMap tmap;
    HalfArgs h;
};
template <bool TMA>
struct StreamParam {
    typedef HalfArgs type;
    static __device__ __forceinline__ const HalfArgs &args(const HalfArgs &p) { return p; }
};
template <>
struct StreamParam<true> {
    typedef HalfStreamTmaArgs type;
    static __device__ __forceinline__ const HalfArgs &args(const HalfStreamTmaArgs &p) { return p.h; }
};
__host__ __device__ constexpr int streamTmaBoxBytes(int NG)
{
    return (4 * NG * kStreamSrcRowBytes + 127) / 128 * 128;  // TMA destinations are 128-byte aligned
}
__host__ __device__ constexpr int streamTmaWarpBytes(int NG)
{
    return 2 * streamTmaBoxBytes(NG) + 2 * NG * kStreamRowBytes + kStreamSideRows * kStreamSideWords * 4 + 16;
}

// The loop runs in *turns* of NG row pairs (one revolution of the register ring), so that ring
// positions, FIFO slots and W rows are compile-time constants and the per-pair bookkeeping is
// paid once per turn: NG vertical passes, one __syncwarp, then the horizontal pass of the 2 NG rows.
template <int NG, int NXH, bool SYM, bool SKIP0, int Z, bool TMA>
__global__ void IQO_STREAM_BOUNDS resizeHalfStreamKernel(const __grid_constant__ typename StreamParam<TMA>::type prm)
{
    extern __shared__ __align__(128) uint8_t streamSmem[];
    const HalfArgs &a = StreamParam<TMA>::args(prm);
    constexpr int kBase = 5 - NXH / 2;  // pair word of taps 0, 1 of pixel 0
    constexpr int kFifo = IQO_STREAM_FIFO_TURNS * NG;      // groups in a lane's FIFO
    constexpr int kWBuf = 2 * NG * kStreamRowBytes;        // W rows of one turn
    constexpr int kBox = streamTmaBoxBytes(NG);            // TMA: one box = the NG source groups of a turn
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tx0 = (blockIdx.x * kStreamWarps + warp) * kHalfTileW;
    if (tx0 >= a.DW) return;
    const int pairs = (a.DH + 1) >> 1;
    const int k0 = blockIdx.y * a.bandPairs;
    const int k1 = min(k0 + a.bandPairs, pairs);
    if (k0 >= k1) return;
    const uint8_t *__restrict__ src = a.src + (long long)blockIdx.z * a.srcFrameStride;
    uint8_t *__restrict__ dst = a.dst + (long long)blockIdx.z * a.dstFrameStride;
    const int xs0 = 2 * tx0 - 8;  // first source column of lane 0 (multiple of 8); W element i is column xs0 + i - 1
    const int txEnd = min(tx0 + kHalfTileW, a.DW);
    // shared bytes of a warp: [W rows | FIFO | side] (cp.async) or [box 0 | box 1 | W rows | side | 2 mbarriers] (TMA)
    uint8_t *warpSmem = streamSmem + warp * (TMA ? streamTmaWarpBytes(NG) : streamWarpBytes(NG));
    const uint32_t fifoBase = smemAddr(warpSmem) + (TMA ? 0 : kWBuf);
    const uint32_t wBase = smemAddr(warpSmem) + (TMA ? 2 * kBox : 0);
    uint32_t *side = reinterpret_cast<uint32_t *>(warpSmem + (TMA ? 2 * kBox + kWBuf : kWBuf + kFifo * kStreamSlotBytes));
    const uint32_t mbarBase = smemAddr(side) + kStreamSideRows * kStreamSideWords * 4;  // TMA only

    // source role: lanes 0..16 copy the 16-byte aligned chunks [xs0 - 8 + 16 lane, +16) of a source row (L1 bypassed);
    // chunks outside the image are zero-filled without a read, one that straddles the right edge reads 8 bytes
    const int ccol = xs0 - 8 + 16 * lane;
    const bool copyLane = lane < 17;
    const int csize = (ccol < 0 || ccol >= a.SW) ? 0 : min(16, a.SW - ccol);
    const uint8_t *base = src + (csize ? ccol : 0);
    const long long pitch = a.srcPitch;
    const int SHm1 = a.SH - 1;
    const int B = a.workBias;
    const uint32_t wst = wBase + 16 * lane;  // this lane's chunk of W row 0
    // horizontal role: within every group of eight lanes four read row 0 and four row 1, so that
    // the 16-byte loads of a quarter warp fall on eight different bank groups
    const int hrow = (lane >> 2) & 1, hl = (lane & 3) | ((lane >> 3) << 2);
    const int d0 = tx0 + 8 * hl;
    // 0: nothing to store, 1: one 8-byte store, 2: one 4-byte store (group cut by the image edge), 3: bytes
    int hmode = 0;
    if (hl < 15 && d0 < txEnd) hmode = !a.dstVec ? 3 : d0 + 8 <= a.DW ? 1 : d0 + 4 == a.DW ? 2 : 3;
    const uint32_t wld = wBase + hrow * kStreamRowBytes + 32 * hl;
    uint8_t *outp = dst + (long long)(2 * k0 + hrow) * a.dstPitch + d0;
    const long long ostep = 2 * a.dstPitch;
    // border columns: the lanes whose W chunks they read park them in the side buffer
    // (destination rows are at least 32 pixels wide here, so no lane serves both sides)
    const bool left = tx0 < a.mbX, right = txEnd > a.meX;
    const int rcol0 = max(a.meX, max(tx0, a.mbX));    // first right border column of this strip
    const int rc0 = (rcol0 - tx0 + kBase) >> 2;       // first W chunk the right border columns read
    const bool parkL = left && lane < 4, parkR = right && (unsigned)(lane - rc0) < 4u;
    const bool park = parkL || parkR;
    const bool edgeStrip = left || right;
    const uint32_t sideLane = smemAddr(side) + (parkL ? 16 * lane : 64 + 16 * (lane - rc0));
    uint32_t sideAddr = sideLane;
    int sideRows = 0;

    // turns starting at pairs [kIntB, kIntE] (whole turn inside) touch no border row and request only rows inside the image
    // (TMA: rows outside the image arrive as zeros, only the border rows matter)
    const int gAhead = a.qmin + NG - 1 + kFifo - 1;  // the step of pair k requests group k + gAhead
    const int kIntB = TMA ? (a.mbY + 1) >> 1 : max((a.mbY + 1) >> 1, -gAhead);
    const int kIntE = TMA ? min(a.meY >> 1, k1) : min(min(a.meY >> 1, ((a.SH - a.delta) >> 2) - gAhead), k1);

    int g = k0 + a.qmin;  // next source group to request
    const uint8_t *gp = base + (long long)(4 * g + a.delta) * pitch;
    // FIFO slot of the n-th group of the band is n mod kFifo.  With a FIFO two turns deep the slots
    // of a turn alternate between the halves fifoCur / fifoOth.
    uint32_t fifoCur = fifoBase, fifoOth = fifoBase + (IQO_STREAM_FIFO_TURNS > 1 ? NG * kStreamSlotBytes : 0);
    auto issue = [&](auto edgeTag, const uint32_t slot) {  // request the four rows of group g into `slot`
        const uint32_t sa = slot + 16 * lane;
        if (copyLane) {
            if (decltype(edgeTag)::value) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int row = min(max(4 * g + a.delta + j, 0), SHm1);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sa + kStreamSrcRowBytes * j),
                                 "l"(base + (long long)row * pitch), "r"(csize)
                                 : "memory");
                }
            } else {
                const uint8_t *p1 = gp + pitch;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sa), "l"(gp), "r"(csize) : "memory");
                asm volatile("cp.async.cg.shared.global [%0+272], [%1], 16, %2;" ::"r"(sa), "l"(p1), "r"(csize) : "memory");
                asm volatile("cp.async.cg.shared.global [%0+544], [%1], 16, %2;" ::"r"(sa), "l"(gp + 2 * pitch), "r"(csize) : "memory");
                asm volatile("cp.async.cg.shared.global [%0+816], [%1], 16, %2;" ::"r"(sa), "l"(p1 + 2 * pitch), "r"(csize) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        gp += 4 * pitch;
        ++g;
    };
    // TMA: box j of the band holds the source groups NG j - 1 ... NG j + NG - 2 (group 0 = k0 + qmin), i.e. box 0 the
    // groups of the initial register ring and box t + 1 the groups turn t consumes; it lives in slot j & 1 and completes
    // phase (j >> 1) & 1 of that slot's mbarrier.
    const int boxX = tx0 - 8;                                      // u16 elements: byte column xs0 - 8
    const int boxY0 = 4 * (k0 + a.qmin - 1) + a.delta;             // first row of box 0
    auto issueBox = [&](const int j) {
        if (lane == 0) {
            const uint32_t bar = mbarBase + 8 * (j & 1);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "n"(4 * NG * kStreamSrcRowBytes) : "memory");
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                ::"r"(fifoBase + (j & 1) * kBox), "l"(reinterpret_cast<unsigned long long>(&prm)), "r"(boxX),
                  "r"(boxY0 + 4 * NG * j), "r"((int)blockIdx.z), "r"(bar)
                : "memory");
        }
    };
    auto waitBox = [&](const int j) {
        asm volatile(
            "{\n\t.reg .pred q;\n\tIQO_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%0], %1;\n\t@!q bra IQO_WAIT;\n\t}"
            ::"r"(mbarBase + 8 * (j & 1)), "r"((j >> 1) & 1)
            : "memory");
    };
    // oldest requested group (slot `ra`) -> transposed columns: .x/.y/.z/.w = four vertical bytes of column 0..3 (A) and 4..7 (B)
    auto consume = [&](const uint32_t slot, uint4 &ca, uint4 &cb) {
        if (!TMA) {
            asm volatile("cp.async.wait_group %0;" ::"n"(kFifo - 2) : "memory");
            __syncwarp();  // the chunks were copied by other lanes; everybody has also read the slot refilled next
        }
        const uint32_t ra = slot + 8 + 8 * lane;
        const uint2 r0 = ldsV2<0>(ra), r1 = ldsV2<kStreamSrcRowBytes>(ra), r2 = ldsV2<2 * kStreamSrcRowBytes>(ra),
                    r3 = ldsV2<3 * kStreamSrcRowBytes>(ra);
        const uint32_t t0 = prmt(r0.x, r1.x, 0x5140), t1 = prmt(r0.x, r1.x, 0x7362);
        const uint32_t t2 = prmt(r2.x, r3.x, 0x5140), t3 = prmt(r2.x, r3.x, 0x7362);
        ca.x = prmt(t0, t2, 0x5410);
        ca.y = prmt(t0, t2, 0x7632);
        ca.z = prmt(t1, t3, 0x5410);
        ca.w = prmt(t1, t3, 0x7632);
        const uint32_t u0 = prmt(r0.y, r1.y, 0x5140), u1 = prmt(r0.y, r1.y, 0x7362);
        const uint32_t u2 = prmt(r2.y, r3.y, 0x5140), u3 = prmt(r2.y, r3.y, 0x7362);
        cb.x = prmt(u0, u2, 0x5410);
        cb.y = prmt(u0, u2, 0x7632);
        cb.z = prmt(u1, u3, 0x5410);
        cb.w = prmt(u1, u3, 0x7632);
    };
    // Group n of the band (n = 0 is group k0 + qmin) lives in slot n mod kFifo; the request of group
    // n + kFifo - 1 goes out right after group n is read and lands in the slot read one step earlier.
    // Before the first turn: groups 0 .. NG-2 are in the register ring, groups up to NG + kFifo - 3 requested.
    uint4 winA[NG], winB[NG];
    int boxJ = 1;  // TMA: box of the running turn
    if (TMA) {
        if (lane == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbarBase));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbarBase + 8));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        issueBox(0);
        issueBox(1);
        waitBox(0);
#pragma unroll
        for (int j = 0; j < NG - 1; ++j) consume(fifoBase + (j + 1) * kStreamSlotBytes, winA[j], winB[j]);
        __syncwarp();  // box 0 has been read by every lane: its slot may be refilled
    } else {
#pragma unroll
        for (int j = 0; j < kFifo - 1; ++j) issue(BoolTag<true>(), fifoBase + j * kStreamSlotBytes);
#pragma unroll
        for (int j = 0; j < NG - 1; ++j) {
            consume(fifoBase + j * kStreamSlotBytes, winA[j], winB[j]);
            issue(BoolTag<true>(), fifoBase + ((j + kFifo - 1) % kFifo) * kStreamSlotBytes);
        }
    }

    // border columns of the parked rows; `yEnd` is the row after the last parked one
    auto flush = [&](const int yEnd) {
        __syncwarp();
        const int y0 = yEnd - sideRows;
        const int nrows = min(sideRows, a.DH - y0);
        uint8_t *drow = dst + (long long)y0 * a.dstPitch;
        if (nrows > 0) {
            if (left) streamBorderColumns<NXH>(a, side, drow, a.borderX + 8 * tx0, kBase - tx0, tx0, min(a.mbX, txEnd), nrows, lane);
            if (right)
                streamBorderColumns<NXH>(a, side, drow, a.borderX + 8 * (a.mbX + rcol0 - a.meX), kBase - tx0 - 4 * rc0 + 16, rcol0,
                                         txEnd, nrows, lane);
        }
        __syncwarp();
        sideRows = 0;
        sideAddr = sideLane;
    };

    // vertical pass of the s-th row pair (k) of a turn; s is also the position of the register ring
    auto vertical = [&](auto edgeTag, auto stepTag, const int k) {
        constexpr bool EDGE = decltype(edgeTag)::value;
        constexpr int s = decltype(stepTag)::value;
        // step s of a turn reads group n = NG t + s + NG - 1: slot NG-1 of the current half for s == 0, else slot s-1 of
        // the other half; the request goes into the slot read one step earlier
        const uint32_t ra = s == 0 ? fifoCur + (NG - 1) * kStreamSlotBytes : fifoOth + (s - 1) * kStreamSlotBytes;
        const uint32_t sa = s == 0 ? fifoCur + (NG - 2) * kStreamSlotBytes
                          : s == 1 ? fifoCur + (NG - 1) * kStreamSlotBytes : fifoOth + (s - 2) * kStreamSlotBytes;
        if (TMA) {
            consume(fifoBase + (boxJ & 1) * kBox + s * kStreamSlotBytes, winA[(s + NG - 1) % NG], winB[(s + NG - 1) % NG]);
        } else {
            consume(ra, winA[(s + NG - 1) % NG], winB[(s + NG - 1) % NG]);
            issue(edgeTag, sa);
        }
#pragma unroll
        for (int par = 0; par < 2; ++par) {
            uint32_t c0 = a.cwY[par][0], c1 = NG > 1 ? a.cwY[par][1] : 0u, c2 = NG > 2 ? a.cwY[par][2] : 0u;
            int deno = 0;
            uint32_t magic = 0;
            if (EDGE) {
                const int y = 2 * k + par;
                if (y < a.DH && (y < a.mbY || y >= a.meY)) {
                    const int row = __ldg(a.rowY + y);
                    deno = __ldg(a.denoY + row);
                    magic = __ldg(a.magicY + row);
                    c0 = __ldg(a.borderY + row * 3);
                    c1 = __ldg(a.borderY + row * 3 + 1);
                    c2 = __ldg(a.borderY + row * 3 + 2);
                }
            }
            const int init = (EDGE && deno) ? 0 : B;
            int v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = init;
#pragma unroll
            for (int t = 0; t < NG; ++t) {
                if (!EDGE && ((Z >> (par * 3 + t)) & 1)) continue;  // zero word of the main phase
                const uint4 qa = winA[(s + t) % NG], qb = winB[(s + t) % NG];
                const uint32_t c = t == 0 ? c0 : t == 1 ? c1 : c2;
                v[0] = dp4a_us(qa.x, c, v[0]);
                v[1] = dp4a_us(qa.y, c, v[1]);
                v[2] = dp4a_us(qa.z, c, v[2]);
                v[3] = dp4a_us(qa.w, c, v[3]);
                v[4] = dp4a_us(qb.x, c, v[4]);
                v[5] = dp4a_us(qb.y, c, v[5]);
                v[6] = dp4a_us(qb.z, c, v[6]);
                v[7] = dp4a_us(qb.w, c, v[7]);
            }
            if (EDGE && deno) {
                // resizeYborder: see halfVerticalStrip
                auto bdiv = [&](int x) -> int {
                    const int n = (int)(short)x * 64;
                    const uint32_t m = (uint32_t)abs(n);
                    const int q = magic ? (int)__umulhi(m, magic) : (int)m;
                    return (int)(short)(n < 0 ? -q : q) + B;
                };
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = bdiv(v[i]);
            }
            // pair words (column 2m-1, column 2m): the lane's last column travels to its right neighbour
            const uint32_t prev = __shfl_up_sync(0xffffffffu, (uint32_t)v[7], 1);
            uint4 o;
            o.x = prmt(prev, (uint32_t)v[0], 0x5410);
            o.y = prmt((uint32_t)v[1], (uint32_t)v[2], 0x5410);
            o.z = prmt((uint32_t)v[3], (uint32_t)v[4], 0x5410);
            o.w = prmt((uint32_t)v[5], (uint32_t)v[6], 0x5410);
            if (par == 0) {
                stsV4<2 * s * kStreamRowBytes>(wst, o);
                if (park) stsV4<2 * s * 4 * kStreamSideWords>(sideAddr, o);
            } else {
                stsV4<(2 * s + 1) * kStreamRowBytes>(wst, o);
                if (park) stsV4<(2 * s + 1) * 4 * kStreamSideWords>(sideAddr, o);
            }
        }
    };

    for (int k = k0; k < k1; k += NG) {
        if (TMA) {
            if (k + NG < k1) issueBox(boxJ + 1);  // next turn's groups, into the slot the previous turn has finished with
            waitBox(boxJ);
        }
        if (k >= kIntB && k + NG <= kIntE) {
            vertical(BoolTag<false>(), std::integral_constant<int, 0>(), k);
            if (NG > 1) vertical(BoolTag<false>(), std::integral_constant<int, 1 % NG>(), k + 1);
            if (NG > 2) vertical(BoolTag<false>(), std::integral_constant<int, 2 % NG>(), k + 2);
        } else {
            vertical(BoolTag<true>(), std::integral_constant<int, 0>(), k);
            if (NG > 1) vertical(BoolTag<true>(), std::integral_constant<int, 1 % NG>(), k + 1);
            if (NG > 2) vertical(BoolTag<true>(), std::integral_constant<int, 2 % NG>(), k + 2);
        }
        __syncwarp();
        // horizontal pass of the turn's row pairs that belong to the band: lane = (row of the pair, 8-pixel group)
        const int np = min(NG, k1 - k);
        if (hmode == 1 && k + NG <= k1 && 2 * (k + NG) <= a.DH) {
            // whole turn inside the band and the image, 8-byte stores: no per-row checks
#pragma unroll
            for (int i = 0; i < NG; ++i) {
                uint32_t n[16];
                const uint32_t wl = wld + i * 2 * kStreamRowBytes;
                const uint4 q0 = ldsV4<0>(wl), q1 = ldsV4<16>(wl), q2 = ldsV4<32>(wl), q3 = ldsV4<48>(wl);
                n[0] = q0.x, n[1] = q0.y, n[2] = q0.z, n[3] = q0.w;
                n[4] = q1.x, n[5] = q1.y, n[6] = q1.z, n[7] = q1.w;
                n[8] = q2.x, n[9] = q2.y, n[10] = q2.z, n[11] = q2.w;
                n[12] = q3.x, n[13] = q3.y, n[14] = q3.z, n[15] = q3.w;
                const uint2 o = halfGroupPixelsOdd<NXH, SYM, SKIP0>(a, n);
                *reinterpret_cast<uint2 *>(outp + i * ostep) = o;
            }
        } else if (hmode != 0) {
            uint32_t wl = wld;
            uint8_t *op = outp;
#pragma unroll
            for (int i = 0; i < NG; ++i, wl += 2 * kStreamRowBytes, op += ostep) {
                if (i >= np) break;
                if (2 * (k + i) + hrow >= a.DH) break;
                uint32_t n[16];
                const uint4 q0 = ldsV4<0>(wl), q1 = ldsV4<16>(wl), q2 = ldsV4<32>(wl), q3 = ldsV4<48>(wl);
                n[0] = q0.x, n[1] = q0.y, n[2] = q0.z, n[3] = q0.w;
                n[4] = q1.x, n[5] = q1.y, n[6] = q1.z, n[7] = q1.w;
                n[8] = q2.x, n[9] = q2.y, n[10] = q2.z, n[11] = q2.w;
                n[12] = q3.x, n[13] = q3.y, n[14] = q3.z, n[15] = q3.w;
                const uint2 o = halfGroupPixelsOdd<NXH, SYM, SKIP0>(a, n);
                if (hmode == 1)
                    *reinterpret_cast<uint2 *>(op) = o;
                else if (hmode == 2)
                    *reinterpret_cast<uint32_t *>(op) = o.x;
                else
                    halfStoreBytes(op, o, 0, min(8, a.DW - d0));
            }
        }
        outp += NG * ostep;
        __syncwarp();  // the W rows are free again
        if (TMA) {
            ++boxJ;
        } else if (IQO_STREAM_FIFO_TURNS > 1) {
            const uint32_t t = fifoCur;
            fifoCur = fifoOth;
            fifoOth = t;
        }
        if (edgeStrip) {
            sideRows += 2 * np;
            sideAddr += 2 * NG * 4 * kStreamSideWords;
            if (sideRows + 2 * NG > kStreamSideRows || k + NG >= k1) flush(2 * (k + np));
        }
    }
    if (!TMA) asm volatile("cp.async.wait_all;" ::: "memory");
}

template <int NG, int NXH, bool SYM, bool SKIP0, int Z>
cudaError_t launchHalfStreamT(const HalfArgs &a, const CUtensorMap *tmap, cudaStream_t stream)
{
    const int strips = (a.DW + kHalfTileW - 1) / kHalfTileW;
    const int pairs = (a.DH + 1) / 2;
    dim3 grid((strips + kStreamWarps - 1) / kStreamWarps, (pairs + a.bandPairs - 1) / a.bandPairs, a.nFrames);
    const int dev = currentDevice();
    if (tmap) {
        constexpr int smem = kStreamWarps * streamTmaWarpBytes(NG);
        static PerDeviceOnce attrSet;  // per instantiation; a benign race sets it twice at worst
        if (!attrSet.done(dev)) {
            cudaError_t e = cudaFuncSetAttribute(resizeHalfStreamKernel<NG, NXH, SYM, SKIP0, Z, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return e;
            attrSet.set(dev);
        }
        HalfStreamTmaArgs p;
        p.tmap = *tmap;
        p.h = a;
        resizeHalfStreamKernel<NG, NXH, SYM, SKIP0, Z, true><<<grid, 32 * kStreamWarps, smem, stream>>>(p);
    } else {
        constexpr int smem = kStreamWarps * streamWarpBytes(NG);
        static PerDeviceOnce attrSet;
        if (!attrSet.done(dev)) {
            cudaError_t e = cudaFuncSetAttribute(resizeHalfStreamKernel<NG, NXH, SYM, SKIP0, Z, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return e;
            attrSet.set(dev);
        }
        resizeHalfStreamKernel<NG, NXH, SYM, SKIP0, Z, false><<<grid, 32 * kStreamWarps, smem, stream>>>(a);
    }
    g_launches.fetch_add(1);
    return cudaGetLastError();
}


// ---------------------------------------------------------------------------------------
// General packed kernel: any kind, ratio, phase count (plan.hpp PackedPlan).
//
//  vertical pass   a thread owns 4 adjacent source columns (one aligned 32-bit word per source
//                  row) of one destination row.  The four bytes are spread into two registers of
//                  two 16-bit lanes (2 PRMT) and each tap is ONE multiply-add per register:
//                  (b1 * 2^16 + b0) * c accumulates both columns at once.  The planner proves that
//                  every partial sum, after a bias, stays inside [0, 65535], so no carry or
//                  borrow ever crosses the lanes and the arithmetic equals two independent
//                  16-bit sums -- which is exactly the packed layout the horizontal pass wants.
//                  Out-of-image taps were trimmed by the planner: no index is clamped.
//  horizontal pass one pixel per lane: dp2a over the 16-bit pairs of its window against
//                  planner-packed coefficient words (low / high byte planes, one variant per
//                  parity of the window start), one shift, clamp, byte store (a warp writes 32
//                  consecutive bytes).  Lanczos border columns divide instead of shifting.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ int dp2a_hi_uu(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

constexpr int kPackedTileW = 128;  // destination columns per tile = threads per row half
constexpr int kPackedTileH = 32;   // destination rows per tile (two halves of 16)
constexpr int kPackedMaxTapsY = 64;

template <bool SIGNED, int NPT>
__global__ void __launch_bounds__(256) resizePackedKernel(const __grid_constant__ PackedArgs a)
{
    extern __shared__ __align__(16) uint32_t Wp[];  // [tileH][wordsPerRow] packed 16-bit pairs
    __shared__ int sFy[kPackedTileH], sNt[kPackedTileH], sDeno[kPackedTileH];
    __shared__ uint32_t sMagic[kPackedTileH];
    int32_t *sCoef = reinterpret_cast<int32_t *>(Wp + a.tileH * a.wordsPerRow);  // [tileH][ntMax]

    const int tx0 = blockIdx.x * kPackedTileW;
    const int ty0 = blockIdx.y * kPackedTileH;
    const int tw = min(kPackedTileW, a.DW - tx0);
    const int th = min(kPackedTileH, a.dstRows - ty0);
    const uint8_t *__restrict__ src = a.src + (long long)blockIdx.z * a.srcFrameStride;
    uint8_t *__restrict__ dst = a.dst + (long long)blockIdx.z * a.dstFrameStride;
    const int B = a.workBias;

    // per-row metadata and coefficient slices of this tile -> shared memory
    if (threadIdx.x < th) {
        const int y = a.dstRow0 + ty0 + threadIdx.x;
        sFy[threadIdx.x] = __ldg(a.firstY + y) - a.srcRow0;
        sNt[threadIdx.x] = __ldg(a.ntapY + y);
        int deno = 0;
        uint32_t magic = 0;
        if (SIGNED) {
            const int row = __ldg(a.rowY + y);
            deno = __ldg(a.denoY + row);
            magic = __ldg(a.magicY + row);
        }
        sDeno[threadIdx.x] = deno;
        sMagic[threadIdx.x] = magic;
    }
    for (int idx = threadIdx.x; idx < th * a.ntMax; idx += blockDim.x) {
        const int r = idx / a.ntMax, i = idx - r * a.ntMax;
        const int y = a.dstRow0 + ty0 + r;
        sCoef[idx] = (i < __ldg(a.ntapY + y)) ? __ldg(a.coefY + __ldg(a.coefOffY + y) + i) : 0;
    }
    // source column window of the tile (firstX is non-decreasing and >= 0), in 8-column units
    const int x0 = __ldg(&a.recX[tx0].x) & ~7;
    const int x1 = min(__ldg(&a.recX[tx0 + tw - 1].x) + a.NX - 1, a.SW - 1);
    const int ncd = ((x1 - x0) >> 3) + 1;
    const int swWords = (a.SW + 3) >> 2;  // 32-bit words per source row that hold image bytes
    __syncthreads();

    // ================= vertical pass: item = (row, 8 adjacent source columns) =================
    {
        // floor(item / ncd) as a multiply-high by ceil(2^32 / ncd): exact for item < 2^32 / ncd
        // (a 16-bit reciprocal is not: wide source windows of strong down-sampling broke it)
        const uint32_t rcp = (uint32_t)((0x100000000ull + (uint32_t)ncd - 1) / (uint32_t)ncd);
        for (int item = threadIdx.x; item < th * ncd; item += blockDim.x) {
            const int r = ncd == 1 ? item : (int)__umulhi((uint32_t)item, rcp);
            const int cd = item - r * ncd;
            const int col = x0 + 8 * cd;
            const bool ok1 = ((col >> 2) + 1) < swWords;  // the second word still holds image bytes
            const uint8_t *ptr = src + (long long)sFy[r] * a.srcPitch + col;
            const int nt = sNt[r];
            const int deno = sDeno[r];
            const int32_t *cs = sCoef + r * a.ntMax;
            const uint32_t init = deno ? 0u : ((uint32_t)B | ((uint32_t)B << 16));
            uint32_t a0 = init, a1 = init, a2 = init, a3 = init;
            for (int i = 0; i < nt; ++i) {
                const uint32_t w0 = __ldg(reinterpret_cast<const uint32_t *>(ptr));
                const uint32_t w1 = ok1 ? __ldg(reinterpret_cast<const uint32_t *>(ptr + 4)) : 0u;
                ptr += a.srcPitch;
                const uint32_t c = (uint32_t)cs[i];
                a0 += prmt(w0, 0u, 0x4140) * c;  // 16-bit lanes (col 0, col 1)
                a1 += prmt(w0, 0u, 0x4342) * c;  // (col 2, col 3)
                a2 += prmt(w1, 0u, 0x4140) * c;
                a3 += prmt(w1, 0u, 0x4342) * c;
            }
            if (SIGNED && deno) {
                // resizeYborder: the lanes are plain int16 sums here (no bias was added; a negative
                // low lane borrowed one from the high lane); work = int16(nume * 64 / deno), C
                // division as a multiply-high (|nume * 64| <= 2^21, 1 <= deno <= 255)
                const uint32_t magic = sMagic[r];
                auto bdiv = [&](uint32_t packed) -> uint32_t {
                    const uint32_t lo = packed & 0xffffu;
                    const uint32_t hi = (packed >> 16) + ((lo & 0x8000u) ? 1u : 0u);
                    uint32_t out = 0;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int n = (int)(short)(h ? hi : lo) * 64;
                        const uint32_t m = (uint32_t)abs(n);
                        const int qv = magic ? (int)__umulhi(m, magic) : (int)m;
                        const int w = (int)(short)(n < 0 ? -qv : qv) + B;
                        out |= ((uint32_t)w & 0xffffu) << (16 * h);
                    }
                    return out;
                };
                a0 = bdiv(a0);
                a1 = bdiv(a1);
                a2 = bdiv(a2);
                a3 = bdiv(a3);
            }
            *reinterpret_cast<uint4 *>(Wp + r * a.wordsPerRow + 4 * cd) = make_uint4(a0, a1, a2, a3);
        }
    }
    __syncthreads();

    // ================= horizontal pass: thread = one destination column, half of the rows ========
    {
        const int dx = threadIdx.x & (kPackedTileW - 1);
        const int rh = threadIdx.x >> 7;  // 0 / 1
        if (dx < tw) {
            const int4 rec = __ldg(a.recX + tx0 + dx);  // {first column, coefficient word offset, accumulator init, divisor}
            uint32_t cw[NPT];
#pragma unroll
            for (int j = 0; j < NPT; ++j) cw[j] = __ldg(a.cwX + rec.y + j);
            const uint32_t *wp = Wp + ((rec.x - x0) >> 1) + rh * (kPackedTileH / 2) * a.wordsPerRow;
            uint8_t *out = dst + (long long)(ty0 + rh * (kPackedTileH / 2)) * a.dstPitch + tx0 + dx;
            const int rows = min(kPackedTileH / 2, th - rh * (kPackedTileH / 2));
            for (int r = 0; r < rows; ++r, wp += a.wordsPerRow, out += a.dstPitch) {
                int lo = rec.z, hi = 0;
#pragma unroll
                for (int j = 0; j < NPT; ++j) {
                    const uint32_t w = wp[j];
                    lo = dp2a_lo_uu(w, cw[j], lo);
                    hi = SIGNED ? dp2a_hi_us(w, cw[j], hi) : dp2a_hi_uu(w, cw[j], hi);
                }
                const int total = lo + (hi << 8);
                int v = total >> a.shift;
                if (SIGNED && rec.w != 0) v = total / rec.w;  // resizeXborder: truncating division by deno * 64
                v = (int)(short)v;
                *out = (uint8_t)min(max(v, 0), 255);
            }
        }
    }
}


// ---------------------------------------------------------------------------------------
// Area 2:1 x 2:1 (BASELINE config 2a): every destination pixel is a 2 x 2 block of the source,
// weights from the planner's single-phase tables (cY = cX = [1/2, 1/2] in 8 / 15 bit fixed point).
// Pure streaming: a thread reads 16 source bytes of two rows (two 16-byte loads), interleaves
// the rows (PRMT) so that dp4a(u8 x u8) gives the vertical sums of two columns, applies the two
// horizontal weights and stores 8 pixels with one 8-byte store.  No shared memory, no halo.
//   work = u16(cy0*a + cy1*c)   (AreaResizerImpl<Generic>::resizeYmain, ..._Generic.cpp:303-320)
//   dst  = (cx0*work0 + cx1*work1 + 2^22) >> 23   (resizeXmain, :340-368)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ int dp4a_uu(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

struct Area2Args {
    const uint8_t *src;
    uint8_t *dst;
    long long srcPitch, dstPitch, srcFrameStride, dstFrameStride;
    int DW, DH;
    int chunksPerRow;      // ceil(DW / 8)
    uint32_t rcpChunks;    // ceil(2^32 / chunksPerRow)
    uint32_t cyLo, cyHi;   // (cy0, cy1, 0, 0) and (0, 0, cy0, cy1) as bytes
    int cx0, cx1;
};

__global__ void __launch_bounds__(256) resizeArea2Kernel(const __grid_constant__ Area2Args a)
{
    const uint32_t idx = blockIdx.x * 256u + threadIdx.x;
    const int row = (a.chunksPerRow == 1) ? (int)idx : (int)__umulhi(idx, a.rcpChunks);
    const int ch = (int)idx - row * a.chunksPerRow;
    if (row >= a.DH) return;
    const uint8_t *__restrict__ s0 = a.src + (long long)blockIdx.y * a.srcFrameStride + (long long)(2 * row) * a.srcPitch + 16 * ch;
    const uint4 ra = __ldg(reinterpret_cast<const uint4 *>(s0));
    const uint4 rc = __ldg(reinterpret_cast<const uint4 *>(s0 + a.srcPitch));
    const uint32_t wa[4] = {ra.x, ra.y, ra.z, ra.w}, wc[4] = {rc.x, rc.y, rc.z, rc.w};
    int v[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t p01 = prmt(wa[j], wc[j], 0x5140);  // (a0, c0, a1, c1)
        const uint32_t p23 = prmt(wa[j], wc[j], 0x7362);  // (a2, c2, a3, c3)
        const int w0 = dp4a_uu(p01, a.cyLo, 0), w1 = dp4a_uu(p01, a.cyHi, 0);
        const int w2 = dp4a_uu(p23, a.cyLo, 0), w3 = dp4a_uu(p23, a.cyHi, 0);
        v[2 * j] = (w0 * a.cx0 + w1 * a.cx1 + (1 << 22)) >> 23;
        v[2 * j + 1] = (w2 * a.cx0 + w3 * a.cx1 + (1 << 22)) >> 23;
    }
    uint2 o;
    o.x = packSatU8(v[1], v[0], packSatU8(v[3], v[2], 0u));
    o.y = packSatU8(v[5], v[4], packSatU8(v[7], v[6], 0u));
    uint8_t *out = a.dst + (long long)blockIdx.y * a.dstFrameStride + (long long)row * a.dstPitch + 8 * ch;
    if (8 * ch + 8 <= a.DW) {
        *reinterpret_cast<uint2 *>(out) = o;
    } else {
        for (int p = 0; p < a.DW - 8 * ch; ++p) out[p] = (uint8_t)(((p < 4 ? o.x : o.y) >> (8 * (p & 3))) & 0xffu);
    }
}


// ---------------------------------------------------------------------------------------
// Linear up-sampling by an integer factor K (2 or 3) on the X axis (BASELINE config 2b is 3x),
// any Linear ratio on Y.  Pure streaming, no shared memory.
// A thread owns one destination row and one aligned source word (4 columns) = 4K destination
// pixels.  Destination pixel d = 4K j + i blends the source columns 4j + g(i), 4j + g(i) + 1 with
// g(i) = floor((2i + 1 - K) / 2K) in [-1, 3] and the weights of phase i mod K -- a static pattern,
// so the five possible column pairs are built once:
//   vertical    pair word (col a, col a+1) as two 16-bit lanes = lanes(row f) * q0 + lanes(row f+1) * q1
//               (two IMAD per pair; <= 255 * 256 per lane, no carry)      [resizeYmain / resizeYborder]
//   horizontal  dp2a(pair, low byte plane) then dp2a(pair, high byte plane) from (acc >> 8),
//               >> 15, saturating pack                                      [resizeXmain]
// The first and the last pixel of a row replicate the edge column (resizeXborder,
// src/IQOLinearResizerImpl_Generic.cpp:355-366).
// ---------------------------------------------------------------------------------------
struct LinearUpArgs {
    const uint8_t *src;
    uint8_t *dst;
    long long srcPitch, dstPitch, srcFrameStride, dstFrameStride;
    int SW, SH, DW, DH;
    int itemsPerRow;           // SW / (4 RS)
    uint32_t rcpItems;         // ceil(2^32 / itemsPerRow)
    const int32_t *firstY, *rowY, *coefY;   // generic vertical tables (two taps per row)
    int q1X[8];                // per phase: twice the weight of the right column (the left one is 32768 - q1)
};

#ifndef IQO_LINEAR_UP_ROWS
#define IQO_LINEAR_UP_ROWS 12
#endif
constexpr int kLinearUpRows = IQO_LINEAR_UP_ROWS;  // destination rows per item: they mostly share their two source rows

// RS : RD on X (up-sampling at 1:2, 1:3, 1:4, 2:3, 2:5, 3:4, 4:5; the mild reductions 3:2 and 4:3): an item is 4 RS source columns (RS aligned words) and the 4 RD
// destination pixels they produce, for kLinearUpRows destination rows
template <int RS, int RD>
__global__ void __launch_bounds__(256) resizeLinearUpKernel(const __grid_constant__ LinearUpArgs a)
{
    constexpr int NC = 4 * RS + 2;   // source columns 4 RS j - 1 .. 4 RS j + 4 RS
    // item = (group of kLinearUpRows destination rows, RS source words), flattened so that rows whose item
    // count is not a multiple of the block size do not leave threads idle
    const uint32_t item = blockIdx.x * 256u + threadIdx.x;
    const int grp = (a.itemsPerRow == 1) ? (int)item : (int)__umulhi(item, a.rcpItems);
    const int j = (int)item - grp * a.itemsPerRow;
    const int y0 = grp * kLinearUpRows;
    if (y0 >= a.DH) return;
    const uint8_t *__restrict__ src = a.src + (long long)blockIdx.y * a.srcFrameStride;
    uint8_t *__restrict__ out = a.dst + (long long)blockIdx.y * a.dstFrameStride + (long long)y0 * a.dstPitch + 4 * RD * j;
    const int wordsPerRow = a.itemsPerRow * RS;
    const int w0 = RS * j, wm = max(w0 - 1, 0), wp = min(w0 + RS, wordsPerRow - 1);

    // the NC source columns of the two source rows in use
    uint32_t A[NC], Bv[NC];
    int have = -(1 << 30);
#pragma unroll
    for (int r = 0; r < kLinearUpRows; ++r) {
        const int y = y0 + r;
        if (y >= a.DH) break;
        // vertical taps of this destination row (uniform over the rows of a block)
        const int fy = __ldg(a.firstY + y);
        const int ry = __ldg(a.rowY + y);
        const uint32_t q0 = (uint32_t)__ldg(a.coefY + 2 * ry), q1 = (uint32_t)__ldg(a.coefY + 2 * ry + 1);
        if (fy != have) {
            have = fy;
            const int r0 = min(max(fy, 0), a.SH - 1), r1 = min(max(fy + 1, 0), a.SH - 1);
            const uint32_t *row0 = reinterpret_cast<const uint32_t *>(src + (long long)r0 * a.srcPitch);
            const uint32_t *row1 = reinterpret_cast<const uint32_t *>(src + (long long)r1 * a.srcPitch);
            uint32_t aw[RS], bw[RS];
            const uint32_t am = __ldg(row0 + wm), ap = __ldg(row0 + wp);
            const uint32_t bm = __ldg(row1 + wm), bp = __ldg(row1 + wp);
#pragma unroll
            for (int w = 0; w < RS; ++w) aw[w] = __ldg(row0 + w0 + w), bw[w] = __ldg(row1 + w0 + w);
            A[0] = am >> 24, Bv[0] = bm >> 24;
#pragma unroll
            for (int w = 0; w < RS; ++w) {
                A[4 * w + 1] = aw[w] & 0xffu, Bv[4 * w + 1] = bw[w] & 0xffu;
                A[4 * w + 2] = prmt(aw[w], 0u, 0x4441), Bv[4 * w + 2] = prmt(bw[w], 0u, 0x4441);
                A[4 * w + 3] = prmt(aw[w], 0u, 0x4442), Bv[4 * w + 3] = prmt(bw[w], 0u, 0x4442);
                A[4 * w + 4] = aw[w] >> 24, Bv[4 * w + 4] = bw[w] >> 24;
            }
            A[NC - 1] = ap & 0xffu, Bv[NC - 1] = bp & 0xffu;
            // columns -1 and S are the replicated edge columns: beyond 3x more than the first / last pixel (which are
            // patched below) read them
            if (RD > 3 * RS) {
                if (j == 0) A[0] = A[1], Bv[0] = Bv[1];
                if (j == a.itemsPerRow - 1) A[NC - 1] = A[NC - 2], Bv[NC - 1] = Bv[NC - 2];
            }
        }
        // vertical blend per column (<= 255 * 256), then per pixel, with the two horizontal weights of a phase
        // summing to 32768:   2 * (lo * q0 + hi * q1 + 2^22) == (lo << 16) + 2^23 + (hi - lo) * 2 q1
        // -- one multiply-add per pixel; the doubled sum has the pixel in its top byte (a convex blend needs no
        // saturation, and 65280 * 65536 + 2^23 still fits 32 bits)
        uint32_t V[NC], baseN[NC - 1], diffN[NC - 1];
#pragma unroll
        for (int c = 0; c < NC; ++c) V[c] = A[c] * q0 + Bv[c] * q1;
#pragma unroll
        for (int n = 0; n < NC - 1; ++n) {
            baseN[n] = V[n] * 65536u + (1u << 23);
            diffN[n] = V[n + 1] - V[n];
        }
        uint32_t packed[RD];
#pragma unroll
        for (int quad = 0; quad < RD; ++quad) {
            uint32_t v[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int i = 4 * quad + e;
                // pixel i blends the columns (4 RS j - 1 + n, + 1), n = first tap + 1 = floor(((2 i + 1) RS + RD) / (2 RD))
                // (plan.cpp linearAxis: first = floor(((2 d + 1) S - D) / (2 D)))
                const int n = ((2 * i + 1) * RS + RD) / (2 * RD);
                v[e] = baseN[n] + diffN[n] * (uint32_t)a.q1X[i % RD];  // q1X holds 2 * q1
            }
            packed[quad] = prmt(prmt(v[0], v[1], 0x0073), prmt(v[2], v[3], 0x0073), 0x5410);  // the four top bytes
        }
        // replicated edge columns
        if (j == 0) {
            const int v = (int)(V[1] + 128u) >> 8;  // column 0
            packed[0] = (packed[0] & 0xffffff00u) | (uint32_t)min(v, 255);
        }
        if (j == a.itemsPerRow - 1) {
            const int v = (int)(V[NC - 2] + 128u) >> 8;  // column S-1
            packed[RD - 1] = (packed[RD - 1] & 0x00ffffffu) | ((uint32_t)min(v, 255) << 24);
        }
        uint8_t *o = out + (long long)r * a.dstPitch;
#pragma unroll
        for (int quad = 0; quad < RD; ++quad) *reinterpret_cast<uint32_t *>(o + 4 * quad) = packed[quad];
    }
}


// ---------------------------------------------------------------------------------------
// Area reductions at small rational ratios on X (3:2, 4:3, 2:1, 5:2, 3:1, 4:1; any ratio on Y), pure streaming:
// an item is WS aligned source words (4 WS = PS source columns) and the 4 WD = PD destination pixels they cover
// (PS RD == PD RS), for kAreaDownRows destination rows.  Vertical pass: every source row of a destination row's taps is
// loaded (L1 keeps the rows that consecutive destination rows share), its bytes spread into 16-bit lane pairs
// (one PRMT per pair) and accumulated with ONE IMAD per pair and tap (weights sum to 256: a lane never exceeds
// 65280); horizontal pass: compile-time first taps floor(i RS / RD), doubled 15-bit weights from the constant bank,
// accumulator preset to 2^23, so the pixel is the top byte of the sum (no shift, and a convex sum needs no
// saturation).  No shared memory.  Area has no border rows or columns; the one column beyond an item that its last
// pixel's zero-weight tap names is read clamped.
// ---------------------------------------------------------------------------------------
#ifndef IQO_AREA_DOWN_ROWS
#define IQO_AREA_DOWN_ROWS 8
#endif
constexpr int kAreaDownRows = IQO_AREA_DOWN_ROWS;   // destination rows per item
constexpr int kAreaDownMaxNX = 4, kAreaDownMaxRD = 3;

struct AreaDownArgs {
    const uint8_t *src;
    uint8_t *dst;
    long long srcPitch, dstPitch, srcFrameStride, dstFrameStride;
    int SW, SH, DW, DH;
    int itemsPerRow;           // SW / PS
    uint32_t rcpItems;         // ceil(2^32 / itemsPerRow)
    int NY, NYstride;          // vertical taps per row up to the last one that is non-zero in some row; taps per table row
    const int32_t *firstY, *rowY, *coefY;   // generic vertical tables
    int cx2[kAreaDownMaxRD][kAreaDownMaxNX];  // per phase: twice the horizontal weights
};

template <int WS, int WD, int RS, int RD, int NX>
__global__ void __launch_bounds__(256) resizeAreaDownKernel(const __grid_constant__ AreaDownArgs a)
{
    constexpr int PS = 4 * WS, PD = 4 * WD;
    static_assert(PS * RD == PD * RS, "item geometry");
    static_assert(((PD - 1) * RS) / RD + NX - 1 <= PS, "the last pixel reads at most one column beyond the item");
    constexpr bool kExtra = ((PD - 1) * RS) / RD + NX - 1 == PS;   // some tap names the column after the item
    const uint32_t item = blockIdx.x * 256u + threadIdx.x;
    const int grp = (a.itemsPerRow == 1) ? (int)item : (int)__umulhi(item, a.rcpItems);
    const int j = (int)item - grp * a.itemsPerRow;
    const int y0 = grp * kAreaDownRows;
    if (y0 >= a.DH) return;
    const uint8_t *__restrict__ src = a.src + (long long)blockIdx.y * a.srcFrameStride;
    uint8_t *__restrict__ out = a.dst + (long long)blockIdx.y * a.dstFrameStride + (long long)y0 * a.dstPitch + PD * j;
    const int w0 = WS * j, wx = min(w0 + WS, a.itemsPerRow * WS - 1);
    for (int r = 0; r < kAreaDownRows; ++r) {
        const int y = y0 + r;
        if (y >= a.DH) break;
        const int fy = __ldg(a.firstY + y);
        const int32_t *cy = a.coefY + __ldg(a.rowY + y) * a.NYstride;
        uint32_t P[2 * WS], Px = 0u;   // lane pairs (column 4w + 2h, + 1) and the extra column
#pragma unroll
        for (int i = 0; i < 2 * WS; ++i) P[i] = 0u;
        for (int t = 0; t < a.NY; ++t) {
            const uint32_t c = (uint32_t)__ldg(cy + t);
            if (c == 0u) continue;
            const int row = min(max(fy + t, 0), a.SH - 1);
            const uint32_t *rp = reinterpret_cast<const uint32_t *>(src + (long long)row * a.srcPitch);
            uint32_t wv[WS];
#pragma unroll
            for (int w = 0; w < WS; ++w) wv[w] = __ldg(rp + w0 + w);
            const uint32_t we = kExtra ? __ldg(rp + wx) : 0u;
#pragma unroll
            for (int w = 0; w < WS; ++w) {
                P[2 * w] += prmt(wv[w], 0u, 0x4140) * c;
                P[2 * w + 1] += prmt(wv[w], 0u, 0x4342) * c;
            }
            Px += (we & 0xffu) * c;
        }
        uint32_t V[PS + 1];
#pragma unroll
        for (int i = 0; i < 2 * WS; ++i) V[2 * i] = P[i] & 0xffffu, V[2 * i + 1] = P[i] >> 16;
        V[PS] = Px;
        uint32_t packed[WD];
#pragma unroll
        for (int quad = 0; quad < WD; ++quad) {
            uint32_t v[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int i = 4 * quad + e;
                const int n = (i * RS) / RD;   // plan.cpp areaAxis: first = floor(d S / D)
                uint32_t acc = 1u << 23;
#pragma unroll
                for (int k = 0; k < NX; ++k) acc += V[n + k] * (uint32_t)a.cx2[i % RD][k];
                v[e] = acc;
            }
            packed[quad] = prmt(prmt(v[0], v[1], 0x0073), prmt(v[2], v[3], 0x0073), 0x5410);  // the four top bytes
        }
        uint8_t *o = out + (long long)r * a.dstPitch;
#pragma unroll
        for (int quad = 0; quad < WD; ++quad) *reinterpret_cast<uint32_t *>(o + 4 * quad) = packed[quad];
    }
}


// ---------------------------------------------------------------------------------------
// Streaming 2:1 x 2:1 Lanczos for kernels with at most four non-zero taps per axis (the pxScale=2
// chroma planes of YUV420 4K->1080p, Lanczos1).  No shared memory: a thread owns 8 adjacent
// destination pixels (16 source columns + one word either side) for 6 consecutive destination
// rows.  Source rows are read once, top to bottom (LDG.128 + 2 LDG.32), spread into 16-bit lane
// pairs (PRMT) and accumulated straight into the destination rows they belong to: one IMAD per
// lane pair and tap (packed-lane trick of resizePackedKernel; bias keeps lanes in [0, 65535]).
// A finished row goes through the dp2a byte-plane horizontal pass and leaves as one 8-byte store.
// Everything is unrolled, so rows in flight live in registers.  Border rows use masked
// coefficients + the multiply-high division; border columns are recomputed afterwards from the
// generic tables, one pixel per thread.
// ---------------------------------------------------------------------------------------
constexpr int kSmallRows = 6;

// one destination pixel exactly as the reference computes it (any kind of row / column)
__device__ __noinline__ uint8_t genericPixel(const AxisDev &gx, const AxisDev &gy, const uint8_t *src, long long pitch, int y, int d,
                                             int shift, bool lanczos)
{
    const int fx = __ldg(gx.first + d), rx = __ldg(gx.row + d);
    const int fy = __ldg(gy.first + y), ry = __ldg(gy.row + y);
    const int denoY = __ldg(gy.deno + ry);
    int nume = 0;
    for (int i = 0; i < gx.N; ++i) {
        const int cx = __ldg(gx.coef + rx * gx.N + i);
        if (cx == 0) continue;
        const int col = min(max(fx + i, 0), gx.S - 1);
        int acc = 0;
        for (int t = 0; t < gy.N; ++t) {
            const int row = min(max(fy + t, 0), gy.S - 1);
            acc += __ldg(gy.coef + ry * gy.N + t) * (int)__ldg(src + (long long)row * pitch + col);
        }
        if (denoY != 0) acc = ((int)(short)acc * 64) / denoY;
        nume += cx * (int)(short)acc;
    }
    return finishPixel(nume, lanczos ? __ldg(gx.deno + rx) : 0, shift);
}

template <int TY, int NW, int WB, bool EDGE>
__device__ __forceinline__ void smallRows(const SmallArgs &a, const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, int c, int y0)
{
    constexpr int R = kSmallRows;
    constexpr int NS = 2 * R + TY - 2;  // source rows consumed
    const int B = a.workBias;
    const uint32_t biased = (uint32_t)B | ((uint32_t)B << 16);
    int coef[R][TY];
    int deno[R];
    uint32_t magic[R];
#pragma unroll
    for (int j = 0; j < R; ++j) {
        deno[j] = 0;
        magic[j] = 0;
#pragma unroll
        for (int t = 0; t < TY; ++t) coef[j][t] = a.cY[t];
        if (EDGE) {
            const int y = y0 + j;
            if (y < a.DH && (y < a.mbY || y >= a.meY)) {
                const int row = __ldg(a.gy.row + y);
                deno[j] = __ldg(a.gy.deno + row);
                magic[j] = __ldg(a.magicY + row);
#pragma unroll
                for (int t = 0; t < TY; ++t) coef[j][t] = __ldg(a.rowsY + row * 4 + t);
            }
        }
    }
    uint32_t acc[R][10];
#pragma unroll
    for (int j = 0; j < R; ++j)
#pragma unroll
        for (int m = 0; m < 10; ++m) acc[j][m] = (EDGE && deno[j]) ? 0u : biased;

    const uint8_t *base = src + 16 * c;
    const int prevOff = (c > 0) ? -4 : 0, nextOff = (16 * c + 16 < a.SW) ? 16 : 12;  // clamped: edge values meet zero weights
    // source rows are fetched two rows ahead of their use (the loop is fully unrolled, so the
    // three buffers are plain registers): global-load latency overlaps the arithmetic
    auto fetch = [&](int s, uint4 &w, uint32_t &wp, uint32_t &wn) {
        int row = 2 * y0 + a.cy0 + s;
        if (EDGE) row = min(max(row, 0), a.SH - 1);
        const uint8_t *rp = base + (long long)row * a.srcPitch;
        w = __ldg(reinterpret_cast<const uint4 *>(rp));
        wp = __ldg(reinterpret_cast<const uint32_t *>(rp + prevOff));
        wn = __ldg(reinterpret_cast<const uint32_t *>(rp + nextOff));
    };
    uint4 bw[3];
    uint32_t bp[3], bn[3];
    fetch(0, bw[0], bp[0], bn[0]);
    if (NS > 1) fetch(1, bw[1], bp[1], bn[1]);
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        if (s + 2 < NS) fetch(s + 2, bw[(s + 2) % 3], bp[(s + 2) % 3], bn[(s + 2) % 3]);
        const uint4 w = bw[s % 3];
        const uint32_t wp = bp[s % 3], wn = bn[s % 3];
        uint32_t L[10];
        L[0] = prmt(wp, 0u, 0x4342);
        L[1] = prmt(w.x, 0u, 0x4140);
        L[2] = prmt(w.x, 0u, 0x4342);
        L[3] = prmt(w.y, 0u, 0x4140);
        L[4] = prmt(w.y, 0u, 0x4342);
        L[5] = prmt(w.z, 0u, 0x4140);
        L[6] = prmt(w.z, 0u, 0x4342);
        L[7] = prmt(w.w, 0u, 0x4140);
        L[8] = prmt(w.w, 0u, 0x4342);
        L[9] = prmt(wn, 0u, 0x4140);
#pragma unroll
        for (int j = 0; j < R; ++j) {
            const int t = s - 2 * j;
            if (t >= 0 && t < TY) {
                const uint32_t cf = (uint32_t)coef[j][t];
#pragma unroll
                for (int m = 1 + WB; m <= 8 + WB + NW - 1; ++m) acc[j][m] += L[m] * cf;
            }
        }
        // the destination row whose last tap this was is complete
        if (s >= TY - 1 && ((s - (TY - 1)) & 1) == 0) {
            const int j = (s - (TY - 1)) >> 1;
            const int y = y0 + j;
            if (j < R && y < a.DH) {
                if (EDGE && deno[j]) {
#pragma unroll
                    for (int m = 1 + WB; m <= 8 + WB + NW - 1; ++m) {
                        const uint32_t lo = acc[j][m] & 0xffffu;
                        const uint32_t hi = (acc[j][m] >> 16) + ((lo & 0x8000u) ? 1u : 0u);
                        uint32_t out = 0;
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int n = (int)(short)(h ? hi : lo) * 64;
                            const uint32_t mag = (uint32_t)abs(n);
                            const int qv = magic[j] ? (int)__umulhi(mag, magic[j]) : (int)mag;
                            const int wv = (int)(short)(n < 0 ? -qv : qv) + B;
                            out |= ((uint32_t)wv & 0xffffu) << (16 * h);
                        }
                        acc[j][m] = out;
                    }
                }
                int v[8];
#pragma unroll
                for (int p = 0; p < 8; ++p) {
                    int x = a.accInit;
#pragma unroll
                    for (int q = 0; q < NW; ++q) x = dp2a_lo_uu(acc[j][1 + p + WB + q], a.cwX[q], x);
                    x >>= 8;
#pragma unroll
                    for (int q = 0; q < NW; ++q) x = dp2a_hi_us(acc[j][1 + p + WB + q], a.cwX[q], x);
                    v[p] = x >> 12;
                }
                uint2 o;
                o.x = packSatU8(v[1], v[0], packSatU8(v[3], v[2], 0u));
                o.y = packSatU8(v[5], v[4], packSatU8(v[7], v[6], 0u));
                *reinterpret_cast<uint2 *>(dst + (long long)y * a.dstPitch + 8 * c) = o;
            }
        }
    }
}

template <int TY, int NW, int WB>
__global__ void __launch_bounds__(128) resizeHalfSmallKernel(const __grid_constant__ SmallArgs a)
{
    const int c = blockIdx.x * 128 + threadIdx.x;  // 8-pixel chunk of the row
    const int y0 = blockIdx.y * kSmallRows;
    const uint8_t *__restrict__ src = a.src + (long long)blockIdx.z * a.srcFrameStride;
    uint8_t *__restrict__ dst = a.dst + (long long)blockIdx.z * a.dstFrameStride;
    const int chunks = a.DW >> 3;
    if (c < chunks) {
        const int lastSrcRow = 2 * y0 + a.cy0 + 2 * kSmallRows + TY - 3;
        const bool edge = (2 * y0 + a.cy0 < 0) || (lastSrcRow >= a.SH) || (y0 < a.mbY) || (y0 + kSmallRows > a.meY);
        if (edge)
            smallRows<TY, NW, WB, true>(a, src, dst, c, y0);
        else
            smallRows<TY, NW, WB, false>(a, src, dst, c, y0);
    }
    // border columns of this block's pixel range
    const int px0 = blockIdx.x * 128 * 8, px1 = min(px0 + 128 * 8, a.DW);
    const int l0 = px0, l1 = min(px1, a.mbX);                 // left border columns [l0, l1)
    const int r0 = max(px0, max(a.meX, a.mbX)), r1 = px1;     // right border columns [r0, r1)
    const int nl = max(l1 - l0, 0), nr = max(r1 - r0, 0);
    if (nl + nr > 0) {
        __syncthreads();  // the streamed stores of these pixels come first
        const int rows = min(kSmallRows, a.DH - y0);
        for (int item = threadIdx.x; item < rows * (nl + nr); item += blockDim.x) {
            const int r = item / (nl + nr), k = item - r * (nl + nr);
            const int d = (k < nl) ? l0 + k : r0 + (k - nl);
            dst[(long long)(y0 + r) * a.dstPitch + d] = genericPixel(a.gx, a.gy, src, a.srcPitch, y0 + r, d, 20, true);
        }
    }
}


// ---------------------------------------------------------------------------------------
// Rational-ratio streaming kernel (plan.hpp RatioPlan): Lanczos at RS:RD with RD | 8 (3:2, 1:2, 3:4),
// at most 10 horizontal taps.  Same organisation as resizeHalfStreamKernel -- one warp per CTA
// walks down a column strip, dp4a vertical pass, dp2a horizontal pass on 8 pixels per lane --
// but table-driven where the 2:1 kernel is hard-wired:
//   vertical    a lane owns 8 source columns.  Every new group of four source rows is transposed
//               once (2 x 8 PRMT) and parked in a lane-private ring of 4 groups in shared memory;
//               a destination row reads the <= 4 groups its record names (2 LDS.128 + 8 dp4a each,
//               coefficient words from the per-row record, so phases and border rows need no code).
//   horizontal  every 8 destination rows: lane = (row, group of 8 pixels).  Pixel p of a group starts
//               at W element GS q + floor(p RS / RD) (+ an even offset): word offsets and parities
//               are compile-time, the coefficient words of (phase, parity) sit in the constant bank.
//   borders     border rows: masked words + multiply-high division from the record; border columns
//               are recomputed per pixel from the generic tables.
// ---------------------------------------------------------------------------------------
constexpr int kRatioRing = 4;                                   // transposed groups per lane
constexpr int kRatioRingBytes = kRatioRing * 32 * 32;           // [slot][half A/B][lane][16 bytes]
constexpr int kRatioRecBytes = 2 * 8 * 32;                      // row records of two turns
constexpr int kRatioSmem = kRatioRingBytes + 8 * kStreamRowBytes + kRatioRecBytes;

#ifndef IQO_RATIO_MINB
#define IQO_RATIO_MINB 24
#endif
template <int RS, int RD, int NX, int TZ, bool ODD>
__global__ void __launch_bounds__(32, IQO_RATIO_MINB) resizeRatioStreamKernel(const __grid_constant__ RatioArgs a)
{
    extern __shared__ __align__(16) uint8_t ratioSmem[];
    constexpr int GS = 8 * RS / RD;
    const int lane = threadIdx.x;
    const int tx0 = blockIdx.x * (8 * a.groupsPerStrip);
    const int y0 = blockIdx.y * a.bandRows;
    const int y1 = min(y0 + a.bandRows, a.DH);
    const uint8_t *__restrict__ src = a.src + (long long)blockIdx.z * a.srcFrameStride;
    uint8_t *__restrict__ dst = a.dst + (long long)blockIdx.z * a.dstFrameStride;
    const int ngs = min(a.groupsPerStrip, (a.DW - tx0) >> 3);   // 8-pixel groups of this strip
    const int first0 = GS * (tx0 >> 3) + a.c0;                  // first tap of pixel tx0
    const int xs0 = first0 & ~7;                                 // first source column of lane 0 (may be negative)
    // W element i is source column xs0 + i (ODD: xs0 + i - 1, so that the first tap sits on an even element)
    const int i0 = first0 - xs0 + (ODD ? 1 : 0);
    const uint32_t ringBase = smemAddr(ratioSmem) + 16 * lane;
    const uint32_t wBase = smemAddr(ratioSmem) + kRatioRingBytes;
    const int B = a.workBias;
    const int SHm1 = a.SH - 1;
    const long long pitch = a.srcPitch;

    // vertical role: columns outside the image only ever meet zero coefficients (read column 0)
    const int col = xs0 + 8 * lane;
    const uint8_t *base = src + ((col >= 0 && col < a.SW) ? col : 0);
    uint2 raw[4];  // group gNext (loads in flight; two groups ahead measured no faster)
    auto fetch = [&](int g, uint2 (&buf)[4]) {
        if (g >= 0 && 4 * g + 3 <= SHm1) {
            const uint8_t *p = base + (long long)(4 * g) * pitch;
#pragma unroll
            for (int j = 0; j < 4; ++j) buf[j] = __ldg(reinterpret_cast<const uint2 *>(p + j * pitch));
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int row = min(max(4 * g + j, 0), SHm1);
                buf[j] = __ldg(reinterpret_cast<const uint2 *>(base + (long long)row * pitch));
            }
        }
    };
    auto park = [&](int g) {  // transposed group -> ring slot g mod 8
        const uint32_t t0 = prmt(raw[0].x, raw[1].x, 0x5140), t1 = prmt(raw[0].x, raw[1].x, 0x7362);
        const uint32_t t2 = prmt(raw[2].x, raw[3].x, 0x5140), t3 = prmt(raw[2].x, raw[3].x, 0x7362);
        uint4 ca, cb;
        ca.x = prmt(t0, t2, 0x5410);
        ca.y = prmt(t0, t2, 0x7632);
        ca.z = prmt(t1, t3, 0x5410);
        ca.w = prmt(t1, t3, 0x7632);
        const uint32_t u0 = prmt(raw[0].y, raw[1].y, 0x5140), u1 = prmt(raw[0].y, raw[1].y, 0x7362);
        const uint32_t u2 = prmt(raw[2].y, raw[3].y, 0x5140), u3 = prmt(raw[2].y, raw[3].y, 0x7362);
        cb.x = prmt(u0, u2, 0x5410);
        cb.y = prmt(u0, u2, 0x7632);
        cb.z = prmt(u1, u3, 0x5410);
        cb.w = prmt(u1, u3, 0x7632);
        const uint32_t sa = ringBase + (g & (kRatioRing - 1)) * 1024;
        stsV4<0>(sa, ca);
        stsV4<512>(sa, cb);
    };

    int gNext = __ldg(a.rowRec + 8 * y0);  // groups are non-decreasing in y
    fetch(gNext, raw);
    const uint32_t rcp = (65536u + ngs - 1) / ngs;
    const bool left = tx0 < a.mbX, right = tx0 + 8 * ngs > a.meX;

    // The 8 row records of a turn (256 bytes) are staged in shared memory one turn ahead: lanes 0..15
    // each carry 16 bytes of the next turn's records in a register while the current turn runs.
    const uint32_t recBase = wBase + 8 * kStreamRowBytes;
    const int recMax = 8 * a.DH - 4;  // last 16-byte piece of the table
    auto recLoad = [&](int yTurn) -> int4 {
        const int w = min(8 * yTurn + 4 * (lane & 15), recMax);
        return __ldg(reinterpret_cast<const int4 *>(a.rowRec + w));
    };
    {
        const int4 cur = recLoad(y0);
        if (lane < 16) stsV4<0>(recBase + 16 * lane, make_uint4(cur.x, cur.y, cur.z, cur.w));
        __syncwarp();
    }
    int4 recNext = recLoad(y0 + 8);
    uint32_t recCur = recBase;  // records of the running turn
    for (int y = y0; y < y1; ++y) {
        const uint32_t ra = recCur + 32 * ((y - y0) & 7);
        const uint4 u0 = ldsV4<0>(ra), u1 = ldsV4<16>(ra);
        const int4 r0 = make_int4(u0.x, u0.y, u0.z, u0.w), r1 = make_int4(u1.x, u1.y, u1.z, u1.w);
        const int g0 = r0.x, ng = r0.y;
        while (gNext < g0 + ng) {  // uniform
            park(gNext);
            ++gNext;
            fetch(gNext, raw);
        }
        const int deno = r1.z;
        const uint32_t magic = (uint32_t)r1.w;
        const int init = deno ? 0 : B;
        int v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = init;
        // straight-line code per group count: all ring loads go out before the first dp4a
        auto rowSum = [&](auto ngTag) {
            constexpr int NGc = decltype(ngTag)::value;
            uint4 qa[NGc], qb[NGc];
#pragma unroll
            for (int t = 0; t < NGc; ++t) {
                const uint32_t sa = ringBase + ((g0 + t) & (kRatioRing - 1)) * 1024;
                qa[t] = ldsV4<0>(sa);
                qb[t] = ldsV4<512>(sa);
            }
#pragma unroll
            for (int t = 0; t < NGc; ++t) {
                const uint32_t c = (uint32_t)(t == 0 ? r0.z : t == 1 ? r0.w : t == 2 ? r1.x : r1.y);
                v[0] = dp4a_us(qa[t].x, c, v[0]);
                v[1] = dp4a_us(qa[t].y, c, v[1]);
                v[2] = dp4a_us(qa[t].z, c, v[2]);
                v[3] = dp4a_us(qa[t].w, c, v[3]);
                v[4] = dp4a_us(qb[t].x, c, v[4]);
                v[5] = dp4a_us(qb[t].y, c, v[5]);
                v[6] = dp4a_us(qb[t].z, c, v[6]);
                v[7] = dp4a_us(qb[t].w, c, v[7]);
            }
        };
        if (ng == 3)
            rowSum(std::integral_constant<int, 3>());
        else if (ng == 2)
            rowSum(std::integral_constant<int, 2>());
        else if (ng == 4)
            rowSum(std::integral_constant<int, 4>());
        else
            rowSum(std::integral_constant<int, 1>());
        if (deno) {
            // resizeYborder: see halfVerticalStrip
            auto bdiv = [&](int x) -> int {
                const int n = (int)(short)x * 64;
                const uint32_t m = (uint32_t)abs(n);
                const int q = magic ? (int)__umulhi(m, magic) : (int)m;
                return (int)(short)(n < 0 ? -q : q) + B;
            };
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = bdiv(v[i]);
        }
        uint4 o;
        if (ODD) {
            // pair words (column 2m-1, column 2m): the lane's last column travels to its right neighbour
            const uint32_t prev = __shfl_up_sync(0xffffffffu, (uint32_t)v[7], 1);
            o.x = prmt(prev, (uint32_t)v[0], 0x5410);
            o.y = prmt((uint32_t)v[1], (uint32_t)v[2], 0x5410);
            o.z = prmt((uint32_t)v[3], (uint32_t)v[4], 0x5410);
            o.w = prmt((uint32_t)v[5], (uint32_t)v[6], 0x5410);
        } else {
            o.x = prmt((uint32_t)v[0], (uint32_t)v[1], 0x5410);
            o.y = prmt((uint32_t)v[2], (uint32_t)v[3], 0x5410);
            o.z = prmt((uint32_t)v[4], (uint32_t)v[5], 0x5410);
            o.w = prmt((uint32_t)v[6], (uint32_t)v[7], 0x5410);
        }
        const int slot = (y - y0) & 7;
        stsV4<0>(wBase + slot * kStreamRowBytes + 16 * lane, o);
        if (slot != 7 && y != y1 - 1) continue;

        // ---- horizontal pass of the parked rows ----
        __syncwarp();
        const int nr = slot + 1, yt = y - slot;
        for (int item = lane; item < nr * ngs; item += 32) {
            const int r = (int)(((uint32_t)item * rcp) >> 16);
            const int q = item - r * ngs;
            const uint32_t wl = wBase + r * kStreamRowBytes + 2 * (i0 + GS * q);
            constexpr int NXE = NX - TZ;  // taps left after the zero taps that end every phase
            constexpr int kWords = (((7 * RS / RD) >> 1) + (NXE + 2) / 2 + 1) & ~1;  // pair words the 8 pixels span (even count)
            uint32_t n[kWords];
            if ((wl & 7) == 0) {  // uniform: 8-byte aligned rows of words
#pragma unroll
                for (int j = 0; j < kWords; j += 2)
                    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(n[j]), "=r"(n[j + 1]) : "r"(wl + 4 * j) : "memory");
            } else {
#pragma unroll
                for (int j = 0; j < kWords; ++j) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(n[j]) : "r"(wl + 4 * j) : "memory");
            }
            int px[8];
#pragma unroll
            for (int p = 0; p < 8; ++p) {
                const int off = (p * RS) / RD;
                const int wp = off >> 1, par = off & 1, ph = p % RD;
                const int nw = (NXE + par + 1) / 2;
                int acc = a.accInit;
#pragma unroll
                for (int j = 0; j < nw; ++j) acc = dp2a_lo_uu(n[wp + j], a.cwX[ph][par][j], acc);
                acc >>= 8;
#pragma unroll
                for (int j = 0; j < nw; ++j) acc = dp2a_hi_us(n[wp + j], a.cwX[ph][par][j], acc);
                px[p] = acc >> 12;
            }
            uint2 o2;
            o2.x = packSatU8(px[1], px[0], packSatU8(px[3], px[2], 0u));
            o2.y = packSatU8(px[5], px[4], packSatU8(px[7], px[6], 0u));
            uint8_t *out = dst + (long long)(yt + r) * a.dstPitch + tx0 + 8 * q;
            if (a.dstVec)
                *reinterpret_cast<uint2 *>(out) = o2;
            else
                halfStoreBytes(out, o2, 0, 8);
        }
        if (left || right) {
            __syncwarp();  // the main stores of these pixels come first
            const int c0 = left ? tx0 : max(a.meX, tx0);
            const int c1 = left ? min(a.mbX, tx0 + 8 * ngs) : tx0 + 8 * ngs;
            // a strip that holds both borders (narrow images) handles the right one in a second sweep
            for (int side = 0; side < 2; ++side) {
                int b0 = c0, b1 = c1;
                if (side == 1) {
                    if (!(left && right)) break;
                    b0 = max(a.meX, max(tx0, a.mbX));
                    b1 = tx0 + 8 * ngs;
                }
                const int nb = b1 - b0;
                for (int item = lane; item < nr * nb; item += 32) {
                    // resizeXborder from the parked W row: masked taps, truncating division (finishPixel)
                    const int r = item / nb, d = b0 + item - r * nb;
                    const int fx = __ldg(a.gx.first + d), rx = __ldg(a.gx.row + d);
                    const int32_t *cx = a.gx.coef + rx * NX;
                    const uint32_t wr = wBase + r * kStreamRowBytes + 2 * (fx - xs0 + (ODD ? 1 : 0));
                    int nume = 0;
#pragma unroll
                    for (int i = 0; i < NX; ++i) {
                        uint32_t w16;
                        asm volatile("ld.shared.u16 %0, [%1];" : "=r"(w16) : "r"(wr + 2 * i) : "memory");
                        nume += ((int)w16 - B) * __ldg(cx + i);
                    }
                    dst[(long long)(yt + r) * a.dstPitch + d] = finishPixel(nume, __ldg(a.gx.deno + rx), 20);
                }
            }
        }
        // next turn's records: register -> the other half of the record buffer, then fetch the turn after
        recCur = recBase + (recCur == recBase ? 256 : 0);
        if (lane < 16) stsV4<0>(recCur + 16 * lane, make_uint4(recNext.x, recNext.y, recNext.z, recNext.w));
        recNext = recLoad(y + 9);
        __syncwarp();  // the W rows are free again, the records are visible
    }
}

template <int RS, int RD, int NX, int TZ, bool ODD>
cudaError_t launchRatioT(const RatioArgs &a, cudaStream_t stream)
{
    const int strip = 8 * a.groupsPerStrip;
    dim3 grid((a.DW + strip - 1) / strip, (a.DH + a.bandRows - 1) / a.bandRows, a.nFrames);
    resizeRatioStreamKernel<RS, RD, NX, TZ, ODD><<<grid, 32, kRatioSmem, stream>>>(a);
    g_launches.fetch_add(1);
    return cudaGetLastError();
}


// ---------------------------------------------------------------------------------------
// General Lanczos streaming kernel (plan.hpp LStreamPlan): any ratio, phase count and row band.
// The record-driven dp4a vertical pass of resizeRatioStreamKernel (a warp walks down a strip whose
// source window fits 256 columns; a lane owns 8 source columns; transposed 4-row groups in a
// lane-private ring of 8) feeds the column-resident horizontal pass of resizePackedKernel: every 8
// destination rows a lane takes the strip's destination columns lane, lane + 32, ..., loads the
// column's record and its NPT coefficient pair words once and produces the column's 8 pixels
// (NPT LDS + 2 NPT dp2a each; Lanczos border columns divide).
// ---------------------------------------------------------------------------------------
constexpr int kLStreamRing = 8;
constexpr int kLStreamSmem = kLStreamRing * 1024 + 8 * kStreamRowBytes + 2 * 8 * 64;

#ifndef IQO_LSTREAM_MINB
#define IQO_LSTREAM_MINB 12
#endif
template <int NPT>
__global__ void __launch_bounds__(32, IQO_LSTREAM_MINB) resizeLanczosStreamKernel(const __grid_constant__ LStreamArgs a)
{
    extern __shared__ __align__(16) uint8_t lsSmem[];
    const int lane = threadIdx.x;
    const int tx0 = blockIdx.x * a.stripW;
    const int tw = min(a.stripW, a.DW - tx0);
    const int y0 = a.dstRow0 + blockIdx.y * a.bandRows;
    const int y1 = min(y0 + a.bandRows, a.dstRow0 + a.dstRows);
    const uint8_t *__restrict__ src = a.src + (long long)blockIdx.z * a.srcFrameStride;
    uint8_t *__restrict__ dst = a.dst + (long long)blockIdx.z * a.dstFrameStride;
    const int x0 = __ldg(&a.recX[tx0].x) & ~7;  // source column of W element 0
    const uint32_t ringBase = smemAddr(lsSmem) + 16 * lane;
    const uint32_t wBase = smemAddr(lsSmem) + kLStreamRing * 1024;
    const uint32_t recBase = wBase + 8 * kStreamRowBytes;
    const int B = a.workBias;
    const long long pitch = a.srcPitch;
    const int rowsM1 = a.srcRows - 1;

    // vertical role: 8 source columns; columns at and beyond the image edge only ever meet zero coefficients
    const int col = x0 + 8 * lane;
    const int colMode = col + 8 <= a.SW ? 0 : col < a.SW ? 1 : 2;  // whole word / straddles the edge / outside
    const uint8_t *base = src + (colMode == 2 ? 0 : col);
    uint2 raw[4];
    auto fetch = [&](int g) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int row = min(max(4 * g + j - a.srcRow0, 0), rowsM1);
            const uint8_t *p = base + (long long)row * pitch;
            if (colMode != 1) {
                raw[j] = __ldg(reinterpret_cast<const uint2 *>(p));
            } else {  // the image ends inside this word: never read past the row
                uint32_t lo = 0, hi = 0;
                for (int b = 0; b < 8 && col + b < a.SW; ++b) {
                    const uint32_t v = __ldg(p + b);
                    if (b < 4) lo |= v << (8 * b);
                    else hi |= v << (8 * (b - 4));
                }
                raw[j] = make_uint2(lo, hi);
            }
        }
    };
    auto park = [&](int g) {  // transposed group -> ring slot g mod 8
        const uint32_t t0 = prmt(raw[0].x, raw[1].x, 0x5140), t1 = prmt(raw[0].x, raw[1].x, 0x7362);
        const uint32_t t2 = prmt(raw[2].x, raw[3].x, 0x5140), t3 = prmt(raw[2].x, raw[3].x, 0x7362);
        uint4 ca, cb;
        ca.x = prmt(t0, t2, 0x5410);
        ca.y = prmt(t0, t2, 0x7632);
        ca.z = prmt(t1, t3, 0x5410);
        ca.w = prmt(t1, t3, 0x7632);
        const uint32_t u0 = prmt(raw[0].y, raw[1].y, 0x5140), u1 = prmt(raw[0].y, raw[1].y, 0x7362);
        const uint32_t u2 = prmt(raw[2].y, raw[3].y, 0x5140), u3 = prmt(raw[2].y, raw[3].y, 0x7362);
        cb.x = prmt(u0, u2, 0x5410);
        cb.y = prmt(u0, u2, 0x7632);
        cb.z = prmt(u1, u3, 0x5410);
        cb.w = prmt(u1, u3, 0x7632);
        const uint32_t sa = ringBase + (g & (kLStreamRing - 1)) * 1024;
        stsV4<0>(sa, ca);
        stsV4<512>(sa, cb);
    };

    // the 8 row records of a turn (512 bytes) are staged in shared memory one turn ahead (16 bytes per lane)
    const int recMax = 16 * (a.dstRow0 + a.dstRows) - 4;
    auto recLoad = [&](int yTurn) -> int4 {
        const int w = min(16 * yTurn + 4 * lane, recMax);
        return __ldg(reinterpret_cast<const int4 *>(a.rowRec + w));
    };
    {
        const int4 cur = recLoad(y0);
        stsV4<0>(recBase + 16 * lane, make_uint4(cur.x, cur.y, cur.z, cur.w));
        __syncwarp();
    }
    int4 recNext = recLoad(y0 + 8);
    uint32_t recCur = recBase;
    int gNext = __ldg(a.rowRec + 16 * y0);  // groups are non-decreasing in y
    fetch(gNext);

    for (int y = y0; y < y1; ++y) {
        const uint32_t ra = recCur + 64 * ((y - y0) & 7);
        const uint4 r0 = ldsV4<0>(ra);
        const int g0 = (int)r0.x, ng = (int)r0.y, deno = (int)r0.z;
        const uint32_t magic = r0.w;
        while (gNext < g0 + ng) {  // uniform
            park(gNext);
            ++gNext;
            fetch(gNext);
        }
        const int init = deno ? 0 : B;
        int v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = init;
        for (int t = 0; t < ng; ++t) {  // uniform
            uint32_t c;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(c) : "r"(ra + 16 + 4 * t) : "memory");
            const uint32_t sa = ringBase + ((g0 + t) & (kLStreamRing - 1)) * 1024;
            const uint4 qa = ldsV4<0>(sa), qb = ldsV4<512>(sa);
            v[0] = dp4a_us(qa.x, c, v[0]);
            v[1] = dp4a_us(qa.y, c, v[1]);
            v[2] = dp4a_us(qa.z, c, v[2]);
            v[3] = dp4a_us(qa.w, c, v[3]);
            v[4] = dp4a_us(qb.x, c, v[4]);
            v[5] = dp4a_us(qb.y, c, v[5]);
            v[6] = dp4a_us(qb.z, c, v[6]);
            v[7] = dp4a_us(qb.w, c, v[7]);
        }
        if (deno) {
            // resizeYborder: see halfVerticalStrip
            auto bdiv = [&](int x) -> int {
                const int n = (int)(short)x * 64;
                const uint32_t m = (uint32_t)abs(n);
                const int q = magic ? (int)__umulhi(m, magic) : (int)m;
                return (int)(short)(n < 0 ? -q : q) + B;
            };
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = bdiv(v[i]);
        }
        uint4 o;
        o.x = prmt((uint32_t)v[0], (uint32_t)v[1], 0x5410);
        o.y = prmt((uint32_t)v[2], (uint32_t)v[3], 0x5410);
        o.z = prmt((uint32_t)v[4], (uint32_t)v[5], 0x5410);
        o.w = prmt((uint32_t)v[6], (uint32_t)v[7], 0x5410);
        const int slot = (y - y0) & 7;
        stsV4<0>(wBase + slot * kStreamRowBytes + 16 * lane, o);
        if (slot != 7 && y != y1 - 1) continue;

        // ---- horizontal pass of the parked rows: a lane per destination column ----
        __syncwarp();
        const int nr = slot + 1, yt = y - slot;
        for (int dx = lane; dx < tw; dx += 32) {
            const int4 rec = __ldg(a.recX + tx0 + dx);  // {first column, coefficient word offset, accumulator init, divisor}
            uint32_t cw[NPT];
#pragma unroll
            for (int j = 0; j < NPT; ++j) cw[j] = __ldg(a.cwX + rec.y + j);
            uint32_t wp = wBase + 4 * ((rec.x - x0) >> 1);
            uint8_t *out = dst + (long long)(yt - a.dstRow0) * a.dstPitch + tx0 + dx;
            for (int r = 0; r < nr; ++r, wp += kStreamRowBytes, out += a.dstPitch) {
                int lo = rec.z, hi = 0;
#pragma unroll
                for (int j = 0; j < NPT; ++j) {
                    uint32_t w;
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(wp + 4 * j) : "memory");
                    lo = dp2a_lo_uu(w, cw[j], lo);
                    hi = dp2a_hi_us(w, cw[j], hi);
                }
                const int total = lo + (hi << 8);
                int px = total >> 20;
                if (rec.w != 0) px = total / rec.w;  // resizeXborder: truncating division by deno * 64
                px = (int)(short)px;
                *out = (uint8_t)min(max(px, 0), 255);
            }
        }
        // next turn's records: register -> the other half of the record buffer, then fetch the turn after
        recCur = recBase + (recCur == recBase ? 512 : 0);
        stsV4<0>(recCur + 16 * lane, make_uint4(recNext.x, recNext.y, recNext.z, recNext.w));
        recNext = recLoad(y + 9);
        __syncwarp();  // the W rows are free again, the records are visible
    }
}

template <int NPT>
cudaError_t launchLStreamT(const LStreamArgs &a, cudaStream_t stream)
{
    dim3 grid((a.DW + a.stripW - 1) / a.stripW, (a.dstRows + a.bandRows - 1) / a.bandRows, a.nFrames);
    resizeLanczosStreamKernel<NPT><<<grid, 32, kLStreamSmem, stream>>>(a);
    g_launches.fetch_add(1);
    return cudaGetLastError();
}


// ---------------------------------------------------------------------------------------
// Tensor-path kernel (plan.hpp MmaPlan): Lanczos at any ratio, phase count and row band; Area / Linear at general
// ratios.  Both passes are banded integer matrix products on the legacy integer tensor path (mma.sync.m16n8k16 / k32,
// SASS IMMA.16816 / IMMA.16832: measured on B200 at 2048 MAC/clk/SM against 256 for dp4a, tools/mma_probe.cu), exact
// like every integer sum.
//   A CTA of a.warps warps owns a strip of destination columns and walks down 16-row destination blocks; FIFO, W tile,
//   output tile and tables are shared, segments / tiles / store rows are dealt out round robin from the last warp down
//   (warp 0 also requests the rows), two CTA barriers per block separate the phases.
//   source      raw source rows travel global -> shared by TMA in chunks of 8 rows: chunk c lives in FIFO slot
//               c mod a.nChunks, one mbarrier phase per block's requests; the chunks the next block needs are requested
//               right after the vertical pass of the running block, so their latency hides behind its horizontal pass.
//   vertical    per 16 source columns: ldmatrix.m16n16.trans.b8 turns 32 source rows x 16 columns straight into the
//               B fragments (four vertically adjacent bytes per register -- no PRMT transposes, no lane-private ring);
//               k slot -> FIFO row is a planner table of byte offsets; the A fragments are the block's coefficient
//               bytes (1 - 3 k-steps of 32 source rows); results + bias are packed to 16-bit pairs and written with
//               stmatrix into the block's W tile.
//   horizontal  per 8 destination columns: ldmatrix of W (16 rows x 32 columns per k-step), PRMT splits the low and
//               high bytes (the k order inside a fragment is {2t, 2t+1, 8+2t, 9+2t}; the planner permutes the
//               coefficient fragments the same way), four mma per k-step (W low/high byte x coefficient low/high
//               plane; the low product starts from the columns' rounding constants), recombination, rounding shift
//               or border division, saturation.  Table addresses run from tile to tile, a tile's descriptor and the
//               operands of a step are fetched one tile / step ahead.
//   output      the block's 16 x stripW result tile is staged in shared memory and leaves as 16- or 8-byte row pieces.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void mmaS8U8(int (&d)[4], const uint4 &a, uint32_t b0, uint32_t b1)
{
    asm("mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]) : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}
// k = 16 form: the B operand is a single register (ldmatrix.m16n16.trans.b8 hands out the two 16-row halves of a
// 32-row k range in registers that are not adjacent, which the k = 32 form would need)
__device__ __forceinline__ void mmaS8U8k16(int (&d)[4], uint32_t a0, uint32_t a1, uint32_t b)
{
    asm("mma.sync.aligned.m16n8k16.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]) : "r"(a0), "r"(a1), "r"(b));
}
// first product of a chain: zero accumulator input (RZ: no register copies to initialise d)
__device__ __forceinline__ void mmaS8U8k16Zero(int (&d)[4], uint32_t a0, uint32_t a1, uint32_t b)
{
    asm("mma.sync.aligned.m16n8k16.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%7,%7,%7,%7};"
        : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]) : "r"(a0), "r"(a1), "r"(b), "r"(0));
}
template <bool SIGNED>
__device__ __forceinline__ void mmaCoefU8k16(int (&d)[4], uint32_t a0, uint32_t a1, uint32_t b, const bool zero)
{
    // A = coefficient bytes (s8 for Lanczos, u8 for Area / Linear), B = source bytes (u8)
    if (SIGNED) {
        if (zero)
            asm("mma.sync.aligned.m16n8k16.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%7,%7,%7,%7};"
                : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]) : "r"(a0), "r"(a1), "r"(b), "r"(0));
        else
            asm("mma.sync.aligned.m16n8k16.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]) : "r"(a0), "r"(a1), "r"(b));
    } else {
        if (zero)
            asm("mma.sync.aligned.m16n8k16.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%7,%7,%7,%7};"
                : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]) : "r"(a0), "r"(a1), "r"(b), "r"(0));
        else
            asm("mma.sync.aligned.m16n8k16.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]) : "r"(a0), "r"(a1), "r"(b));
    }
}
__device__ __forceinline__ uint32_t addU16x2(uint32_t a, uint32_t b)
{
    uint32_t d;
    asm("add.u16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ void mmaU8U8(int (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    asm("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mmaU8S8(int (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    asm("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct MmaKernelArgs {
    alignas(64) CUtensorMap tmap;
    MmaArgs a;
};

__host__ __device__ constexpr int mmaWStride(int wcols) { return 2 * wcols + 16; }          // bytes; (stride / 16) is odd
__host__ __device__ constexpr int mmaOutStride(int stripTiles) { return (8 * stripTiles + 15) / 16 * 16 + 16 + ((((8 * stripTiles + 15) / 16) & 1) ? 16 : 0); }   // multiple of 16, odd number of 16-byte units: the 8 rows of an epilogue store fall on different banks
// shared bytes of a warp: [FIFO chunks | W tile | output tile | strip tables: B fragments, W offsets, {init, divisor} | mbarriers]
__host__ __device__ constexpr int mmaTableBytes(int stripTiles, int hks) { return stripTiles * (hks * 512 + 8 + 64 + 32) + 8; }
constexpr int kMmaChunk = 8;   // plan.hpp kMmaChunkRows

#ifndef IQO_MMA_MINB
#define IQO_MMA_MINB 8
#endif
// VKS / HKS: k-steps of the vertical / horizontal products.  Blocks and tiles that need fewer carry zero coefficient
// fragments for the rest (the planner's tables are zero padded), so the loops have no data-dependent structure.
// The warps of a CTA (a.warps = 1, 2 or 4) share one strip: FIFO, W tile, output tile and tables are common, the
// segments of the vertical pass, the tiles of the horizontal pass and the rows of the store are dealt out round robin;
// two CTA barriers per block separate the phases.
template <int VKS, int HKS, bool SIGNED>
__global__ void __launch_bounds__(128, IQO_MMA_MINB / 2) resizeLanczosMmaKernel(const __grid_constant__ MmaKernelArgs prm)
{
    extern __shared__ __align__(128) uint8_t mmaSmem[];
    const MmaArgs &a = prm.a;
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), nw = blockDim.x >> 5;   // (the shuffle tells the compiler it is warp-uniform)
    const int wv = nw - 1 - warp;   // segments and tiles are dealt out from the last warp down: warp 0, which also requests the rows, gets the short end
    const int strip = blockIdx.x;
    const int T0 = strip * a.stripTiles;                        // first destination tile of the strip
    const int tx0 = 8 * T0;
    const int nt = min(a.stripTiles, (a.DW + 7) / 8 - T0);      // tiles of this strip
    const int tw = min(8 * nt, a.DW - tx0);                     // destination columns of this strip
    const int xs = __ldg(a.stripXs + strip);                    // source column of W element 0
    const int rowBytes = a.wcols;                               // FIFO row
    const int chunkBytes = kMmaChunk * rowBytes;
    const int wStride = mmaWStride(a.wcols), oStride = mmaOutStride(a.stripTiles);
    const uint32_t fifoBase = smemAddr(mmaSmem);
    const uint32_t wBase = fifoBase + a.nChunks * chunkBytes;
    uint8_t *oTile = mmaSmem + a.nChunks * chunkBytes + 16 * wStride;
    const uint32_t oTileS = wBase + 16 * wStride;
    uint8_t *tabs = oTile + 16 * oStride;
    uint4 *sFrag = reinterpret_cast<uint4 *>(tabs);                                        // [tile][HKS][lane]
    int2 *sOff = reinterpret_cast<int2 *>(tabs + a.stripTiles * HKS * 512);                // [tile] {byte offset of the tile's k range in a W row, tile has border columns}
    int *sInit = reinterpret_cast<int *>(sOff + ((a.stripTiles + 1) & ~1));                           // [tile][4 t][4]: rounding constants of columns 2t, 2t+1 as an accumulator quad {c0, c1, c0, c1}
    int *sDiv = sInit + 16 * a.stripTiles;                                                 // [tile * 8] border divisor (0: ordinary column)
    const uint32_t mbarBase = smemAddr(sDiv + 8 * a.stripTiles);
    uint8_t *__restrict__ dst = a.dst + (long long)blockIdx.z * a.dstFrameStride;
    const int nseg = a.wcols >> 4;

    // blocks of this warp (global block indices)
    const int blkFirst = (a.dstRow0 >> 4) + blockIdx.y * a.bandBlocks;
    const int blkEnd = min(blkFirst + a.bandBlocks, (a.dstRow0 + a.dstRows + 15) >> 4);
    if (blkFirst >= blkEnd) return;

    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbarBase));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbarBase + 8));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // chunk c (global source rows 8 c ... 8 c + 7) lives in FIFO slot c mod nChunks (the planner's row map holds the byte
    // offsets that follow from it).  The chunks requested for one block form a group that completes one phase of an
    // mbarrier: group i uses barrier i & 1, phase (i >> 1) & 1.  Warp 0 requests, so only it keeps the request state.
    int2 vb = __ldg(a.vBlock + blkFirst);   // {first source row, rows}
    int cIssued = vb.x >> 3;                // next chunk to request (arithmetic shift: floor)
    uint32_t slotAddr = 0, barIssue = mbarBase;   // where chunk cIssued goes, the barrier of the next group
    if (warp == 0) {
        int sl = cIssued % a.nChunks;
        if (sl < 0) sl += a.nChunks;
        slotAddr = fifoBase + sl * chunkBytes;
    }
    const uint32_t fifoEnd = wBase;
    int grpWaited = 0;                      // groups waited for
    auto issueUpTo = [&](const int cHi) {
        if (warp == 0) {
            const int nNew = cHi + 1 - cIssued;
            if (lane == 0) {
                if (nNew > 0)
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(barIssue), "r"(nNew * chunkBytes) : "memory");
                else
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(barIssue) : "memory");   // nothing new: the phase completes at once
            }
            for (; cIssued <= cHi; ++cIssued) {
                if (lane == 0)
                    asm volatile(
                        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                        ::"r"(slotAddr), "l"(reinterpret_cast<unsigned long long>(&prm)), "r"(xs >> 1),
                          "r"(kMmaChunk * cIssued - a.srcRow0), "r"((int)blockIdx.z), "r"(barIssue)
                        : "memory");
                slotAddr += chunkBytes;
                if (slotAddr == fifoEnd) slotAddr = fifoBase;
            }
            barIssue ^= 8u;
        }
    };
    auto waitGroup = [&]() {
        asm volatile(
            "{\n\t.reg .pred q;\n\tIQO_MMA_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%0], %1;\n\t@!q bra IQO_MMA_WAIT;\n\t}"
            ::"r"(mbarBase + 8 * (grpWaited & 1)), "r"((grpWaited >> 1) & 1)
            : "memory");
        ++grpWaited;
    };
    issueUpTo((vb.x + vb.y - 1) >> 3);   // the first block's rows: in flight while the tables are staged

    // the strip's horizontal tables -> shared memory (every block of the band uses them)
    // (asynchronous 16-byte copies: all of a thread's pieces are in flight at once)
    for (int i = threadIdx.x; i < nt * HKS * 32; i += blockDim.x)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smemAddr(sFrag + i)), "l"(a.hFrag + (size_t)T0 * HKS * 32 + i) : "memory");
    for (int i = threadIdx.x; i < 8 * nt; i += blockDim.x) {
        const int2 hc = __ldg(a.hCol + 8 * T0 + i);   // {init, divisor} of a destination column
        int *q = sInit + 16 * (i >> 3) + 4 * ((i & 7) >> 1) + (i & 1);
        q[0] = q[2] = hc.x;
        sDiv[i] = hc.y;
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    for (int i = threadIdx.x; i < nt; i += blockDim.x) {
        int any = 0;   // tiles that hold border columns divide in the epilogue: a warp-uniform flag per tile
#pragma unroll
        for (int j = 0; j < 8; ++j) any |= sDiv[8 * i + j];
        sOff[i] = make_int2(2 * (__ldg(a.hTile + T0 + i).x - xs), any);
    }
    __syncthreads();

    // lane roles of the matrix loads / stores
    const int mi = lane >> 3, ri = lane & 7;
    const uint32_t wSt = wBase + ((mi & 1) * 8 + ri) * wStride + (mi >> 1) * 16;   // stmatrix: M0/M1 rows 0-7/8-15 of columns 0-7, M2/M3 of columns 8-15
    const uint32_t wLd = wBase + ((mi >> 1) * 8 + ri) * wStride + (mi & 1) * 16;   // ldmatrix: M0/M1 columns 0-7/8-15 of rows 0-7, M2/M3 of rows 8-15
    const int bias = a.workBias;
    // output tile -> global in pieces of 16 bytes (a.dstVec == 2), 8 bytes (1: strips start on multiples of 8 pixels) or
    // single bytes (0); threads per row = the power of two that covers the row's pieces
    const int pieceShift = a.dstVec == 2 ? 4 : a.dstVec == 1 ? 3 : 0;
    const int nch = (tw + (1 << pieceShift) - 1) >> pieceShift;
    int lprShift = 0;
    while ((1 << lprShift) < nch && (1 << lprShift) < (int)blockDim.x) ++lprShift;
    const int stRow = threadIdx.x >> lprShift, stCh = threadIdx.x & ((1 << lprShift) - 1), stRows = (int)blockDim.x >> lprShift;
    // with 16- or 8-byte pieces a thread owns one piece position of the rows stRow, stRow + stRows, ...
    const int stX = stCh << pieceShift;
    const int stKind = pieceShift == 0 ? 0 : stX >= tw ? -1 : (pieceShift == 4 && stX + 16 <= tw) ? 16 : (pieceShift == 3 && stX + 8 <= tw) ? 8 : 1;

    uint4 af[VKS];   // A fragments (coefficients) of the running block
    int rmap[VKS];   // source rows of this lane's k slots
#pragma unroll
    for (int s = 0; s < VKS; ++s) {
        af[s] = __ldg(a.vFrag + ((size_t)blkFirst * VKS + s) * 32 + lane);
        rmap[s] = __ldg(a.vRowMap + ((size_t)blkFirst * VKS + s) * 32 + lane);
    }

    for (int b = blkFirst; b < blkEnd; ++b) {
        // the next block's row range is needed right after this block's vertical pass (by the warp that requests its rows):
        // fetched here, a whole pass ahead, so that nobody waits on the L2 at that point
        if (b + 1 < blkEnd) vb = __ldg(a.vBlock + b + 1);
        waitGroup();   // the rows of this block have landed

        // ---------------- vertical pass ----------------
        uint32_t ra[VKS];   // k slot 32 s + lane reads the FIFO row the planner's map names (unused slots: any resident row, zero coefficients)
#pragma unroll
        for (int s = 0; s < VKS; ++s) ra[s] = fifoBase + (uint32_t)rmap[s];
        const bool borderBlock = SIGNED && ((16 * b < a.mbY) || (16 * b + 16 > a.meY));
        int denoLo = 0, denoHi = 0;
        uint32_t magicLo = 0, magicHi = 0;
        if (borderBlock) {
            const int2 v0 = __ldg(a.vRow + 16 * b + g), v1 = __ldg(a.vRow + 16 * b + g + 8);
            denoLo = v0.x, magicLo = (uint32_t)v0.y, denoHi = v1.x, magicHi = (uint32_t)v1.y;
        }
        // ordinary rows get the bias as a packed 16-bit add on the pair words (every biased value is a non-negative u16);
        // border rows add it inside their division
        const uint32_t biasPair = (uint32_t)bias | ((uint32_t)bias << 16);
        const uint32_t biasLo = denoLo ? 0u : biasPair, biasHi = denoHi ? 0u : biasPair;
        // the B fragments of a segment (16 columns x 32 VKS source rows) are fetched one segment ahead of their mma
        auto loadSeg = [&](uint32_t (&bf)[VKS][4], const int seg) {
#pragma unroll
            for (int s = 0; s < VKS; ++s)
                asm volatile("ldmatrix.sync.aligned.m16n16.x2.trans.shared.b8 {%0, %1, %2, %3}, [%4];"
                             : "=r"(bf[s][0]), "=r"(bf[s][1]), "=r"(bf[s][2]), "=r"(bf[s][3]) : "r"(ra[s] + 16 * seg));
        };
        auto computeSeg = [&](const uint32_t (&bf)[VKS][4], const int seg) {
            int dA[4], dB[4];
#pragma unroll
            for (int s = 0; s < VKS; ++s) {
                mmaCoefU8k16<SIGNED>(dA, af[s].x, af[s].y, bf[s][0], s == 0);   // columns 0..7 of the segment, k slots 32 s ... + 15
                mmaCoefU8k16<SIGNED>(dB, af[s].x, af[s].y, bf[s][1], s == 0);   // columns 8..15
                mmaCoefU8k16<SIGNED>(dA, af[s].z, af[s].w, bf[s][2], false);    // k slots 32 s + 16 ... + 31
                mmaCoefU8k16<SIGNED>(dB, af[s].z, af[s].w, bf[s][3], false);
            }
            if (borderBlock) {
                // resizeYborder: int16 numerator * 64 / denominator, C (truncating) division (see halfVerticalStrip)
                auto bdiv = [&](int x, int deno, uint32_t magic) -> int {
                    if (!deno) return x;
                    const int n = (int)(short)x * 64;
                    const uint32_t mm = (uint32_t)abs(n);
                    const int q = magic ? (int)__umulhi(mm, magic) : (int)mm;
                    return (int)(short)(n < 0 ? -q : q) + bias;
                };
                dA[0] = bdiv(dA[0], denoLo, magicLo), dA[1] = bdiv(dA[1], denoLo, magicLo);
                dB[0] = bdiv(dB[0], denoLo, magicLo), dB[1] = bdiv(dB[1], denoLo, magicLo);
                dA[2] = bdiv(dA[2], denoHi, magicHi), dA[3] = bdiv(dA[3], denoHi, magicHi);
                dB[2] = bdiv(dB[2], denoHi, magicHi), dB[3] = bdiv(dB[3], denoHi, magicHi);
            }
            const uint32_t w0 = addU16x2(prmt((uint32_t)dA[0], (uint32_t)dA[1], 0x5410), biasLo), w1 = addU16x2(prmt((uint32_t)dA[2], (uint32_t)dA[3], 0x5410), biasHi);
            const uint32_t w2 = addU16x2(prmt((uint32_t)dB[0], (uint32_t)dB[1], 0x5410), biasLo), w3 = addU16x2(prmt((uint32_t)dB[2], (uint32_t)dB[3], 0x5410), biasHi);
            asm volatile("stmatrix.sync.aligned.m8n8.x4.shared.b16 [%0], {%1, %2, %3, %4};" ::"r"(wSt + 32 * seg), "r"(w0), "r"(w1), "r"(w2), "r"(w3) : "memory");
        };
        if (wv < nseg) {
            // this warp's segments: warp, warp + nw, ...
            uint32_t bf0[VKS][4], bf1[VKS][4];
            loadSeg(bf0, wv);
            int seg = wv;
            for (; seg + nw < nseg; seg += 2 * nw) {
                loadSeg(bf1, seg + nw);
                computeSeg(bf0, seg);
                if (seg + 2 * nw < nseg) loadSeg(bf0, seg + 2 * nw);
                computeSeg(bf1, seg + nw);
            }
            if (seg < nseg) computeSeg(bf0, seg);
        }
        __syncthreads();   // W is complete; every warp has finished reading the source rows of this block
        if (b + 1 < blkEnd) {
            // the next block's rows and coefficient fragments arrive during the horizontal pass
            issueUpTo((vb.x + vb.y - 1) >> 3);
#pragma unroll
            for (int s = 0; s < VKS; ++s) {
                af[s] = __ldg(a.vFrag + ((size_t)(b + 1) * VKS + s) * 32 + lane);
                rmap[s] = __ldg(a.vRowMap + ((size_t)(b + 1) * VKS + s) * 32 + lane);
            }
        }

        // ---------------- horizontal pass: tiles in pairs, the operands of a step fetched one step ahead ----------------
        {
            struct Step {
                uint32_t r[8];
                uint4 bf;
            };
            // tile ti of the strip: W offset + border flag at offS + 8 ti, B fragments at fragS + 512 HKS ti, rounding quad
            // at initS + 64 ti, output bytes at outS + 8 ti.  This warp's tiles are warp, warp + nw, ...: the addresses run
            auto loadStep = [&](Step &st, const int wOff, const uint32_t frag, const int s) {
                const uint32_t wl = wLd + wOff + 64 * s;
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                             : "=r"(st.r[0]), "=r"(st.r[1]), "=r"(st.r[2]), "=r"(st.r[3]) : "r"(wl));
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                             : "=r"(st.r[4]), "=r"(st.r[5]), "=r"(st.r[6]), "=r"(st.r[7]) : "r"(wl + 32));
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(st.bf.x), "=r"(st.bf.y), "=r"(st.bf.z), "=r"(st.bf.w) : "r"(frag + 512 * s));
            };
            auto ldsInt2 = [](const uint32_t addr) -> int2 {
                int2 v;
                asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
                return v;
            };
            int ll[4], mid[4], hh[4];
            auto computeStep = [&](const Step &st, const bool first, const uint32_t init) {
                if (first) {
                    // the low product starts from the columns' rounding constants (one 16-byte load straight into the
                    // accumulator quad), the other two from zero
                    asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(ll[0]), "=r"(ll[1]), "=r"(ll[2]), "=r"(ll[3]) : "r"(init));
#pragma unroll
                    for (int e = 0; e < 4; ++e) mid[e] = hh[e] = 0;
                }
                uint32_t alo[4], ahi[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    alo[i] = prmt(st.r[2 * i], st.r[2 * i + 1], 0x6420);
                    ahi[i] = prmt(st.r[2 * i], st.r[2 * i + 1], 0x7531);
                }
                // coefficient high plane: signed for Lanczos, 0 .. 128 for Area / Linear
                if (SIGNED) mmaU8S8(mid, alo, st.bf.z, st.bf.w); else mmaU8U8(mid, alo, st.bf.z, st.bf.w);
                mmaU8U8(ll, alo, st.bf.x, st.bf.y);
                if (SIGNED) mmaU8S8(hh, ahi, st.bf.z, st.bf.w); else mmaU8U8(hh, ahi, st.bf.z, st.bf.w);
                mmaU8U8(mid, ahi, st.bf.x, st.bf.y);
            };
            auto finishTile = [&](const int ti, const int border, const uint32_t out) {
                // thread (g, t): rows g, g + 8; columns 8 (T0 + ti) + 2 t, + 1
                int v[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) v[e] = ll[e] + (mid[e] << 8) + (hh[e] << 16);
                if (!SIGNED) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) v[e] = (int)(short)(v[e] >> 23);   // Area / Linear: 8 + 15 fixed-point bits
                } else if (border == 0) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) v[e] = (int)(short)(v[e] >> 20);
                } else {
                    const int2 dv2 = *reinterpret_cast<const int2 *>(sDiv + 8 * ti + 2 * t);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int dv = (e & 1) ? dv2.y : dv2.x;
                        v[e] = (int)(short)(dv != 0 ? v[e] / dv : v[e] >> 20);   // resizeXborder: truncating division by deno * 64
                    }
                }
                const uint32_t p01 = packSatU8(v[1], v[0], 0u), p23 = packSatU8(v[3], v[2], 0u);
                asm volatile("st.shared.u16 [%0], %1;" ::"r"(out), "h"((uint16_t)p01) : "memory");
                asm volatile("st.shared.u16 [%0], %1;" ::"r"(out + 8 * oStride), "h"((uint16_t)p23) : "memory");
            };
            // steps are numbered i = (tile sequence) * HKS + s; a pair of tiles is 2 HKS steps, so the two operand buffers
            // alternate at compile-time positions inside the unrolled pair
            const uint32_t dOff = 8 * nw, dFrag = 512 * HKS * nw, dInit = 64 * nw, dOut = 8 * nw;
            uint32_t aOff = smemAddr(sOff) + 8 * wv, aFrag = smemAddr(sFrag) + 512 * HKS * wv + 16 * lane;
            uint32_t aInit = smemAddr(sInit) + 64 * wv + 16 * t, aOut = oTileS + g * oStride + 8 * wv + 2 * t;
            Step st[2];
            int2 odCur = make_int2(0, 0), odNext = make_int2(0, 0);   // {W offset, border flag} of the running tile and the one after it
            if (wv < nt) {
                odCur = ldsInt2(aOff);
                loadStep(st[0], odCur.x, aFrag, 0);
                if (wv + nw < nt) odNext = ldsInt2(aOff + dOff);
            }
            for (int tp = wv; tp < nt; tp += 2 * nw) {
#pragma unroll
                for (int u = 0; u < 2 * HKS; ++u) {
                    const int ti = tp + (u / HKS) * nw, s = u % HKS;
                    if (ti < nt) {
                        if (s + 1 < HKS) {
                            loadStep(st[(u + 1) & 1], odCur.x, aFrag, s + 1);
                        } else if (ti + nw < nt) {
                            loadStep(st[(u + 1) & 1], odNext.x, aFrag + dFrag, 0);
                        }
                        computeStep(st[u & 1], s == 0, aInit);
                        if (s == HKS - 1) {
                            finishTile(ti, odCur.y, aOut);
                            odCur = odNext;
                            aOff += dOff, aFrag += dFrag, aInit += dInit, aOut += dOut;
                            if (ti + 2 * nw < nt) odNext = ldsInt2(aOff + dOff);
                        }
                    }
                }
            }
        }
        __syncthreads();

        // ---------------- store the 16 x tw tile ----------------
        {
            const int yb = 16 * b;
            const int rLo = max(a.dstRow0 - yb, 0), rHi = min(a.dstRow0 + a.dstRows - yb, 16);   // rows of the block inside the launch
            uint8_t *drow = dst + (long long)(yb + stRow - a.dstRow0) * a.dstPitch + tx0;
            const long long stStep = (long long)stRows * a.dstPitch;
            if (stKind > 0) {
                uint8_t *d = drow + stX;
                uint32_t sp = oTileS + stRow * oStride + stX;
                const bool allRows = rLo == 0 && rHi == 16;   // every block but the first and last of a row band
                if (stKind == 16) {
                    for (int r = stRow; r < 16; r += stRows, d += stStep, sp += stRows * oStride)
                        if (allRows || (r >= rLo && r < rHi)) {
                            uint4 v;
                            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(sp));
                            *reinterpret_cast<uint4 *>(d) = v;
                        }
                } else if (stKind == 8) {
                    for (int r = stRow; r < 16; r += stRows, d += stStep, sp += stRows * oStride)
                        if (allRows || (r >= rLo && r < rHi)) {
                            uint2 v;
                            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(sp));
                            *reinterpret_cast<uint2 *>(d) = v;
                        }
                } else {
                    // the ragged piece at the end of the last strip
                    for (int r = stRow; r < 16; r += stRows, d += stStep)
                        if (r >= rLo && r < rHi)
                            for (int i = 0; i < min(1 << pieceShift, tw - stX); ++i) d[i] = oTile[(r * oStride + stX) + i];
                }
            } else if (stKind == 0) {
                // single bytes (unaligned destination): a row may have more bytes than the CTA has threads
                for (int r = stRow; r < 16; r += stRows, drow += stStep) {
                    if (r < rLo || r >= rHi) continue;
                    const uint8_t *srow = oTile + r * oStride;
                    for (int x = stCh; x < tw; x += 1 << lprShift) drow[x] = srow[x];
                }
            }
        }
        // (the next block's first barrier orders these reads of the output tile before its horizontal pass writes it;
        //  W was last read before the barrier above, so the next vertical pass may overwrite it)
    }
}


// ---------------------------------------------------------------------------------------
// Tensor-path vertical pass + compile-time dp2a horizontal pass, for the reductions of the RD | 8 family (cfg1: 3:2).
// Geometry, source FIFO (TMA chunks, mbarrier groups) and vertical pass are resizeLanczosMmaKernel's; the horizontal
// pass is resizeRatioStreamKernel's: an item is (row of the 16-row W tile, group of 8 destination pixels), the pair
// words of the group's window come from the W row with 8-byte loads, the coefficient words of each (phase, parity) are
// constant-bank operands, and the 8 pixels leave as one 8-byte store straight from registers -- no output tile, no
// store phase, no horizontal tables in shared memory.  Border columns (first / last strip) are recomputed from the
// W rows with the generic tables.  ODD: the first tap of a group sits on an odd W element (the pair words start one
// element earlier and every pixel's offset grows by one).  Needs DW % 8 == 0.
// ---------------------------------------------------------------------------------------
struct MmaRatioKernelArgs {
    alignas(64) CUtensorMap tmap;
    MmaArgs a;
    uint32_t cwX[4][2][7];     // [phase][parity][pair word] (RatioArgs)
    int accInit;               // 2^19 - workBias * sum of the low plane ... (start of the low-plane sum)
    int c0;                    // first tap of destination pixel 0
    int dstVec8;               // destination rows may be written with 8-byte stores
    AxisDev gx;                // generic horizontal tables: border columns
};

#ifndef IQO_MMAR_MINB
#define IQO_MMAR_MINB 4
#endif
// pair words the 8 pixels of a group span (even count: 8-byte loads), ODD: the first tap sits on an odd W element
template <int RS, int RD, int NXE, int ODD>
__host__ __device__ constexpr int mmaRatioWords()
{
    int m = 0;
    for (int p = 0; p < 8; ++p) {
        const int off = (p * RS) / RD + ODD;
        const int e = (off >> 1) + (NXE + (off & 1) + 1) / 2;
        m = e > m ? e : m;
    }
    return (m + 1) & ~1;
}

template <int VKS, int RS, int RD, int NX, int TZ, int ODD>
__global__ void __launch_bounds__(128, IQO_MMAR_MINB) resizeLanczosMmaRatioKernel(const __grid_constant__ MmaRatioKernelArgs prm)
{
    extern __shared__ __align__(128) uint8_t mmaSmem[];
    constexpr bool SIGNED = true;
    constexpr int GS = 8 * RS / RD;
    const MmaArgs &a = prm.a;
    const int lane = threadIdx.x & 31, g = lane >> 2;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), nw = blockDim.x >> 5;
    const int wv = nw - 1 - warp;
    const int strip = blockIdx.x;
    const int T0 = strip * a.stripTiles;
    const int tx0 = 8 * T0;
    const int ngs = min(a.stripTiles, (a.DW >> 3) - T0);       // 8-pixel groups of this strip
    const int xs = __ldg(a.stripXs + strip);                    // source column of W element 0
    const int rowBytes = a.wcols;
    const int chunkBytes = kMmaChunk * rowBytes;
    const int wStride = mmaWStride(a.wcols);
    const uint32_t fifoBase = smemAddr(mmaSmem);
    const uint32_t wBase = fifoBase + a.nChunks * chunkBytes;
    const uint32_t mbarBase = wBase + 16 * wStride;
    uint8_t *__restrict__ dst = a.dst + (long long)blockIdx.z * a.dstFrameStride;
    const int nseg = a.wcols >> 4;
    const int i0 = GS * T0 + prm.c0 - xs - ODD;                 // W element the pair words of pixel tx0 start at (even; ODD: one before its first tap)

    const int blkFirst = (a.dstRow0 >> 4) + blockIdx.y * a.bandBlocks;
    const int blkEnd = min(blkFirst + a.bandBlocks, (a.dstRow0 + a.dstRows + 15) >> 4);
    if (blkFirst >= blkEnd) return;

    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbarBase));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbarBase + 8));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    int2 vb = __ldg(a.vBlock + blkFirst);   // {first source row, rows}
    int cIssued = vb.x >> 3;
    uint32_t slotAddr = 0, barIssue = mbarBase;
    if (warp == 0) {
        int sl = cIssued % a.nChunks;
        if (sl < 0) sl += a.nChunks;
        slotAddr = fifoBase + sl * chunkBytes;
    }
    const uint32_t fifoEnd = wBase;
    int grpWaited = 0;
    auto issueUpTo = [&](const int cHi) {
        if (warp == 0) {
            const int nNew = cHi + 1 - cIssued;
            if (lane == 0) {
                if (nNew > 0)
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(barIssue), "r"(nNew * chunkBytes) : "memory");
                else
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(barIssue) : "memory");
            }
            for (; cIssued <= cHi; ++cIssued) {
                if (lane == 0)
                    asm volatile(
                        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                        ::"r"(slotAddr), "l"(reinterpret_cast<unsigned long long>(&prm)), "r"(xs >> 1),
                          "r"(kMmaChunk * cIssued - a.srcRow0), "r"((int)blockIdx.z), "r"(barIssue)
                        : "memory");
                slotAddr += chunkBytes;
                if (slotAddr == fifoEnd) slotAddr = fifoBase;
            }
            barIssue ^= 8u;
        }
    };
    auto waitGroup = [&]() {
        asm volatile(
            "{\n\t.reg .pred q;\n\tIQO_MMAR_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%0], %1;\n\t@!q bra IQO_MMAR_WAIT;\n\t}"
            ::"r"(mbarBase + 8 * (grpWaited & 1)), "r"((grpWaited >> 1) & 1)
            : "memory");
        ++grpWaited;
    };
    issueUpTo((vb.x + vb.y - 1) >> 3);

    const int mi = lane >> 3, ri = lane & 7;
    const uint32_t wSt = wBase + ((mi & 1) * 8 + ri) * wStride + (mi >> 1) * 16;
    const int bias = a.workBias;
    const uint32_t rcp = (65536u + ngs - 1) / ngs;   // item / ngs for item < 16 * 34
    const bool left = tx0 < a.mbX, right = tx0 + 8 * ngs > a.meX;

    uint4 af[VKS];
    int rmap[VKS];
#pragma unroll
    for (int s = 0; s < VKS; ++s) {
        af[s] = __ldg(a.vFrag + ((size_t)blkFirst * VKS + s) * 32 + lane);
        rmap[s] = __ldg(a.vRowMap + ((size_t)blkFirst * VKS + s) * 32 + lane);
    }

    for (int b = blkFirst; b < blkEnd; ++b) {
        if (b + 1 < blkEnd) vb = __ldg(a.vBlock + b + 1);
        waitGroup();

        // ---------------- vertical pass (as in resizeLanczosMmaKernel) ----------------
        uint32_t ra[VKS];
#pragma unroll
        for (int s = 0; s < VKS; ++s) ra[s] = fifoBase + (uint32_t)rmap[s];
        const bool borderBlock = (16 * b < a.mbY) || (16 * b + 16 > a.meY);
        int denoLo = 0, denoHi = 0;
        uint32_t magicLo = 0, magicHi = 0;
        if (borderBlock) {
            const int2 v0 = __ldg(a.vRow + 16 * b + g), v1 = __ldg(a.vRow + 16 * b + g + 8);
            denoLo = v0.x, magicLo = (uint32_t)v0.y, denoHi = v1.x, magicHi = (uint32_t)v1.y;
        }
        const uint32_t biasPair = (uint32_t)bias | ((uint32_t)bias << 16);
        const uint32_t biasLo = denoLo ? 0u : biasPair, biasHi = denoHi ? 0u : biasPair;
        auto loadSeg = [&](uint32_t (&bf)[VKS][4], const int seg) {
#pragma unroll
            for (int s = 0; s < VKS; ++s)
                asm volatile("ldmatrix.sync.aligned.m16n16.x2.trans.shared.b8 {%0, %1, %2, %3}, [%4];"
                             : "=r"(bf[s][0]), "=r"(bf[s][1]), "=r"(bf[s][2]), "=r"(bf[s][3]) : "r"(ra[s] + 16 * seg));
        };
        auto computeSeg = [&](const uint32_t (&bf)[VKS][4], const int seg) {
            int dA[4], dB[4];
#pragma unroll
            for (int s = 0; s < VKS; ++s) {
                mmaCoefU8k16<SIGNED>(dA, af[s].x, af[s].y, bf[s][0], s == 0);
                mmaCoefU8k16<SIGNED>(dB, af[s].x, af[s].y, bf[s][1], s == 0);
                mmaCoefU8k16<SIGNED>(dA, af[s].z, af[s].w, bf[s][2], false);
                mmaCoefU8k16<SIGNED>(dB, af[s].z, af[s].w, bf[s][3], false);
            }
            if (borderBlock) {
                auto bdiv = [&](int x, int deno, uint32_t magic) -> int {
                    if (!deno) return x;
                    const int n = (int)(short)x * 64;
                    const uint32_t mm = (uint32_t)abs(n);
                    const int q = magic ? (int)__umulhi(mm, magic) : (int)mm;
                    return (int)(short)(n < 0 ? -q : q) + bias;
                };
                dA[0] = bdiv(dA[0], denoLo, magicLo), dA[1] = bdiv(dA[1], denoLo, magicLo);
                dB[0] = bdiv(dB[0], denoLo, magicLo), dB[1] = bdiv(dB[1], denoLo, magicLo);
                dA[2] = bdiv(dA[2], denoHi, magicHi), dA[3] = bdiv(dA[3], denoHi, magicHi);
                dB[2] = bdiv(dB[2], denoHi, magicHi), dB[3] = bdiv(dB[3], denoHi, magicHi);
            }
            const uint32_t w0 = addU16x2(prmt((uint32_t)dA[0], (uint32_t)dA[1], 0x5410), biasLo), w1 = addU16x2(prmt((uint32_t)dA[2], (uint32_t)dA[3], 0x5410), biasHi);
            const uint32_t w2 = addU16x2(prmt((uint32_t)dB[0], (uint32_t)dB[1], 0x5410), biasLo), w3 = addU16x2(prmt((uint32_t)dB[2], (uint32_t)dB[3], 0x5410), biasHi);
            asm volatile("stmatrix.sync.aligned.m8n8.x4.shared.b16 [%0], {%1, %2, %3, %4};" ::"r"(wSt + 32 * seg), "r"(w0), "r"(w1), "r"(w2), "r"(w3) : "memory");
        };
        if (wv < nseg) {
            uint32_t bf0[VKS][4], bf1[VKS][4];
            loadSeg(bf0, wv);
            int seg = wv;
            for (; seg + nw < nseg; seg += 2 * nw) {
                loadSeg(bf1, seg + nw);
                computeSeg(bf0, seg);
                if (seg + 2 * nw < nseg) loadSeg(bf0, seg + 2 * nw);
                computeSeg(bf1, seg + nw);
            }
            if (seg < nseg) computeSeg(bf0, seg);
        }
        __syncthreads();   // W is complete; the source rows of this block are free
        if (b + 1 < blkEnd) {
            issueUpTo((vb.x + vb.y - 1) >> 3);
#pragma unroll
            for (int s = 0; s < VKS; ++s) {
                af[s] = __ldg(a.vFrag + ((size_t)(b + 1) * VKS + s) * 32 + lane);
                rmap[s] = __ldg(a.vRowMap + ((size_t)(b + 1) * VKS + s) * 32 + lane);
            }
        }

        // ---------------- horizontal pass: (row, group of 8 pixels) items, compile-time tap pattern ----------------
        const int yb = 16 * b;
        const int rLo = max(a.dstRow0 - yb, 0), rHi = min(a.dstRow0 + a.dstRows - yb, 16);   // rows of the block inside the launch
        for (int item = threadIdx.x; item < 16 * ngs; item += blockDim.x) {
            const int r = (int)(((uint32_t)item * rcp) >> 16);
            const int q = item - r * ngs;
            if (r < rLo || r >= rHi) continue;
            const uint32_t wl = wBase + r * wStride + 2 * (i0 + GS * q);
            constexpr int NXE = NX - TZ;  // taps left after the zero taps that end every phase
            constexpr int kWords = mmaRatioWords<RS, RD, NXE, ODD>();
            uint32_t n[kWords];
            if ((wl & 7) == 0) {
#pragma unroll
                for (int j = 0; j < kWords; j += 2)
                    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(n[j]), "=r"(n[j + 1]) : "r"(wl + 4 * j) : "memory");
            } else {
#pragma unroll
                for (int j = 0; j < kWords; ++j) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(n[j]) : "r"(wl + 4 * j) : "memory");
            }
            int px[8];
#pragma unroll
            for (int p = 0; p < 8; ++p) {
                const int off = (p * RS) / RD + ODD;
                const int wp = off >> 1, par = off & 1, ph = p % RD;
                const int nwd = (NXE + par + 1) / 2;
                int acc = prm.accInit;
#pragma unroll
                for (int j = 0; j < nwd; ++j) acc = dp2a_lo_uu(n[wp + j], prm.cwX[ph][par][j], acc);
                acc >>= 8;
#pragma unroll
                for (int j = 0; j < nwd; ++j) acc = dp2a_hi_us(n[wp + j], prm.cwX[ph][par][j], acc);
                px[p] = acc >> 12;
            }
            uint2 o2;
            o2.x = packSatU8(px[1], px[0], packSatU8(px[3], px[2], 0u));
            o2.y = packSatU8(px[5], px[4], packSatU8(px[7], px[6], 0u));
            uint8_t *out = dst + (long long)(yb + r - a.dstRow0) * a.dstPitch + tx0 + 8 * q;
            if (prm.dstVec8)
                *reinterpret_cast<uint2 *>(out) = o2;
            else
                halfStoreBytes(out, o2, 0, 8);
        }
        if (left || right) {
            __syncthreads();  // the main stores of these pixels come first
            const int c0b = left ? tx0 : max(a.meX, tx0);
            const int c1b = left ? min(a.mbX, tx0 + 8 * ngs) : tx0 + 8 * ngs;
            for (int side = 0; side < 2; ++side) {
                int b0 = c0b, b1 = c1b;
                if (side == 1) {
                    if (!(left && right)) break;
                    b0 = max(a.meX, max(tx0, a.mbX));
                    b1 = tx0 + 8 * ngs;
                }
                const int nb = b1 - b0;
                for (int item = threadIdx.x; item < 16 * nb; item += blockDim.x) {
                    // resizeXborder from the W row: masked taps, truncating division (finishPixel)
                    const int r = item / nb, d = b0 + item - r * nb;
                    if (r < rLo || r >= rHi) continue;
                    const int fx = __ldg(prm.gx.first + d), rx = __ldg(prm.gx.row + d);
                    const int32_t *cx = prm.gx.coef + rx * NX;
                    const uint32_t wr = wBase + r * wStride + 2 * (fx - xs);
                    int nume = 0;
#pragma unroll
                    for (int i = 0; i < NX; ++i) {
                        const int cxi = __ldg(cx + i);
                        if (cxi == 0) continue;   // masked taps may lie outside the W row
                        uint32_t w16;
                        asm volatile("ld.shared.u16 %0, [%1];" : "=r"(w16) : "r"(wr + 2 * i) : "memory");
                        nume += ((int)w16 - bias) * cxi;
                    }
                    dst[(long long)(yb + r - a.dstRow0) * a.dstPitch + d] = finishPixel(nume, __ldg(prm.gx.deno + rx), 20);
                }
            }
        }
        __syncthreads();   // the next vertical pass may overwrite W
    }
}

}  // namespace

size_t mmaSmemBytes(int wcols, int stripTiles, int nChunks, int hKMax)
{
    return size_t(nChunks) * kMmaChunk * wcols + 16 * size_t(mmaWStride(wcols)) + 16 * size_t(mmaOutStride(stripTiles)) +
           size_t(mmaTableBytes(stripTiles, hKMax)) + 8 * size_t(nChunks) + 16;
}

template <int VKS, int HKS, bool SIGNED>
cudaError_t launchMmaT(const MmaKernelArgs &p, dim3 grid, size_t smem, cudaStream_t stream)
{
    static PerDeviceOnce attrSet;
    const int dev = currentDevice();
    if (!attrSet.done(dev)) {
        cudaError_t e = cudaFuncSetAttribute(resizeLanczosMmaKernel<VKS, HKS, SIGNED>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        attrSet.set(dev);
    }
    resizeLanczosMmaKernel<VKS, HKS, SIGNED><<<grid, 32 * p.a.warps, smem, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launchMma(const MmaArgs &a, const CUtensorMap &tmap, cudaStream_t stream)
{
    const int tiles = (a.DW + 7) / 8;
    const int strips = (tiles + a.stripTiles - 1) / a.stripTiles;
    const int blocks = ((a.dstRow0 + a.dstRows + 15) >> 4) - (a.dstRow0 >> 4);
    const int bands = (blocks + a.bandBlocks - 1) / a.bandBlocks;
    if (bands > 65535 || a.nFrames > 65535) return cudaErrorInvalidConfiguration;
    const size_t smem = mmaSmemBytes(a.wcols, a.stripTiles, a.nChunks, a.hKMax);
    MmaKernelArgs p;
    p.tmap = tmap;
    p.a = a;
    dim3 grid(strips, bands, a.nFrames);
    g_launches.fetch_add(1);
#define IQO_MMA_CASE(V, H) \
    if (a.vKMax == V && a.hKMax == H) return a.isSigned ? launchMmaT<V, H, true>(p, grid, smem, stream) : launchMmaT<V, H, false>(p, grid, smem, stream);
    IQO_MMA_CASE(1, 1) IQO_MMA_CASE(1, 2) IQO_MMA_CASE(1, 3)
    IQO_MMA_CASE(2, 1) IQO_MMA_CASE(2, 2) IQO_MMA_CASE(2, 3)
    IQO_MMA_CASE(3, 1) IQO_MMA_CASE(3, 2) IQO_MMA_CASE(3, 3)
#undef IQO_MMA_CASE
    return cudaErrorInvalidValue;
}

size_t mmaRatioSmemBytes(int wcols, int nChunks)
{
    return size_t(nChunks) * kMmaChunk * wcols + 16 * size_t(mmaWStride(wcols)) + 16;
}

bool mmaRatioHasKernel(int RS, int RD, int NX, int odd)
{
    // measured in one run (512 frames 1080p): Lanczos3 at 3:2 0.557 ms here, 0.61 on the all-mma kernel, 0.67 on the 3:2
    // kernel; the 12-tap patterns (Lanczos4 at 3:2: 0.72 vs 0.69, 2:1 on X only: 0.66 vs 0.56) are faster all-mma
    return odd == 0 && RS == 3 && RD == 2 && NX == 10;
}

template <int VKS, int RS, int RD, int NX, int TZ, int ODD>
cudaError_t launchMmaRatioT(const MmaRatioKernelArgs &p, dim3 grid, size_t smem, cudaStream_t stream)
{
    static PerDeviceOnce attrSet;
    const int dev = currentDevice();
    if (!attrSet.done(dev)) {
        cudaError_t e = cudaFuncSetAttribute(resizeLanczosMmaRatioKernel<VKS, RS, RD, NX, TZ, ODD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        attrSet.set(dev);
    }
    resizeLanczosMmaRatioKernel<VKS, RS, RD, NX, TZ, ODD><<<grid, 32 * p.a.warps, smem, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launchMmaRatio(const MmaArgs &a, const CUtensorMap &tmap, const RatioArgs &ra, cudaStream_t stream)
{
    if (!mmaRatioHasKernel(ra.RS, ra.RD, ra.NX, ra.odd) || !a.isSigned || (a.DW & 7)) return cudaErrorInvalidValue;
    const int tiles = a.DW / 8;
    const int strips = (tiles + a.stripTiles - 1) / a.stripTiles;
    const int blocks = ((a.dstRow0 + a.dstRows + 15) >> 4) - (a.dstRow0 >> 4);
    const int bands = (blocks + a.bandBlocks - 1) / a.bandBlocks;
    if (bands > 65535 || a.nFrames > 65535) return cudaErrorInvalidConfiguration;
    const size_t smem = mmaRatioSmemBytes(a.wcols, a.nChunks);
    MmaRatioKernelArgs p;
    p.tmap = tmap;
    p.a = a;
    memcpy(p.cwX, ra.cwX, sizeof p.cwX);
    // the low-plane sum starts at 2^19 minus the bias share of the whole sum: RatioPlan's accInit is stated for its own
    // bias, the same formula with this plan's bias (the horizontal weights of a main column sum to 16384)
    p.accInit = (1 << 19) - a.workBias * 16384;
    p.c0 = ra.c0;
    p.dstVec8 = a.dstVec >= 1;
    p.gx = ra.gx;
    dim3 grid(strips, bands, a.nFrames);
    g_launches.fetch_add(1);
#define IQO_MMAR_CASE(V) \
    if (a.vKMax == V) return ra.tailZeros ? launchMmaRatioT<V, 3, 2, 10, 1, 0>(p, grid, smem, stream) : launchMmaRatioT<V, 3, 2, 10, 0, 0>(p, grid, smem, stream);
    IQO_MMAR_CASE(1) IQO_MMAR_CASE(2) IQO_MMAR_CASE(3)
#undef IQO_MMAR_CASE
    return cudaErrorInvalidValue;
}

GenericGeom chooseGenericGeom(const int32_t *firstX, int N, int S, int D)
{
    GenericGeom g;
    g.tileW = 128;
    g.tileH = 16;
    for (;;) {
        int widest = 1;
        for (int t0 = 0; t0 < D; t0 += g.tileW) {
            int t1 = std::min(D, t0 + g.tileW) - 1;
            int lo = std::min(std::max(firstX[t0], 0), S - 1);
            int hi = std::min(std::max(firstX[t1] + N - 1, 0), S - 1);
            widest = std::max(widest, hi - lo + 1);
        }
        g.workW = (widest + 7) & ~7;
        g.smemBytes = size_t(g.tileH) * g.workW * 2;
        if (g.smemBytes <= 96 * 1024) break;
        // extreme down-sampling ratios: narrower tile first, then fewer rows (capi.cu rejects the shape at
        // create time if even an 8 x 1 tile does not fit the 200 KB opt-in limit)
        if (g.tileW > 8)
            g.tileW /= 2;
        else if (g.tileH > 1)
            g.tileH /= 2;
        else
            break;
    }
    return g;
}

cudaError_t initKernels()
{
    cudaError_t e = cudaFuncSetAttribute(resizeGenericKernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(resizeGenericKernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
}

cudaError_t launchGeneric(const ResizeArgs &a, const GenericGeom &g, cudaStream_t stream)
{
    const int tilesX = (a.x.D + g.tileW - 1) / g.tileW;
    const int tilesY = (a.dstRows + g.tileH - 1) / g.tileH;
    if (tilesY > 65535) return cudaErrorInvalidConfiguration;
    // gridDim.z is limited to 65535 frames per launch
    for (int f0 = 0; f0 < a.nFrames; f0 += 65535) {
        ResizeArgs b = a;
        b.nFrames = std::min(65535, a.nFrames - f0);
        b.src = a.src + (long long)f0 * a.srcFrameStride;
        b.dst = a.dst + (long long)f0 * a.dstFrameStride;
        dim3 grid(tilesX, tilesY, b.nFrames);
        if (a.workSigned)
            resizeGenericKernel<true><<<grid, 256, g.smemBytes, stream>>>(b, g.tileW, g.tileH, g.workW);
        else
            resizeGenericKernel<false><<<grid, 256, g.smemBytes, stream>>>(b, g.tileW, g.tileH, g.workW);
        g_launches.fetch_add(1);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

cudaError_t launchFloat(const FloatArgs &a, const GenericGeom &g, cudaStream_t stream)
{
    static PerDeviceOnce attrSet;
    const int dev = currentDevice();
    if (!attrSet.done(dev)) {
        cudaError_t e = cudaFuncSetAttribute(resizeFloatKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        attrSet.set(dev);
    }
    // the float work tile needs twice the bytes of the 16-bit one: halve the tile height when it would not fit
    int tileH = g.tileH;
    while (tileH > 1 && size_t(tileH) * g.workW * 4 > 200 * 1024) tileH /= 2;
    const size_t smem = size_t(tileH) * g.workW * 4;
    if (smem > 200 * 1024) return cudaErrorInvalidConfiguration;
    const int tilesX = (a.x.D + g.tileW - 1) / g.tileW;
    const int tilesY = (a.dstRows + tileH - 1) / tileH;
    if (tilesY > 65535) return cudaErrorInvalidConfiguration;
    for (int f0 = 0; f0 < a.nFrames; f0 += 65535) {
        FloatArgs b = a;
        b.nFrames = std::min(65535, a.nFrames - f0);
        b.src = a.src + (long long)f0 * a.srcFrameStride;
        b.dst = a.dst + (long long)f0 * a.dstFrameStride;
        dim3 grid(tilesX, tilesY, b.nFrames);
        resizeFloatKernel<<<grid, 256, smem, stream>>>(b, g.tileW, tileH, g.workW);
        g_launches.fetch_add(1);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

cudaError_t launchHalf(const HalfArgs &a, const CUtensorMap *tmap, int boxRows, cudaStream_t stream)
{
#define IQO_HALF_CASE(G, NW)                                                                           \
    if (a.NG == G && a.NWX == NW)                                                                      \
        return !a.symmetric ? launchHalfT<G, NW, false, true>(a, tmap, boxRows, stream)                \
               : a.endsHi   ? launchHalfT<G, NW, true, true>(a, tmap, boxRows, stream)                 \
                            : launchHalfT<G, NW, true, false>(a, tmap, boxRows, stream);
    IQO_HALF_CASE(3, 7)
    IQO_HALF_CASE(3, 5)
    IQO_HALF_CASE(2, 3)
    IQO_HALF_CASE(2, 5)
    IQO_HALF_CASE(2, 7)
    IQO_HALF_CASE(1, 3)
    IQO_HALF_CASE(3, 3)
    IQO_HALF_CASE(1, 5)
    IQO_HALF_CASE(1, 7)
#undef IQO_HALF_CASE
    return cudaErrorInvalidValue;
}

bool halfStreamHasKernel(int NG, int NXH)
{
    return (NG == 2 || NG == 3) && (NXH == 2 || NXH == 4 || NXH == 6);
}

int halfStreamBoxRows(int NG)
{
    return 4 * NG;
}

cudaError_t launchHalfStream(const HalfArgs &a, const CUtensorMap *tmap, cudaStream_t stream)
{
#define IQO_STREAM_CASE(G, NH, ZM)                                                              \
    if (a.NG == G && a.NXH == NH && a.zmask == ZM)                                              \
        return !a.symmetric ? launchHalfStreamT<G, NH, false, false, ZM>(a, tmap, stream)       \
               : a.skipHi0  ? launchHalfStreamT<G, NH, true, true, ZM>(a, tmap, stream)         \
                            : launchHalfStreamT<G, NH, true, false, ZM>(a, tmap, stream);
    IQO_STREAM_CASE(3, 6, 0)
    IQO_STREAM_CASE(3, 4, 0)
    IQO_STREAM_CASE(3, 2, 0)
    IQO_STREAM_CASE(3, 6, 4)
    IQO_STREAM_CASE(3, 4, 4)
    IQO_STREAM_CASE(3, 2, 4)
    IQO_STREAM_CASE(2, 6, 0)
    IQO_STREAM_CASE(2, 4, 0)
    IQO_STREAM_CASE(2, 2, 0)
#undef IQO_STREAM_CASE
    return cudaErrorInvalidValue;
}

// instantiated (RS, RD, NX) triples; the parity of the first tap follows from NX (NX / 2 odd: even)
bool ratioHasKernel(int RS, int RD, int NX, int odd)
{
    if (odd != (((NX / 2) & 1) ? 0 : 1)) return false;
    return (RS == 3 && RD == 2 && (NX == 10 || NX == 6 || NX == 12 || NX == 4)) || (RS == 1 && RD == 2 && (NX == 6 || NX == 4)) ||
           (RS == 3 && RD == 4 && NX == 6) || (RS == 2 && RD == 1 && (NX == 12 || NX == 8 || NX == 4));
}

cudaError_t launchRatio(const RatioArgs &a, cudaStream_t stream)
{
    if (a.RS == 3 && a.RD == 2 && a.NX == 10)
        return a.tailZeros ? launchRatioT<3, 2, 10, 1, false>(a, stream) : launchRatioT<3, 2, 10, 0, false>(a, stream);
    if (a.RS == 3 && a.RD == 2 && a.NX == 6) return launchRatioT<3, 2, 6, 0, false>(a, stream);
    if (a.RS == 3 && a.RD == 2 && a.NX == 12) return launchRatioT<3, 2, 12, 0, true>(a, stream);
    if (a.RS == 3 && a.RD == 2 && a.NX == 4) return launchRatioT<3, 2, 4, 0, true>(a, stream);
    if (a.RS == 1 && a.RD == 2 && a.NX == 6) return launchRatioT<1, 2, 6, 0, false>(a, stream);
    if (a.RS == 1 && a.RD == 2 && a.NX == 4) return launchRatioT<1, 2, 4, 0, true>(a, stream);
    if (a.RS == 3 && a.RD == 4 && a.NX == 6) return launchRatioT<3, 4, 6, 0, false>(a, stream);
    if (a.RS == 2 && a.RD == 1 && a.NX == 12) return launchRatioT<2, 1, 12, 0, true>(a, stream);
    if (a.RS == 2 && a.RD == 1 && a.NX == 8) return launchRatioT<2, 1, 8, 0, true>(a, stream);
    if (a.RS == 2 && a.RD == 1 && a.NX == 4) return launchRatioT<2, 1, 4, 0, true>(a, stream);
    return cudaErrorInvalidValue;
}

bool lstreamHasKernel(int NP)
{
    return NP == 2 || NP == 3 || NP == 4 || NP == 6 || NP == 8 || NP == 12;
}

cudaError_t launchLStream(const LStreamArgs &a, cudaStream_t stream)
{
    switch (a.NP) {
    case 2: return launchLStreamT<2>(a, stream);
    case 3: return launchLStreamT<3>(a, stream);
    case 4: return launchLStreamT<4>(a, stream);
    case 6: return launchLStreamT<6>(a, stream);
    case 8: return launchLStreamT<8>(a, stream);
    case 12: return launchLStreamT<12>(a, stream);
    }
    return cudaErrorInvalidValue;
}

int packedPadNP(int np)
{
    const int sizes[] = {2, 3, 4, 6, 8, 12};
    for (int i = 0; i < 6; ++i)
        if (np <= sizes[i]) return sizes[i];
    return 0;  // longer kernels: not handled by the packed kernel
}

PackedGeom choosePackedGeom(const int32_t *firstX, int N, int S, int D, int npt, int ntMax)
{
    PackedGeom g;
    g.tileW = kPackedTileW;
    g.tileH = kPackedTileH;
    int widest = 1;
    for (int t0 = 0; t0 < D; t0 += g.tileW) {
        const int t1 = std::min(D, t0 + g.tileW) - 1;
        const int lo = firstX[t0] & ~7;
        const int hi = std::min(firstX[t1] + N - 1, S - 1);
        widest = std::max(widest, ((hi - lo) >> 3) + 1);
    }
    // four 32-bit words per 8-column unit, plus slack for the zero-weight pair words the
    // horizontal pass touches beyond the last written column; multiple of 4 (16-byte rows)
    g.wordsPerRow = (4 * widest + npt + 4 + 3) & ~3;
    g.smemBytes = size_t(g.tileH) * g.wordsPerRow * 4 + size_t(g.tileH) * ntMax * 4;
    return g;
}

template <bool SIGNED, int NPT>
cudaError_t launchPackedT(const PackedArgs &a, dim3 grid, size_t smem, cudaStream_t stream)
{
    static PerDeviceOnce attrSet;
    const int dev = currentDevice();
    if (!attrSet.done(dev)) {
        cudaError_t e = cudaFuncSetAttribute(resizePackedKernel<SIGNED, NPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        if (e != cudaSuccess) return e;
        attrSet.set(dev);
    }
    resizePackedKernel<SIGNED, NPT><<<grid, 256, smem, stream>>>(a);
    return cudaGetLastError();
}

cudaError_t launchPacked(const PackedArgs &a, cudaStream_t stream)
{
    const int tilesX = (a.DW + kPackedTileW - 1) / kPackedTileW;
    const int tilesY = (a.dstRows + kPackedTileH - 1) / kPackedTileH;
    if (tilesY > 65535) return cudaErrorInvalidConfiguration;
    dim3 grid(tilesX, tilesY, a.nFrames);
    const size_t smem = size_t(a.tileH) * a.wordsPerRow * 4 + size_t(a.tileH) * a.ntMax * 4;
    g_launches.fetch_add(1);
#define IQO_PACKED_CASE(N)                                                              \
    if (a.NP == N)                                                                      \
        return a.isSigned ? launchPackedT<true, N>(a, grid, smem, stream) : launchPackedT<false, N>(a, grid, smem, stream);
    IQO_PACKED_CASE(2)
    IQO_PACKED_CASE(3)
    IQO_PACKED_CASE(4)
    IQO_PACKED_CASE(6)
    IQO_PACKED_CASE(8)
    IQO_PACKED_CASE(12)
#undef IQO_PACKED_CASE
    return cudaErrorInvalidValue;
}

cudaError_t launchArea2(const uint8_t *src, uint8_t *dst, long long srcPitch, long long dstPitch, long long srcFrameStride,
                        long long dstFrameStride, int DW, int DH, int nFrames, const int32_t cy[2], const int32_t cx[2],
                        cudaStream_t stream)
{
    Area2Args a;
    a.src = src;
    a.dst = dst;
    a.srcPitch = srcPitch;
    a.dstPitch = dstPitch;
    a.srcFrameStride = srcFrameStride;
    a.dstFrameStride = dstFrameStride;
    a.DW = DW;
    a.DH = DH;
    a.chunksPerRow = (DW + 7) / 8;
    a.rcpChunks = (uint32_t)((0x100000000ull + a.chunksPerRow - 1) / a.chunksPerRow);
    a.cyLo = uint32_t(cy[0]) | (uint32_t(cy[1]) << 8);
    a.cyHi = a.cyLo << 16;
    a.cx0 = cx[0];
    a.cx1 = cx[1];
    const long long items = (long long)a.chunksPerRow * DH;
    dim3 grid((unsigned)((items + 255) / 256), (unsigned)nFrames);
    resizeArea2Kernel<<<grid, 256, 0, stream>>>(a);
    g_launches.fetch_add(1);
    return cudaGetLastError();
}

bool areaDownHasKernel(int RS, int RD, int NX)
{
    // NX: taps up to the last one that is non-zero in some phase
    return (RS == 3 && RD == 2 && NX <= 2) || (RS == 4 && RD == 3 && NX <= 2) || (RS == 2 && RD == 1 && NX <= 2) ||
           (RS == 5 && RD == 2 && NX <= 3) || (RS == 3 && RD == 1 && NX <= 3) || (RS == 4 && RD == 1 && NX <= 4);
}

int areaDownItemColumns(int RS, int RD)
{
    // source columns per item: 4 WS with WS RD == WD RS
    return RD == 1 ? 4 * RS : RD == 2 ? 4 * RS : 4 * RS;
}

cudaError_t launchAreaDown(int RS, int RD, int NX, int NXeff, const uint8_t *src, uint8_t *dst, long long srcPitch, long long dstPitch,
                           long long srcFrameStride, long long dstFrameStride, int SW, int SH, int DW, int DH, int nFrames, int NY,
                           int NYeff, const int32_t *firstY, const int32_t *rowY, const int32_t *coefY, const int32_t *cx /* [RD][NX] */,
                           cudaStream_t stream)
{
    if (!areaDownHasKernel(RS, RD, NXeff)) return cudaErrorInvalidValue;
    AreaDownArgs a;
    a.src = src;
    a.dst = dst;
    a.srcPitch = srcPitch;
    a.dstPitch = dstPitch;
    a.srcFrameStride = srcFrameStride;
    a.dstFrameStride = dstFrameStride;
    a.SW = SW;
    a.SH = SH;
    a.DW = DW;
    a.DH = DH;
    const int PS = areaDownItemColumns(RS, RD);
    if (SW % PS != 0) return cudaErrorInvalidValue;
    a.itemsPerRow = SW / PS;
    a.rcpItems = (uint32_t)((0x100000000ull + a.itemsPerRow - 1) / a.itemsPerRow);
    a.NY = NYeff;
    a.NYstride = NY;
    a.firstY = firstY;
    a.rowY = rowY;
    a.coefY = coefY;
    for (int ph = 0; ph < kAreaDownMaxRD; ++ph)
        for (int k = 0; k < kAreaDownMaxNX; ++k) a.cx2[ph][k] = (ph < RD && k < NX) ? 2 * cx[ph * NX + k] : 0;
    const long long items = (long long)a.itemsPerRow * ((DH + kAreaDownRows - 1) / kAreaDownRows);
    if (items >= (1ll << 31)) return cudaErrorInvalidConfiguration;
    dim3 grid((unsigned)((items + 255) / 256), nFrames);
    if (RS == 3 && RD == 2)
        resizeAreaDownKernel<3, 2, 3, 2, 2><<<grid, 256, 0, stream>>>(a);
    else if (RS == 4 && RD == 3)
        resizeAreaDownKernel<4, 3, 4, 3, 2><<<grid, 256, 0, stream>>>(a);
    else if (RS == 2 && RD == 1)
        resizeAreaDownKernel<2, 1, 2, 1, 2><<<grid, 256, 0, stream>>>(a);
    else if (RS == 5 && RD == 2)
        resizeAreaDownKernel<5, 2, 5, 2, 3><<<grid, 256, 0, stream>>>(a);
    else if (RS == 3 && RD == 1)
        resizeAreaDownKernel<3, 1, 3, 1, 3><<<grid, 256, 0, stream>>>(a);
    else
        resizeAreaDownKernel<4, 1, 4, 1, 4><<<grid, 256, 0, stream>>>(a);
    g_launches.fetch_add(1);
    return cudaGetLastError();
}

cudaError_t launchLinearUp(int RS, int RD, const uint8_t *src, uint8_t *dst, long long srcPitch, long long dstPitch,
                           long long srcFrameStride, long long dstFrameStride, int SW, int SH, int DW, int DH, int nFrames,
                           const int32_t *firstY, const int32_t *rowY, const int32_t *coefY, const int q1X[8],
                           cudaStream_t stream)
{
    LinearUpArgs a;
    a.src = src;
    a.dst = dst;
    a.srcPitch = srcPitch;
    a.dstPitch = dstPitch;
    a.srcFrameStride = srcFrameStride;
    a.dstFrameStride = dstFrameStride;
    a.SW = SW;
    a.SH = SH;
    a.DW = DW;
    a.DH = DH;
    if (RS <= 0 || SW % (4 * RS) != 0) return cudaErrorInvalidValue;
    a.itemsPerRow = SW / (4 * RS);
    a.firstY = firstY;
    a.rowY = rowY;
    a.coefY = coefY;
    for (int i = 0; i < 8; ++i) a.q1X[i] = 2 * q1X[i];
    a.rcpItems = (uint32_t)((0x100000000ull + a.itemsPerRow - 1) / a.itemsPerRow);
    const long long items = (long long)a.itemsPerRow * ((DH + kLinearUpRows - 1) / kLinearUpRows);
    if (items >= (1ll << 31)) return cudaErrorInvalidConfiguration;
    dim3 grid((unsigned)((items + 255) / 256), nFrames);
    if (RS == 1 && RD == 2)
        resizeLinearUpKernel<1, 2><<<grid, 256, 0, stream>>>(a);
    else if (RS == 1 && RD == 3)
        resizeLinearUpKernel<1, 3><<<grid, 256, 0, stream>>>(a);
    else if (RS == 2 && RD == 3)
        resizeLinearUpKernel<2, 3><<<grid, 256, 0, stream>>>(a);
    else if (RS == 3 && RD == 4)
        resizeLinearUpKernel<3, 4><<<grid, 256, 0, stream>>>(a);
    else if (RS == 1 && RD == 4)
        resizeLinearUpKernel<1, 4><<<grid, 256, 0, stream>>>(a);
    else if (RS == 2 && RD == 5)
        resizeLinearUpKernel<2, 5><<<grid, 256, 0, stream>>>(a);
    else if (RS == 4 && RD == 5)
        resizeLinearUpKernel<4, 5><<<grid, 256, 0, stream>>>(a);
    else if (RS == 3 && RD == 2)   // mild reductions share the item geometry and the first-tap formula
        resizeLinearUpKernel<3, 2><<<grid, 256, 0, stream>>>(a);
    else if (RS == 4 && RD == 3)
        resizeLinearUpKernel<4, 3><<<grid, 256, 0, stream>>>(a);
    else
        return cudaErrorInvalidValue;
    g_launches.fetch_add(1);
    return cudaGetLastError();
}

cudaError_t launchSmall(const SmallArgs &a, cudaStream_t stream)
{
    const int chunks = a.DW / 8;
    dim3 grid((chunks + 127) / 128, (a.DH + kSmallRows - 1) / kSmallRows, a.nFrames);
    if (grid.y > 65535) return cudaErrorInvalidConfiguration;
    g_launches.fetch_add(1);
#define IQO_SMALL_CASE(T, N, W)                                      \
    if (a.TY == T && a.NW == N && a.wbase == W) {                    \
        resizeHalfSmallKernel<T, N, W><<<grid, 128, 0, stream>>>(a); \
        return cudaGetLastError();                                   \
    }
#define IQO_SMALL_TY(T) IQO_SMALL_CASE(T, 1, 0) IQO_SMALL_CASE(T, 2, 0) IQO_SMALL_CASE(T, 2, -1) IQO_SMALL_CASE(T, 3, -1)
    IQO_SMALL_TY(1)
    IQO_SMALL_TY(2)
    IQO_SMALL_TY(3)
    IQO_SMALL_TY(4)
#undef IQO_SMALL_TY
#undef IQO_SMALL_CASE
    return cudaErrorInvalidValue;
}

int halfSourceRowsMax()
{
    return kHalfSrcMaxRows - 4;  // the last group is prefetch slack
}

unsigned long long launchCount()
{
    return g_launches.load();
}

}  // namespace iqo_b200
