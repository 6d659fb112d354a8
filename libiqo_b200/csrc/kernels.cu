// sm_100a kernels for the libiqo resize hot path.
//
// Arithmetic contract (bit-exact with the reference's Generic implementation, SURVEY.md
// Appendix A): integer coefficient tables, 16-bit wrapped intermediate ("work") after the
// vertical pass, 32-bit horizontal sums, one rounding shift, the truncating divisions of the
// Lanczos border rows/columns.  Integer sums are order independent, so any tap order / packing
// is exact.
#include "kernels.cuh"

#include <algorithm>
#include <atomic>

namespace iqo_b200 {

namespace {

std::atomic<unsigned long long> g_launches(0);

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

}  // namespace

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
        if (g.smemBytes <= 96 * 1024 || g.tileW <= 8) break;
        g.tileW /= 2;  // extreme down-sampling ratios: narrower tile
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

unsigned long long launchCount()
{
    return g_launches.load();
}

}  // namespace iqo_b200
