// Minimal TMA (cp.async.bulk.tensor) bring-up test: load a 256 x ROWS x 1 box of bytes from a
// 3-D tensor (x, y, frame) at possibly negative coordinates and write it back out.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

struct Args {
    alignas(64) CUtensorMap tmap;
    unsigned char *out;
    int x, y, f, rows;
};

__device__ __forceinline__ unsigned smemAddr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void __launch_bounds__(256) k(const __grid_constant__ Args p)
{
    extern __shared__ __align__(128) unsigned char tile[];
    __shared__ __align__(8) unsigned long long mbar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smemAddr(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned bytes = p.rows * 256;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(&mbar)), "r"(bytes) : "memory");
        if (MODE == 0)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(smemAddr(tile)), "l"(reinterpret_cast<unsigned long long>(&p.tmap)), "r"(p.x), "r"(p.y), "r"(p.f), "r"(smemAddr(&mbar)) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(smemAddr(tile)), "l"(reinterpret_cast<unsigned long long>(&p.tmap)), "r"(p.x), "r"(p.y), "r"(p.f), "r"(smemAddr(&mbar)) : "memory");
    }
    unsigned done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(done) : "r"(smemAddr(&mbar)) : "memory");
    }
    for (int i = threadIdx.x; i < p.rows * 256; i += blockDim.x) p.out[i] = tile[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int run(int mode, int W, int H, int F, int rows, int x, int y, int f)
{
    void *fnp = 0;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)fnp;
    std::vector<unsigned char> h((size_t)W * H * F);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (unsigned char)(i * 7 + i / W);
    unsigned char *d, *o;
    cudaMalloc(&d, h.size());
    cudaMalloc(&o, rows * 256);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    Args a;
    cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)F};
    cuuint64_t strides[2] = {(cuuint64_t)W, (cuuint64_t)W * H};
    cuuint32_t box[3] = {256, (cuuint32_t)rows, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult cr = enc(&a.tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("mode %d W %d H %d F %d rows %d at (%d,%d,%d): encode=%d ", mode, W, H, F, rows, x, y, f, (int)cr);
    a.out = o;
    a.x = x; a.y = y; a.f = f; a.rows = rows;
    size_t smem = rows * 256;
    if (mode == 0) { cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); k<0><<<1, 256, smem>>>(a); }
    else { cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); k<1><<<1, 256, smem>>>(a); }
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel=%s ", cudaGetErrorString(e));
    if (e == cudaSuccess) {
        std::vector<unsigned char> r(rows * 256);
        cudaMemcpy(r.data(), o, r.size(), cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int j = 0; j < rows; ++j)
            for (int i = 0; i < 256; ++i) {
                int sx = x + i, sy = y + j;
                unsigned char want = (sx >= 0 && sx < W && sy >= 0 && sy < H) ? h[(size_t)f * W * H + (size_t)sy * W + sx] : 0;
                bad += (r[j * 256 + i] != want);
            }
        printf("mismatches=%d", bad);
    }
    printf("\n");
    return e != cudaSuccess;
}

int main(int argc, char **argv)
{
    int which = argc > 1 ? atoi(argv[1]) : 0;
    switch (which) {
    case 0: return run(0, 1920, 1080, 4, 128, 232, 116, 2);
    case 1: return run(0, 1920, 1080, 4, 128, -8, -4, 0);
    case 2: return run(0, 64, 48, 1, 32, -8, -4, 0);
    case 3: return run(1, 1920, 1080, 4, 128, 232, 116, 2);
    case 4: return run(0, 1920, 1080, 4, 136, 1672, 1000, 3);
    case 5: return run(0, 1920, 1080, 4, 16, 0, 0, 0);
    case 6: return run(0, 1920, 1080, 4, 128, 0, 0, 0);
    case 7: return run(0, 1920, 1080, 4, 128, 16, 5, 1);
    case 8: return run(0, 1920, 1080, 4, 16, 8, 0, 0);
    case 9: return run(0, 1920, 1080, 4, 128, -16, -4, 3);
    case 10: return run(0, 1920, 1080, 4, 136, 1680, 1000, 3);
    case 11: return run(0, 1920, 1080, 4, 16, 4, 0, 0);
    }
    return 0;
}
