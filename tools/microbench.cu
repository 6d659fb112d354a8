// Pipe-throughput microbenchmark for B200 (sm_100a): how many warp-instructions per clock per
// SM the scalar pipes sustain for the instruction mix the resize kernels are made of.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/microbench tools/microbench.cu
// Output: one line per test: warp-instr/clk/SM (and lanes/clk/SM = x32).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define ITERS 256
#define ACC 8

#define KERNEL(name, BODY)                                                                   \
    __global__ void __launch_bounds__(1024) name(unsigned *out, long long *cyc, unsigned seed) \
    {                                                                                        \
        unsigned a[ACC];                                                                     \
        unsigned b = seed * 3 + threadIdx.x, c = seed + 7, d = seed ^ 0x01020304u;           \
        float fb = __uint_as_float(0x3f800001u + seed), fc = 1e-9f;                          \
        _Pragma("unroll") for (int i = 0; i < ACC; ++i) a[i] = threadIdx.x * 17 + i + seed;   \
        __syncthreads();                                                                     \
        long long t0 = clock64();                                                            \
        for (int it = 0; it < ITERS; ++it) {                                                 \
            _Pragma("unroll") for (int u = 0; u < 4; ++u) {                                  \
                _Pragma("unroll") for (int i = 0; i < ACC; ++i) { BODY }                     \
            }                                                                                \
        }                                                                                    \
        long long t1 = clock64();                                                            \
        unsigned s = 0;                                                                      \
        _Pragma("unroll") for (int i = 0; i < ACC; ++i) s += a[i];                            \
        out[blockIdx.x * blockDim.x + threadIdx.x] = s + (unsigned)__float_as_uint(fb) + c + d + (unsigned)__float_as_uint(fc); \
        __syncthreads();                                                                     \
        if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;                                     \
    }

KERNEL(k_ffma, asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(*(float *)&a[i]) : "f"(fb), "f"(fc));)
KERNEL(k_imad, asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));)
KERNEL(k_imad_acc, asm volatile("mad.lo.s32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b), "r"(c));)
KERNEL(k_dp4a, asm volatile("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b), "r"(c));)
KERNEL(k_dp2a, asm volatile("dp2a.lo.s32.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b), "r"(c));)
KERNEL(k_prmt, asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(d));)
KERNEL(k_prmt_imm, asm volatile("prmt.b32 %0, %0, %1, 0x5140;" : "+r"(a[i]) : "r"(b));)
KERNEL(k_iadd, asm volatile("add.s32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));)
KERNEL(k_iadd3, asm volatile("{.reg .s32 t; add.s32 t, %0, %1; add.s32 %0, t, %2;}" : "+r"(a[i]) : "r"(b), "r"(c));)
KERNEL(k_lop3, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));)
KERNEL(k_shr, asm volatile("shr.s32 %0, %0, 3;" : "+r"(a[i]));)
KERNEL(k_shf, asm volatile("shf.r.clamp.b32 %0, %0, %1, 7;" : "+r"(a[i]) : "r"(b));)
KERNEL(k_cvtpack, asm volatile("cvt.pack.sat.u8.s32.b32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));)
KERNEL(k_cvtpack16, asm volatile("cvt.pack.sat.s16.s32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));)
KERNEL(k_imnmx, asm volatile("max.s32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));)
KERNEL(k_i2f, asm volatile("{.reg .f32 t; cvt.rn.f32.s32 t, %0; mov.b32 %0, t;}" : "+r"(a[i]));)
KERNEL(k_mix_imad_iadd, asm volatile("mad.lo.s32 %0, %1, %2, %0; add.s32 %0, %0, %1;" : "+r"(a[i]) : "r"(b), "r"(c));)
KERNEL(k_mix_dp4a_prmt, asm volatile("dp4a.u32.s32 %0, %1, %2, %0; prmt.b32 %0, %0, %1, %3;" : "+r"(a[i]) : "r"(b), "r"(c), "r"(d));)
KERNEL(k_mix_dp4a_imad, asm volatile("dp4a.u32.s32 %0, %1, %2, %0; mad.lo.s32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b), "r"(c));)
KERNEL(k_mix_ffma_imad, asm volatile("fma.rn.f32 %0, %0, %2, %3; mad.lo.s32 %1, %4, %5, %1;" : "+f"(*(float *)&a[i]), "+r"(a[(i + 4) % ACC]) : "f"(fb), "f"(fc), "r"(b), "r"(c));)
KERNEL(k_mix_ffma_dp4a, asm volatile("fma.rn.f32 %0, %0, %2, %3; dp4a.u32.s32 %1, %4, %5, %1;" : "+f"(*(float *)&a[i]), "+r"(a[(i + 4) % ACC]) : "f"(fb), "f"(fc), "r"(b), "r"(c));)
KERNEL(k_mix_dp2a_iadd_prmt, asm volatile("dp2a.lo.s32.u32 %0, %1, %2, %0; dp2a.hi.s32.s32 %0, %1, %2, %0; add.s32 %0, %0, %1; prmt.b32 %0, %0, %1, %3;" : "+r"(a[i]) : "r"(b), "r"(c), "r"(d));)

// shared memory loads
template <int VEC>
__global__ void __launch_bounds__(1024) k_lds(unsigned *out, long long *cyc, unsigned seed)
{
    __shared__ __align__(16) unsigned sm[8192];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i * seed;
    __syncthreads();
    unsigned s = 0;
    unsigned base = (threadIdx.x * VEC) & 8191;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int u = 0; u < 32; ++u) {
            unsigned idx = (base + u * 1024 * VEC / 4 + it * VEC) & (8191 & ~(VEC - 1));
            if (VEC == 1) {
                unsigned v;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(&sm[idx])));
                s += v;
            } else if (VEC == 2) {
                unsigned v0, v1;
                asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v0), "=r"(v1) : "r"((unsigned)__cvta_generic_to_shared(&sm[idx])));
                s += v0 ^ v1;
            } else {
                unsigned v0, v1, v2, v3;
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"((unsigned)__cvta_generic_to_shared(&sm[idx])));
                s += v0 ^ v1 ^ v2 ^ v3;
            }
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

typedef void (*kern_t)(unsigned *, long long *, unsigned);

static void run(const char *name, kern_t k, double instrPerThread, int nsm)
{
    unsigned *out;
    long long *cyc;
    cudaMalloc(&out, sizeof(unsigned) * 1024 * nsm);
    cudaMalloc(&cyc, sizeof(long long) * nsm);
    k<<<nsm, 1024>>>(out, cyc, 1);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<<<nsm, 1024>>>(out, cyc, 2);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long *h = (long long *)malloc(sizeof(long long) * nsm);
    cudaMemcpy(h, cyc, sizeof(long long) * nsm, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < nsm; ++i) avg += double(h[i]);
    avg /= nsm;
    double warpInstr = instrPerThread * 32.0;  // 32 warps per block
    printf("%-22s %8.3f warp-instr/clk/SM  (%7.1f lanes/clk/SM)  cycles=%.0f  ms=%.4f  clk=%.0f MHz %s\n", name,
           warpInstr / avg, 32.0 * warpInstr / avg, avg, ms, avg / (ms * 1e3), err == cudaSuccess ? "" : cudaGetErrorString(err));
    free(h);
    cudaFree(out);
    cudaFree(cyc);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int nsm = p.multiProcessorCount;
    printf("device %s, %d SMs, clock %d kHz\n", p.name, nsm, p.clockRate);
    const double n1 = double(ITERS) * 4 * ACC;
    run("ffma", k_ffma, n1, nsm);
    run("imad (a*b+c chain)", k_imad, n1, nsm);
    run("imad (acc)", k_imad_acc, n1, nsm);
    run("dp4a", k_dp4a, n1, nsm);
    run("dp2a", k_dp2a, n1, nsm);
    run("prmt (reg sel)", k_prmt, n1, nsm);
    run("prmt (imm sel)", k_prmt_imm, n1, nsm);
    run("iadd", k_iadd, n1, nsm);
    run("iadd x2 (iadd3?)", k_iadd3, n1, nsm);
    run("lop3", k_lop3, n1, nsm);
    run("shr.s32", k_shr, n1, nsm);
    run("shf", k_shf, n1, nsm);
    run("cvt.pack.sat.u8", k_cvtpack, n1, nsm);
    run("cvt.pack.sat.s16", k_cvtpack16, n1, nsm);
    run("max.s32", k_imnmx, n1, nsm);
    run("i2f", k_i2f, n1, nsm);
    run("mix imad+iadd", k_mix_imad_iadd, 2 * n1, nsm);
    run("mix dp4a+prmt", k_mix_dp4a_prmt, 2 * n1, nsm);
    run("mix dp4a+imad", k_mix_dp4a_imad, 2 * n1, nsm);
    run("mix ffma+imad", k_mix_ffma_imad, 2 * n1, nsm);
    run("mix ffma+dp4a", k_mix_ffma_dp4a, 2 * n1, nsm);
    run("mix 2dp2a+iadd+prmt", k_mix_dp2a_iadd_prmt, 4 * n1, nsm);
    run("lds.32", k_lds<1>, double(ITERS) * 32, nsm);
    run("lds.64", k_lds<2>, double(ITERS) * 32, nsm);
    run("lds.128", k_lds<4>, double(ITERS) * 32, nsm);
    return 0;
}
