// Probe for the tensor-pipe experiment (VERDICT r1 next #4-iv): fragment layout of
// ldmatrix.m16n16.trans.b8, correctness of mma.sync.m16n8k32.s32.s8.u8, and the issue rate of that mma on B200.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/bin/mma_probe tools/mma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smemAddr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// 16 rows x 16 bytes at smem (row pitch `pitch`), element (r, c) = 16 r + c
__global__ void ldmatrixLayout(uint32_t *out, int pitch)
{
    __shared__ __align__(128) uint8_t m[32 * 64];
    for (int i = threadIdx.x; i < 32 * 64; i += 32) m[i] = 0xee;
    __syncwarp();
    for (int i = threadIdx.x; i < 32 * 16; i += 32) {
        int r = i / 16, c = i % 16;
        m[r * pitch + c] = (uint8_t)(16 * (r & 15) + c + (r >= 16 ? 0 : 0));
    }
    __syncwarp();
    const int lane = threadIdx.x;
    uint32_t a0, a1;
    // x1: threads 0..15 give the 16 row addresses
    uint32_t addr = smemAddr(m + (lane & 15) * pitch);
    asm volatile("ldmatrix.sync.aligned.m16n16.x1.trans.shared.b8 {%0, %1}, [%2];" : "=r"(a0), "=r"(a1) : "r"(addr));
    out[2 * lane] = a0;
    out[2 * lane + 1] = a1;
    uint32_t b0, b1, b2, b3;
    addr = smemAddr(m + lane * pitch);   // x2: threads 16..31 address the second matrix (rows 16..31 here)
    asm volatile("ldmatrix.sync.aligned.m16n16.x2.trans.shared.b8 {%0, %1, %2, %3}, [%4];" : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3) : "r"(addr));
    out[64 + 4 * lane] = b0;
    out[64 + 4 * lane + 1] = b1;
    out[64 + 4 * lane + 2] = b2;
    out[64 + 4 * lane + 3] = b3;
}

// D[16][8] = A[16][32] (s8, row major) * B[32][8] (u8, "col": k contiguous per column) + C
__global__ void mmaCheck(const int8_t *A, const uint8_t *B, int *D)
{
    const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
    uint32_t a[4], b[2];
    auto ldA = [&](int row, int k) { return *reinterpret_cast<const uint32_t *>(A + row * 32 + k); };
    a[0] = ldA(g, 4 * t);
    a[1] = ldA(g + 8, 4 * t);
    a[2] = ldA(g, 16 + 4 * t);
    a[3] = ldA(g + 8, 16 + 4 * t);
    auto ldB = [&](int k, int n) { return *reinterpret_cast<const uint32_t *>(B + n * 32 + k); };  // B stored [n][k]
    b[0] = ldB(4 * t, g);
    b[1] = ldB(16 + 4 * t, g);
    int d[4] = {1000, 1000, 1000, 1000};
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    D[g * 8 + 2 * t] = d[0];
    D[g * 8 + 2 * t + 1] = d[1];
    D[(g + 8) * 8 + 2 * t] = d[2];
    D[(g + 8) * 8 + 2 * t + 1] = d[3];
}

template <int CHAINS>
__global__ void __launch_bounds__(1024) mmaRate(int *out, long long *cyc, int iters, uint32_t seed)
{
    uint32_t a[4] = {seed, seed * 3, seed * 5, seed * 7}, b[2] = {seed ^ 0x01020304u, seed + 9};
    int d[CHAINS][4];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) d[c][0] = d[c][1] = d[c][2] = d[c][3] = c;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c)
            asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+r"(d[c][0]), "+r"(d[c][1]), "+r"(d[c][2]), "+r"(d[c][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    long long t1 = clock64();
    int s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += d[c][0] + d[c][1] + d[c][2] + d[c][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// latency: one warp per SM, one dependent chain of k = 16 products
__global__ void __launch_bounds__(32) mmaLatencyK16(int *out, long long *cyc, int iters, uint32_t seed)
{
    uint32_t a0 = seed, a1 = seed * 3, b = seed ^ 0x01020304u;
    int d[4] = {0, 0, 0, 0};
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                     : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]) : "r"(a0), "r"(a1), "r"(b));
    long long t1 = clock64();
    out[blockIdx.x * 32 + threadIdx.x] = d[0] + d[1] + d[2] + d[3];
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// the same with a dp4a stream of the same warp mixed in (does the tensor pipe run beside the integer pipe?)
__global__ void __launch_bounds__(1024) mmaPlusDp4a(int *out, long long *cyc, int iters, uint32_t seed, int ndp)
{
    uint32_t a[4] = {seed, seed * 3, seed * 5, seed * 7}, b[2] = {seed ^ 0x01020304u, seed + 9};
    int d[2][4], e[8];
#pragma unroll
    for (int c = 0; c < 2; ++c) d[c][0] = d[c][1] = d[c][2] = d[c][3] = c;
#pragma unroll
    for (int i = 0; i < 8; ++i) e[i] = i;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < 2; ++c)
            asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+r"(d[c][0]), "+r"(d[c][1]), "+r"(d[c][2]), "+r"(d[c][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
        if (ndp) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(e[i]) : "r"(a[i & 3]), "r"(b[i & 1]));
        }
    }
    long long t1 = clock64();
    int s = 0;
#pragma unroll
    for (int c = 0; c < 2; ++c) s += d[c][0] + d[c][1] + d[c][2] + d[c][3];
#pragma unroll
    for (int i = 0; i < 8; ++i) s += e[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main()
{
    uint32_t *dout;
    cudaMalloc(&dout, 4096);
    for (int pitch : {16, 48}) {
        ldmatrixLayout<<<1, 32>>>(dout, pitch);
        uint32_t h[192];
        if (cudaMemcpy(h, dout, sizeof h, cudaMemcpyDeviceToHost) != cudaSuccess) { printf("ldmatrix kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        printf("ldmatrix.m16n16.x1.trans.b8 (element = 16*row + col, row pitch %d): lane: reg0 reg1 as (row,col) per byte\n", pitch);
        for (int l = 0; l < 32; ++l) {
            printf("lane %2d:", l);
            for (int r = 0; r < 2; ++r) {
                printf("  [");
                for (int b = 0; b < 4; ++b) { unsigned v = (h[2 * l + r] >> (8 * b)) & 255; printf(" (%2u,%2u)", v >> 4, v & 15); }
                printf(" ]");
            }
            printf("\n");
        }
        if (pitch == 16) {
            printf("x2 (second matrix = rows 16..31 of the same buffer, printed mod 16):\n");
            for (int l = 0; l < 32; l += 5) {
                printf("lane %2d:", l);
                for (int r = 0; r < 4; ++r) {
                    printf("  [");
                    for (int b = 0; b < 4; ++b) { unsigned v = (h[64 + 4 * l + r] >> (8 * b)) & 255; printf(" (%2u,%2u)", v >> 4, v & 15); }
                    printf(" ]");
                }
                printf("\n");
            }
        }
    }
    // mma check
    int8_t hA[16 * 32];
    uint8_t hB[8 * 32];
    srand(1);
    for (int i = 0; i < 512; ++i) hA[i] = (int8_t)(rand() % 256 - 128);
    for (int i = 0; i < 256; ++i) hB[i] = (uint8_t)(rand() % 256);
    int8_t *dA; uint8_t *dB; int *dD;
    cudaMalloc(&dA, 512); cudaMalloc(&dB, 256); cudaMalloc(&dD, 512);
    cudaMemcpy(dA, hA, 512, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB, 256, cudaMemcpyHostToDevice);
    mmaCheck<<<1, 32>>>(dA, dB, dD);
    int hD[128];
    cudaMemcpy(hD, dD, 512, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int m = 0; m < 16; ++m)
        for (int n = 0; n < 8; ++n) {
            int ref = 1000;
            for (int k = 0; k < 32; ++k) ref += (int)hA[m * 32 + k] * (int)hB[n * 32 + k];
            if (ref != hD[m * 8 + n]) ++bad;
        }
    printf("mma.sync.m16n8k32.s32.s8.u8 vs scalar: %d mismatches of 128\n", bad);
    // rates
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount, iters = 4096;
    int *o; long long *cy;
    cudaMalloc(&o, sms * 1024 * 4); cudaMalloc(&cy, sms * 8);
    long long *hc = (long long *)malloc(sms * 8);
    auto report = [&](const char *name, double mmaPerIter, double dpPerIter, int warps) {
        cudaDeviceSynchronize();
        cudaMemcpy(hc, cy, sms * 8, cudaMemcpyDeviceToHost);
        double c = 0;
        for (int i = 0; i < sms; ++i) c += hc[i];
        c /= sms;
        printf("%-34s %2d warps/SM: %.3f mma/clk/SM = %.0f MAC/clk/SM", name, warps, mmaPerIter * iters * warps / c, mmaPerIter * iters * warps / c * 4096);
        if (dpPerIter > 0) printf("  + %.3f dp4a warp-instr/clk/SM", dpPerIter * iters * warps / c);
        printf("  (%.0f cycles)\n", c);
    };
    for (int threads : {128, 256, 512, 1024}) {
        mmaRate<4><<<sms, threads>>>(o, cy, iters, 12345u);
        report("mma x4 independent chains", 4, 0, threads / 32);
    }
    mmaRate<1><<<sms, 1024>>>(o, cy, iters, 12345u);
    report("mma x1 dependent chain", 1, 0, 32);
    mmaRate<1><<<sms, 32>>>(o, cy, iters, 12345u);
    report("k32 latency: 1 warp/SM, dependent chain (cycles per mma = iters / rate)", 1, 0, 1);
    mmaRate<2><<<sms, 32>>>(o, cy, iters, 12345u);
    report("k32: 1 warp/SM, 2 chains", 2, 0, 1);
    mmaRate<4><<<sms, 32>>>(o, cy, iters, 12345u);
    report("k32: 1 warp/SM, 4 chains", 4, 0, 1);
    mmaLatencyK16<<<sms, 32>>>(o, cy, iters, 12345u);
    report("k16 latency: 1 warp/SM, dependent chain", 1, 0, 1);
    mmaPlusDp4a<<<sms, 1024>>>(o, cy, iters, 12345u, 0);
    report("mma x2 alone", 2, 0, 32);
    mmaPlusDp4a<<<sms, 1024>>>(o, cy, iters, 12345u, 1);
    report("mma x2 + 8 dp4a per iteration", 2, 8, 32);
    printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
