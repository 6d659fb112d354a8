// benchmark -- command line twin of the reference's benchmark tool
// (reference benchmark/benchmark.cpp:882-1036): same flags, same report lines, same protocol:
// 256 cycles of "construct the resizers + resize Y, U, V of one planar YUV420 frame" on host
// buffers, minimum time per cycle.  It drives the libiqo C++ classes exactly like the reference's
// IQO*Resizer adapters do (benchmark/benchmark.cpp:141-226), so the constructor (served by the
// plan cache) and the host<->device copies are inside the timed region.  A second figure times
// the same frame device-resident in a batch, which is how the GPU is meant to be fed.
//
//   benchmark -m lanczos3 -iw 1920 -ih 1080 -ow 1280 -oh 720
#include <errno.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <chrono>
#include <random>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../include/iqo_cuda.h"
#include "../include/libiqo/iqo.hpp"

namespace {

// benchmark/benchmark.cpp:51-59
void fillRandom(uint8_t *p, size_t n)
{
    std::mt19937 gen(0);
    std::uniform_int_distribution<int> dist(0, 255);
    for (size_t i = 0; i < n; ++i) p[i] = uint8_t(dist(gen));
}

struct Frame {
    size_t stX, stY, sizeY, sizeU, size;
    Frame(long w, long h)
    {
        stX = size_t(w + w % 2);
        stY = size_t(h + h % 2);
        sizeY = stX * stY;
        sizeU = sizeY / 4;
        size = sizeY + 2 * sizeU;
    }
};

// one cycle through the public classes, resizers constructed inside (like the reference's adapters)
void cycle(int kind, unsigned degree, long iw, long ih, long ow, long oh, const Frame &s, const Frame &d,
           const uint8_t *src, uint8_t *dst)
{
    const uint8_t *sy = src, *su = src + s.sizeY, *sv = su + s.sizeU;
    uint8_t *dy = dst, *du = dst + d.sizeY, *dv = du + d.sizeU;
    if (kind == IQO_CUDA_LANCZOS) {
        {
            iqo::LanczosResizer r(degree, iw, ih, ow, oh);
            r.resize(s.stX, sy, d.stX, dy);
        }
        iqo::LanczosResizer r(degree, s.stX / 2, s.stY / 2, d.stX / 2, d.stY / 2, 2);
        r.resize(s.stX / 2, su, d.stX / 2, du);
        r.resize(s.stX / 2, sv, d.stX / 2, dv);
    } else if (kind == IQO_CUDA_AREA) {
        {
            iqo::AreaResizer r(iw, ih, ow, oh);
            r.resize(s.stX, sy, d.stX, dy);
        }
        iqo::AreaResizer r(s.stX / 2, s.stY / 2, d.stX / 2, d.stY / 2);
        r.resize(s.stX / 2, su, d.stX / 2, du);
        r.resize(s.stX / 2, sv, d.stX / 2, dv);
    } else {
        {
            iqo::LinearResizer r(iw, ih, ow, oh);
            r.resize(s.stX, sy, d.stX, dy);
        }
        iqo::LinearResizer r(s.stX / 2, s.stY / 2, d.stX / 2, d.stY / 2);
        r.resize(s.stX / 2, su, d.stX / 2, du);
        r.resize(s.stX / 2, sv, d.stX / 2, dv);
    }
}

}  // namespace

int main(int argc, char **argv)
{
    std::string method;
    long iw = 0, ih = 0, ow = 0, oh = 0;
    for (int i = 1; i + 1 < argc; i += 2) {
        const std::string key = argv[i];
        if (key == "-m") method = argv[i + 1];
        else if (key == "-iw") iw = atol(argv[i + 1]);
        else if (key == "-ih") ih = atol(argv[i + 1]);
        else if (key == "-ow") ow = atol(argv[i + 1]);
        else if (key == "-oh") oh = atol(argv[i + 1]);
    }
    if (iw <= 0 || ih <= 0 || ow <= 0 || oh <= 0) {
        printf("usage: benchmark -m method -iw in_width -ih in_height -ow out_width -oh out_height\n");
        printf("method: area | linear | lanczos[1-9]\n");
        return EINVAL;
    }
    int kind;
    unsigned degree = 2;
    if (method == "area") {
        kind = IQO_CUDA_AREA;
    } else if (method == "linear") {
        kind = IQO_CUDA_LINEAR;
    } else if (method.size() == 8 && method.compare(0, 7, "lanczos") == 0 && method[7] >= '1' && method[7] <= '9') {
        kind = IQO_CUDA_LANCZOS;
        degree = unsigned(method[7] - '0');
        method = "lanczos";
    } else {
        printf("invalid method: %s\n", method.c_str());
        return EINVAL;
    }
    const int numCycles = 256;
    const Frame s(iw, ih), d(ow, oh);

    printf("method: %s\n", method.c_str());
    if (kind == IQO_CUDA_LANCZOS) printf("quality\n  degree: %u\n", degree);
    if (kind == IQO_CUDA_AREA && (iw < ow || ih < oh)) printf("warning: area supports only down-sampling.\n");
    if (kind == IQO_CUDA_LINEAR && (iw > ow || ih > oh)) printf("warning: linear supports only up-sampling.\n");
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess) {
        printf("no CUDA device: this build of libiqo has no CPU path\n");
        return ENODEV;
    }
    printf("gpu\n  device: %s (%d SMs)\n", prop.name, prop.multiProcessorCount);
    printf("input\n    size: %ldx%ld\n  stride: %zux%zu\n", iw, ih, s.stX, s.stY);
    printf("output\n    size: %ldx%ld\n  stride: %zux%zu\n", ow, oh, d.stX, d.stY);
    printf("benchmark\n  cycles: %d\n", numCycles);

    // pinned host frames so that the copies run at full PCIe speed
    uint8_t *src = (uint8_t *)iqo_cuda_host_alloc(s.size), *dst = (uint8_t *)iqo_cuda_host_alloc(d.size);
    if (!src || !dst) {
        printf("host allocation failed: %s\n", iqo_cuda_last_error());
        return ENOMEM;
    }
    fillRandom(src, s.sizeY);
    fillRandom(src + s.sizeY, s.sizeU);
    fillRandom(src + s.sizeY + s.sizeU, s.sizeU);

    double best = 1e30;
    for (int i = 0; i < numCycles; ++i) {
        auto t0 = std::chrono::high_resolution_clock::now();
        cycle(kind, degree, iw, ih, ow, oh, s, d, src, dst);
        std::chrono::duration<double> dt = std::chrono::high_resolution_clock::now() - t0;
        best = std::min(best, dt.count());
    }
    printf("  elapsed time: %8.3f ms/cycle\n", best * 1000);

    // device-resident batch through the YUV420 entry point
    iqo_cuda_yuv420 *h = 0;
    if (iqo_cuda_yuv420_create(&h, kind, degree, size_t(iw), size_t(ih), size_t(ow), size_t(oh)) == IQO_CUDA_OK) {
        const size_t batch = std::max<size_t>(1, std::min<size_t>(256, (size_t(2) << 30) / (s.size + d.size)));
        uint8_t *ds = 0, *dd = 0;
        if (cudaMalloc(&ds, batch * s.size) == cudaSuccess && cudaMalloc(&dd, batch * d.size) == cudaSuccess) {
            for (size_t f = 0; f < batch; ++f) cudaMemcpy(ds + f * s.size, src, s.size, cudaMemcpyHostToDevice);
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0);
            cudaEventCreate(&e1);
            float bestMs = 1e30f;
            for (int i = 0; i < 8; ++i) {
                cudaEventRecord(e0, 0);
                iqo_cuda_yuv420_resize(h, batch, ds, dd, 0);
                cudaEventRecord(e1, 0);
                cudaEventSynchronize(e1);
                float ms = 0;
                cudaEventElapsedTime(&ms, e0, e1);
                if (i >= 2) bestMs = std::min(bestMs, ms);
            }
            printf("  device-resident batch of %zu frames: %8.4f ms/frame\n", batch, bestMs / batch);
        }
        cudaFree(ds);
        cudaFree(dd);
        iqo_cuda_yuv420_destroy(h);
    }
    iqo_cuda_host_free(src);
    iqo_cuda_host_free(dst);
    return 0;
}
