// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// A thin extern "C" veneer over the UNMODIFIED reference sources, which are
// compiled where they lie under /root/reference by oracle/Makefile.  Nothing of
// the reference is copied into this repository; this file only *calls* it.
//
//  * iqo_ref_generic_resize(): instantiates the reference's Generic (fixed-point)
//    implementation directly through its factory
//    (src/IQOLanczosResizerImpl.hpp:65-75, src/IQOAreaResizerImpl.hpp,
//    src/IQOLinearResizerImpl.hpp), bypassing CPUID dispatch.  This is the
//    parity target.
//  * iqo_ref_public_*(): the reference's public classes
//    (include/libiqo/*.hpp) with its own CPUID dispatch (AVX512/AVX2/SSE4.1) and
//    OpenMP -- only ever used as the *timed* CPU baseline (built only into
//    libiqo_ref_full.so).
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>

#include "IQOLanczosResizerImpl.hpp"
#include "IQOAreaResizerImpl.hpp"
#include "IQOLinearResizerImpl.hpp"

#if defined(_OPENMP)
#include <omp.h>
#endif

namespace {

// The Generic code reads (with coefficient 0) one row past the end for Area
// with non-integer ratios and for Linear at exactly 3x; give it defined zeros.
struct Guarded {
    uint8_t *base;
    uint8_t *img;
    Guarded(size_t st, size_t w, size_t h, const uint8_t *src, size_t guardRows) {
        size_t total = st * (h + 2 * guardRows) + 64;
        base = (uint8_t *)calloc(total, 1);
        img = base + st * guardRows;
        for (size_t y = 0; y < h; ++y) memcpy(img + y * st, src + y * st, w);
    }
    ~Guarded() { free(base); }
};

}

extern "C" {

// kind: 0 lanczos, 1 area, 2 linear.  Returns 0.
int iqo_ref_generic_resize(int kind, unsigned degree,
                           size_t srcW, size_t srcH, size_t dstW, size_t dstH, size_t pxScale,
                           size_t srcSt, const uint8_t *src, size_t dstSt, uint8_t *dst)
{
    Guarded g(srcSt, srcW, srcH, src, 4);
    if (kind == 0) {
        iqo::ILanczosResizerImpl *p = iqo::LanczosResizerImpl_new<iqo::ArchGeneric>();
        p->init(degree, srcW, srcH, dstW, dstH, pxScale);
        p->resize(srcSt, g.img, dstSt, dst);
        delete p;
    } else if (kind == 1) {
        iqo::IAreaResizerImpl *p = iqo::AreaResizerImpl_new<iqo::ArchGeneric>();
        p->init(srcW, srcH, dstW, dstH);
        p->resize(srcSt, g.img, dstSt, dst);
        delete p;
    } else if (kind == 2) {
        iqo::ILinearResizerImpl *p = iqo::LinearResizerImpl_new<iqo::ArchGeneric>();
        p->init(srcW, srcH, dstW, dstH);
        p->resize(srcSt, g.img, dstSt, dst);
        delete p;
    } else {
        return -1;
    }
    return 0;
}

// Coefficient-table builders of the reference (shared by all its impls):
// src/IQOLanczosResizerImpl.hpp:79,94-102, src/IQOAreaResizerImpl.hpp:73,82-88,
// src/IQOLinearResizerImpl.hpp:76-80.
size_t iqo_ref_num_coefs_lanczos(int degree, size_t srcLen, size_t dstLen, size_t pxScale)
{
    return iqo::calcNumCoefsForLanczos(degree, srcLen, dstLen, pxScale);
}
float iqo_ref_lanczos_table(int degree, size_t srcLen, size_t dstLen, ptrdiff_t dstOffset,
                            size_t pxScale, ptrdiff_t numCoefs, float *fTable)
{
    return iqo::setLanczosTable(degree, srcLen, dstLen, dstOffset, pxScale, numCoefs, fTable);
}
size_t iqo_ref_num_coefs_area(size_t srcLen, size_t dstLen)
{
    return iqo::calcNumCoefsForArea(srcLen, dstLen);
}
float iqo_ref_area_table(size_t srcLen, size_t dstLen, ptrdiff_t dstOffset, ptrdiff_t numCoefs, float *fTable)
{
    return iqo::setAreaTable(srcLen, dstLen, dstOffset, numCoefs, fTable);
}
void iqo_ref_linear_table(size_t srcLen, size_t dstLen, float *fTable)
{
    iqo::setLinearTable(srcLen, dstLen, fTable);
}

#if defined(IQO_REF_FULL)
// ---- public API with the reference's own dispatch: timed CPU baseline only ----
struct iqo_ref_public {
    int kind;
    void *obj;
};

void *iqo_ref_public_new(int kind, unsigned degree,
                         size_t srcW, size_t srcH, size_t dstW, size_t dstH, size_t pxScale)
{
    iqo_ref_public *h = new iqo_ref_public;
    h->kind = kind;
    if (kind == 0)      h->obj = new iqo::LanczosResizer(degree, srcW, srcH, dstW, dstH, pxScale);
    else if (kind == 1) h->obj = new iqo::AreaResizer(srcW, srcH, dstW, dstH);
    else                h->obj = new iqo::LinearResizer(srcW, srcH, dstW, dstH);
    return h;
}

void iqo_ref_public_resize(void *hv, size_t srcSt, const uint8_t *src, size_t dstSt, uint8_t *dst)
{
    iqo_ref_public *h = (iqo_ref_public *)hv;
    if (h->kind == 0)      ((iqo::LanczosResizer *)h->obj)->resize(srcSt, src, dstSt, dst);
    else if (h->kind == 1) ((iqo::AreaResizer *)h->obj)->resize(srcSt, src, dstSt, dst);
    else                   ((iqo::LinearResizer *)h->obj)->resize(srcSt, src, dstSt, dst);
}

// Loop over a batch of frames inside C so that Python overhead is not timed.
void iqo_ref_public_resize_batch(void *hv, size_t nFrames,
                                 size_t srcSt, size_t srcFrameBytes, const uint8_t *src,
                                 size_t dstSt, size_t dstFrameBytes, uint8_t *dst)
{
    for (size_t f = 0; f < nFrames; ++f)
        iqo_ref_public_resize(hv, srcSt, src + f * srcFrameBytes, dstSt, dst + f * dstFrameBytes);
}

void iqo_ref_public_delete(void *hv)
{
    iqo_ref_public *h = (iqo_ref_public *)hv;
    if (h->kind == 0)      delete (iqo::LanczosResizer *)h->obj;
    else if (h->kind == 1) delete (iqo::AreaResizer *)h->obj;
    else                   delete (iqo::LinearResizer *)h->obj;
    delete h;
}

// torchrun exports OMP_NUM_THREADS=1 to its workers; the baseline wants all host cores
void iqo_ref_set_threads(int n)
{
#if defined(_OPENMP)
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int iqo_ref_threads(void)
{
#if defined(_OPENMP)
    return omp_get_max_threads();
#else
    return 1;
#endif
}
#endif

}
