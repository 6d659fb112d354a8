// iqo::LanczosResizer / AreaResizer / LinearResizer on top of the C ABI.
//
// Replaces the reference's front ends (src/IQOLanczosResizer.cpp:7-49, src/IQOAreaResizer.cpp:7-47,
// src/IQOLinearResizer.cpp:7-47): instead of a CPUID probe that picks one of five
// implementations there is a single CUDA backend.  The classes have no error channel
// (void functions, like the reference), so failures print the backend's message and abort.
#include <stdio.h>
#include <stdlib.h>

#include "../../include/iqo_cuda.h"
#include "../../include/libiqo/iqo.hpp"

namespace {

void die(const char *what)
{
    fprintf(stderr, "libiqo (CUDA backend): %s: %s\n", what, iqo_cuda_last_error());
    abort();
}

iqo_cuda_resizer *create(int kind, unsigned degree, size_t srcW, size_t srcH, size_t dstW, size_t dstH, size_t pxScale)
{
    iqo_cuda_resizer *r = 0;
    if (iqo_cuda_create(&r, kind, degree, srcW, srcH, dstW, dstH, pxScale) != IQO_CUDA_OK) die("constructor");
    return r;
}

}  // namespace

namespace iqo {

    LanczosResizer::LanczosResizer(unsigned int degree, size_t srcW, size_t srcH, size_t dstW, size_t dstH, size_t pxScale)
        : m_Impl(create(IQO_CUDA_LANCZOS, degree, srcW, srcH, dstW, dstH, pxScale))
    {
    }

    LanczosResizer::~LanczosResizer()
    {
        iqo_cuda_destroy(m_Impl);
    }

    void LanczosResizer::resize(size_t srcSt, const unsigned char * src, size_t dstSt, unsigned char * dst)
    {
        if (iqo_cuda_resize(m_Impl, srcSt, src, dstSt, dst) != IQO_CUDA_OK) die("LanczosResizer::resize");
    }

    AreaResizer::AreaResizer(size_t srcW, size_t srcH, size_t dstW, size_t dstH)
        : m_Impl(create(IQO_CUDA_AREA, 0, srcW, srcH, dstW, dstH, 1))
    {
    }

    AreaResizer::~AreaResizer()
    {
        iqo_cuda_destroy(m_Impl);
    }

    void AreaResizer::resize(size_t srcSt, const unsigned char * src, size_t dstSt, unsigned char * dst)
    {
        if (iqo_cuda_resize(m_Impl, srcSt, src, dstSt, dst) != IQO_CUDA_OK) die("AreaResizer::resize");
    }

    LinearResizer::LinearResizer(size_t srcW, size_t srcH, size_t dstW, size_t dstH)
        : m_Impl(create(IQO_CUDA_LINEAR, 0, srcW, srcH, dstW, dstH, 1))
    {
    }

    LinearResizer::~LinearResizer()
    {
        iqo_cuda_destroy(m_Impl);
    }

    void LinearResizer::resize(size_t srcSt, const unsigned char * src, size_t dstSt, unsigned char * dst)
    {
        if (iqo_cuda_resize(m_Impl, srcSt, src, dstSt, dst) != IQO_CUDA_OK) die("LinearResizer::resize");
    }

}
