// iqo::AreaResizer on the B200 / CUDA backend: box-filter (area average) down-sampling.
//
// Drop-in for the reference's include/libiqo/AreaResizer.hpp:14-56 (same constructor and
// resize() parameter lists); the private pointer is a handle of the C ABI in include/iqo_cuda.h.
#ifndef LIBIQO_AREA_RESIZER_HPP
#define LIBIQO_AREA_RESIZER_HPP

#include "detail/backend.hpp"

namespace iqo {

class IQO_EXPORT AreaResizer {
public:
    // srcW x srcH -> dstW x dstH pixels.  Plans the weight tables and uploads them to the current
    // CUDA device; prints the reason and aborts when that is impossible (no CPU fallback).
    AreaResizer(size_t srcW, size_t srcH, size_t dstW, size_t dstH);
    ~AreaResizer();

    // One U8 plane: rows of srcSt / dstSt bytes, of which srcW / dstW are pixels.  src and dst may
    // each be host or device memory.  Only dstW bytes of a destination row are written.
    void resize(size_t srcSt, const unsigned char *src, size_t dstSt, unsigned char *dst);

private:
    AreaResizer(const AreaResizer &);                   // not copyable (declared, never defined)
    AreaResizer &operator=(const AreaResizer &);

    iqo_cuda_resizer *m_Impl;
};

}  // namespace iqo

#endif
