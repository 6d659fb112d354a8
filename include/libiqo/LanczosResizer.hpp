// iqo::LanczosResizer on the B200 / CUDA backend.
//
// Drop-in for the reference's include/libiqo/LanczosResizer.hpp:14-59: constructor and resize()
// have the reference's parameter lists, so programs written against libiqo compile unchanged.
// The private pointer is a handle of the C ABI in include/iqo_cuda.h instead of a CPUID-selected
// ILanczosResizerImpl.
#ifndef LIBIQO_LANCZOS_RESIZER_HPP
#define LIBIQO_LANCZOS_RESIZER_HPP

#include "detail/backend.hpp"

namespace iqo {

class IQO_EXPORT LanczosResizer {
public:
    // degree: Lanczos window (2 = Lanczos2, 3 = Lanczos3, ...); srcW x srcH -> dstW x dstH pixels;
    // pxScale: size of one sample in luma pixels (2 for the chroma planes of YUV420).
    // Plans the coefficient tables and uploads them to the current CUDA device.  The class has no
    // error channel (neither has the reference) and there is no CPU fallback: without a device,
    // or for a size the reference leaves undefined, the reason is printed and the process aborts.
    LanczosResizer(unsigned int degree, size_t srcW, size_t srcH, size_t dstW, size_t dstH, size_t pxScale = 1);
    ~LanczosResizer();

    // One U8 plane: rows of srcSt / dstSt bytes, of which srcW / dstW are pixels.  src and dst may
    // each be host or device memory.  Only dstW bytes of a destination row are written.
    // Not re-entrant per object (like the reference); distinct objects are independent.
    void resize(size_t srcSt, const unsigned char *src, size_t dstSt, unsigned char *dst);

private:
    LanczosResizer(const LanczosResizer &);             // not copyable (declared, never defined)
    LanczosResizer &operator=(const LanczosResizer &);

    iqo_cuda_resizer *m_Impl;
};

}  // namespace iqo

#endif
