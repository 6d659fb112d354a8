#pragma once

//! @file
//! @brief Lanczos image resampler (B200 / CUDA backend)
//!
//! Same public interface as the reference's include/libiqo/LanczosResizer.hpp:14-59, so code
//! written against libiqo compiles unchanged.  The pimpl no longer points at a CPUID-selected
//! I*ResizerImpl but at a handle of the CUDA backend (include/iqo_cuda.h).

#include <stddef.h>

struct iqo_cuda_resizer;

#if !defined(IQO_EXPORT)
    #if defined(__GNUC__)
        #define IQO_EXPORT __attribute__((visibility("default")))
    #else
        #define IQO_EXPORT
    #endif
#endif

namespace iqo {

    class IQO_EXPORT LanczosResizer
    {
    public:
        //! @param degree   Window size of Lanczos (ex. A=2 means Lanczos2)
        //! @param srcW     Width of source image
        //! @param srcH     Height of source image
        //! @param dstW     Width of destination image
        //! @param dstH     Height of destination image
        //! @param pxScale  Scale of a pixel (ex. 2 when U plane of YUV420 image)
        //!
        //! Builds the coefficient tables and uploads them to the current CUDA device.
        //! On failure (no device, unsupported size) prints the reason and aborts: like the
        //! reference the class has no error channel, and there is no CPU fallback.
        LanczosResizer(
            unsigned int degree,
            size_t srcW,
            size_t srcH,
            size_t dstW,
            size_t dstH,
            size_t pxScale=1
        );

        ~LanczosResizer();

        //! @param srcSt  Stride of src (in byte)
        //! @param src    Source image (host or device memory)
        //! @param dstSt  Stride of dst (in byte)
        //! @param dst    Destination image (host or device memory)
        void resize(
            size_t srcSt,
            const unsigned char * src,
            size_t dstSt,
            unsigned char * dst
        );

    private:
        // no copy
        LanczosResizer(const LanczosResizer &);
        LanczosResizer & operator=(const LanczosResizer &);

        iqo_cuda_resizer * m_Impl;
    };

}
