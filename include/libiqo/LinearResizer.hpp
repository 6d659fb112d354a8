#pragma once

//! @file
//! @brief Linear image resampler (B200 / CUDA backend)
//!
//! Same public interface as the reference's include/libiqo/LinearResizer.hpp:14-56.

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

    class IQO_EXPORT LinearResizer
    {
    public:
        //! @param srcW     Width of source image
        //! @param srcH     Height of source image
        //! @param dstW     Width of destination image
        //! @param dstH     Height of destination image
        LinearResizer(
            size_t srcW,
            size_t srcH,
            size_t dstW,
            size_t dstH
        );

        ~LinearResizer();

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
        LinearResizer(const LinearResizer &);
        LinearResizer & operator=(const LinearResizer &);

        iqo_cuda_resizer * m_Impl;
    };

}
