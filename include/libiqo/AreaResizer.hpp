#pragma once

//! @file
//! @brief Area image resampler (B200 / CUDA backend)
//!
//! Same public interface as the reference's include/libiqo/AreaResizer.hpp:14-56.

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

    class IQO_EXPORT AreaResizer
    {
    public:
        //! @param srcW     Width of source image
        //! @param srcH     Height of source image
        //! @param dstW     Width of destination image
        //! @param dstH     Height of destination image
        AreaResizer(
            size_t srcW,
            size_t srcH,
            size_t dstW,
            size_t dstH
        );

        ~AreaResizer();

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
        AreaResizer(const AreaResizer &);
        AreaResizer & operator=(const AreaResizer &);

        iqo_cuda_resizer * m_Impl;
    };

}
