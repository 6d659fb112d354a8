/*
 * iqo_cuda.h -- C ABI of the B200 (sm_100a) backend for libiqo's one-channel U8 resizers.
 *
 * This is the drop-in boundary.  libiqo's public classes keep their signatures
 * (the headers under include/libiqo/ in this repo); where the reference's constructors probe CPUID and
 * pick `LanczosResizerImpl_new<ArchAVX512|AVX2FMA|SSE4_1|NEON|Generic>()`
 * (reference src/IQOLanczosResizer.cpp:15-36, src/IQOAreaResizer.cpp:13-34,
 * src/IQOLinearResizer.cpp:13-34) there is now exactly one backend, reached through the
 * functions below.  Each entry point names the reference interface it replaces.
 *
 * Plain C: pointers, sizes and an opaque handle.  No CPU fallback exists: every function
 * returns IQO_CUDA_E_CUDA when no usable device/driver is present.
 *
 * Results are bit-identical to the reference's Generic implementation
 * (src/IQO{Lanczos,Area,Linear}ResizerImpl_Generic.cpp) for every input the reference defines.
 *
 * Threading: a handle owns its staging buffers and stream -- one in-flight call per handle
 * (the reference's impl objects are not re-entrant either: m_Work is a member,
 * src/IQOLanczosResizerImpl_Generic.cpp:279,376).  Distinct handles are independent.
 */
#ifndef IQO_CUDA_H_
#define IQO_CUDA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define IQO_CUDA_API __attribute__((visibility("default")))
#else
#define IQO_CUDA_API
#endif

typedef struct iqo_cuda_resizer iqo_cuda_resizer;

/* resampler kinds: iqo::LanczosResizer / iqo::AreaResizer / iqo::LinearResizer */
enum { IQO_CUDA_LANCZOS = 0, IQO_CUDA_AREA = 1, IQO_CUDA_LINEAR = 2 };

/* status codes (0 = success).  The reference has no error channel at all (void functions,
 * -fno-exceptions, src/CMakeLists.txt:28): the conditions below crash it or are undefined. */
enum {
    IQO_CUDA_OK = 0,
    IQO_CUDA_E_ARG = -1,         /* NULL pointer, zero size/degree/pxScale, stride < width        */
    IQO_CUDA_E_UNSUPPORTED = -2, /* Lanczos source shorter than its kernel (reference iterators
                                    desynchronise / write out of bounds, ..._Generic.cpp:390-453)  */
    IQO_CUDA_E_DEGENERATE = -3,  /* a Lanczos border denominator is 0 (reference: SIGFPE at
                                    ..._Generic.cpp:488,572) or a coefficient overflows int16      */
    IQO_CUDA_E_TOO_LARGE = -4,   /* a dimension >= 2^30                                            */
    IQO_CUDA_E_CUDA = -5,        /* CUDA runtime/driver error, or no device (no CPU fallback)      */
    IQO_CUDA_E_NOMEM = -6        /* host or device allocation failed                               */
};

/* kernel selection, for tests and benchmarks (iqo_cuda_set_path) */
enum {
    IQO_CUDA_PATH_AUTO = 0,    /* fastest eligible kernel                                          */
    IQO_CUDA_PATH_GENERIC = 1, /* the general tile kernel (any stride, ratio, kind)                */
    IQO_CUDA_PATH_NO_TMA = 2,  /* like NO_STREAM, but the tiled 2:1 Lanczos kernel reads the source
                                  with plain global loads instead of TMA (what AUTO itself does
                                  when the source pitch or base is not 8/16-byte aligned)          */
    IQO_CUDA_PATH_NO_STREAM = 3, /* like AUTO, but without the warp-streaming kernels: 2:1 Lanczos uses
                                  the tiled (TMA) kernel, other Lanczos ratios the general packed one */
    IQO_CUDA_PATH_STREAM = 4,   /* like AUTO, but the warp-streaming kernels also take launches that are
                                  too small to fill the GPU with one warp per strip (AUTO gives those
                                  to the tiled / packed kernels, which start faster)                */
    IQO_CUDA_PATH_MMA = 5,      /* Lanczos: the tensor-path kernel (both passes as integer mma.sync matrix
                                  products) takes every launch it is eligible for, small ones too     */
    IQO_CUDA_PATH_NO_MMA = 6    /* like AUTO, but never the tensor-path kernel                        */
};

/* Replaces: I{Lanczos,Area,Linear}ResizerImpl::init (reference src/IQOLanczosResizerImpl.hpp:17-22,
 * src/IQOAreaResizerImpl.hpp:17-20, src/IQOLinearResizerImpl.hpp:17-20) together with the
 * `*ResizerImpl_new<ARCH>()` factories (src/IQOLanczosResizerImpl.hpp:65-75).
 * Builds the coefficient tables on the host (calcNumCoefsForLanczos/setLanczosTable/adjustCoefs,
 * calcNumCoefsForArea/setAreaTable, setLinearTable) and uploads them to the current CUDA device.
 * degree and pxScale are ignored for AREA and LINEAR. */
IQO_CUDA_API int iqo_cuda_create(iqo_cuda_resizer **out, int kind, unsigned degree,
                                 size_t srcW, size_t srcH, size_t dstW, size_t dstH, size_t pxScale);

/* Same, on an explicit device ordinal. */
IQO_CUDA_API int iqo_cuda_create_on(iqo_cuda_resizer **out, int device, int kind, unsigned degree,
                                    size_t srcW, size_t srcH, size_t dstW, size_t dstH, size_t pxScale);

/* Replaces: the impl destructor (`delete m_Impl`, reference src/IQOLanczosResizer.cpp:39-42). */
IQO_CUDA_API void iqo_cuda_destroy(iqo_cuda_resizer *r);

/* Replaces: I*ResizerImpl::resize (reference src/IQOLanczosResizerImpl.hpp:24-28).
 * src/dst may each be a host pointer (pageable or pinned) or a device pointer; strides are in
 * bytes.  Only dstW bytes of each destination row are written.  Returns after dst is complete.
 *
 * Ordering contract for device pointers: when src or dst is device memory the call runs on the
 * legacy default stream (stream 0).  It is therefore ordered after all work previously enqueued
 * on the default stream and on any *blocking* stream (cudaStreamCreate), e.g. the kernel or
 * copy that produced src.  Work on cudaStreamNonBlocking streams is NOT ordered with it: a
 * caller that produces src on such a stream must synchronise that stream first, or use
 * iqo_cuda_resize_batch(r, 1, ..., stream), which enqueues on the caller's own stream.
 * Host-to-host calls use a private stream of the handle.
 *
 * Readable extent of device inputs: source rows are read in aligned 4/8/16-byte words inside
 * the row pitch, so srcSt * srcH bytes from src must be readable (a buffer that ends right
 * after the last row's srcW-th byte with srcSt > srcW is too short).  The same holds for the
 * batch and band entry points; host inputs are staged and have no such requirement. */
IQO_CUDA_API int iqo_cuda_resize(iqo_cuda_resizer *r, size_t srcSt, const uint8_t *src,
                                 size_t dstSt, uint8_t *dst);

/* Device-resident batch of independent frames in ONE launch sequence on `stream`
 * (a cudaStream_t passed as void*; NULL = the default stream, as everywhere in CUDA).  Frame f lives at
 * src + f*srcFrameStride / dst + f*dstFrameStride.  Asynchronous: returns after enqueueing.
 * The batched form of resize(): the reference's callers loop over frames
 * (benchmark/benchmark.cpp:1017-1033). */
IQO_CUDA_API int iqo_cuda_resize_batch(iqo_cuda_resizer *r, size_t nFrames,
                                       size_t srcSt, size_t srcFrameStride, const uint8_t *src,
                                       size_t dstSt, size_t dstFrameStride, uint8_t *dst,
                                       void *stream);

/* Host-resident batch: chunks of frames are copied host->device, resized and copied back through
 * a double-buffered pipeline on the handle's streams.  Returns after dst is complete. */
IQO_CUDA_API int iqo_cuda_resize_batch_host(iqo_cuda_resizer *r, size_t nFrames,
                                            size_t srcSt, size_t srcFrameStride, const uint8_t *src,
                                            size_t dstSt, size_t dstFrameStride, uint8_t *dst);

/* Row band of one image (gigapixel sharding): computes destination rows
 * [dstRow0, dstRow0+dstRows) from a device buffer that holds source rows
 * [srcRow0, srcRow0+srcRows) of the full image (band + halo).  `dst` addresses destination row
 * dstRow0.  Use iqo_cuda_band_src_rows() to learn which source rows a band needs. */
IQO_CUDA_API int iqo_cuda_resize_band(iqo_cuda_resizer *r, size_t dstRow0, size_t dstRows,
                                      size_t srcRow0, size_t srcRows,
                                      size_t srcSt, const uint8_t *src, size_t dstSt, uint8_t *dst,
                                      void *stream);
IQO_CUDA_API int iqo_cuda_band_src_rows(const iqo_cuda_resizer *r, size_t dstRow0, size_t dstRows,
                                        size_t *srcRow0, size_t *srcRows);

/* Whole image from host memory, row bands sharded over `nDevices` devices (devices[i] = CUDA
 * ordinal; NULL = 0..nDevices-1), halo rows copied host-side into each band's upload, one host
 * thread and stream per device, no inter-device communication.  Static: builds its own handles. */
IQO_CUDA_API int iqo_cuda_resize_bands_multi(int kind, unsigned degree,
                                             size_t srcW, size_t srcH, size_t dstW, size_t dstH, size_t pxScale,
                                             size_t srcSt, const uint8_t *src, size_t dstSt, uint8_t *dst,
                                             int nDevices, const int *devices);

/* Host-resident batch, frames sharded in contiguous blocks over `nDevices` devices. */
IQO_CUDA_API int iqo_cuda_resize_batch_multi(int kind, unsigned degree,
                                             size_t srcW, size_t srcH, size_t dstW, size_t dstH, size_t pxScale,
                                             size_t nFrames,
                                             size_t srcSt, size_t srcFrameStride, const uint8_t *src,
                                             size_t dstSt, size_t dstFrameStride, uint8_t *dst,
                                             int nDevices, const int *devices);

/* ---- planar YUV420 frames ----
 * Frame layout of the reference's sample program (sample/resize_yuv420p.cpp:66-75,117-123):
 * strides rounded up to even (stX = W + W%2, stY = H + H%2), Y plane stX*stY bytes, then U and V
 * of (stX/2)*(stY/2) bytes each.  Luma is resized W x H -> dstW x dstH with pxScale 1, both
 * chroma planes stX/2 x stY/2 -> dstStX/2 x dstStY/2 with pxScale 2 (:150-163; pxScale only
 * matters for Lanczos). */
typedef struct iqo_cuda_yuv420 iqo_cuda_yuv420;
IQO_CUDA_API int iqo_cuda_yuv420_create(iqo_cuda_yuv420 **out, int kind, unsigned degree,
                                        size_t srcW, size_t srcH, size_t dstW, size_t dstH);
IQO_CUDA_API void iqo_cuda_yuv420_destroy(iqo_cuda_yuv420 *h);
/* bytes of one source / destination frame in the layout above */
IQO_CUDA_API int iqo_cuda_yuv420_frame_bytes(const iqo_cuda_yuv420 *h, size_t *srcBytes, size_t *dstBytes);
/* nFrames consecutive frames.  Device pointers: three launches (Y, U, V planes of the whole
 * batch) enqueued on `stream`, asynchronous.  Host pointers: chunks of frames go through a
 * double-buffered H2D / kernels / D2H pipeline and the call returns when dst is complete. */
IQO_CUDA_API int iqo_cuda_yuv420_resize(iqo_cuda_yuv420 *h, size_t nFrames, const uint8_t *src, uint8_t *dst,
                                        void *stream);

/* ---- introspection (tests, benchmarks) ---- */

/* Coefficient tables as the reference would hold them in m_TablesX / m_TablesY
 * (axis 0 = X, 1 = Y): numTables rows (phases) of numCoefs int32 values. Copies up to `cap`
 * values into `out` (may be NULL). */
IQO_CUDA_API int iqo_cuda_get_table(const iqo_cuda_resizer *r, int axis, int *numCoefs, int *numTables,
                                    int32_t *out, size_t cap);
/* Host-only view of the planner (no CUDA call, works without a device): the integer
 * coefficient table of one axis plus, per destination index, the first source tap and the
 * coefficient row the kernels use.  Any output pointer may be NULL.  `coefs` receives
 * numRows*numCoefs values (phase rows first, then planner-made border rows), `first`/`row`
 * receive one value per destination index of the axis. */
IQO_CUDA_API int iqo_cuda_plan_query(int kind, unsigned degree,
                                     size_t srcW, size_t srcH, size_t dstW, size_t dstH, size_t pxScale,
                                     int axis, int *numCoefs, int *numTables, int *numRows,
                                     long long *mainBegin, long long *mainEnd,
                                     int32_t *coefs, size_t coefCap,
                                     int32_t *first, int32_t *row, size_t indexCap);
/* Host-only: which kernel family IQO_CUDA_PATH_AUTO selects for this shape when the buffers are
 * suitably aligned and the launch is large ("half_small", "half_sym", "half", "area2", "linear_up2",
 * "linear_up3", "ratio_stream", "lanczos_stream", "packed", "generic"), and why the more specialised
 * kernels are not eligible.  iqo_cuda_last_kernel() adds the variant that actually ran ("_stream",
 * "_tma"); small launches of the streaming families run on the tiled / packed kernels instead.
 * Both strings are copied NUL-terminated into the caller's buffers. */
IQO_CUDA_API int iqo_cuda_plan_kernel(int kind, unsigned degree,
                                      size_t srcW, size_t srcH, size_t dstW, size_t dstH, size_t pxScale,
                                      char *kernel, size_t kernelCap, char *why, size_t whyCap);
IQO_CUDA_API int iqo_cuda_set_path(iqo_cuda_resizer *r, int path);

/* Arithmetic of a handle.  FIXED (default) is the parity contract: the reference's Generic fixed-point path, bit-exact.
 * SIMD_FLOAT is the optional mode of SURVEY 8f-4 for callers who today get the results of the reference's
 * AVX-512 / AVX2 / SSE4.1 / NEON implementations, which compute in float and do NOT agree with Generic: float
 * tables normalised by their float sum, FMA accumulation in tap order, round-to-nearest-even, saturation
 * (reference src/IQOLanczosResizerImpl_AVX512.cpp:47-60,179-185,385-431,547-590) -- with correctly masked border
 * denominators, where the reference's resizeXborder divides by the sum of ALL coefficients (:507-519).
 * Lanczos only (IQO_CUDA_E_UNSUPPORTED otherwise); interior pixels agree with the reference's AVX-512 output to
 * within 1 LSB (tests/test_gpu_float_mode.py states and checks the tolerance). */
enum { IQO_CUDA_ARITH_FIXED = 0, IQO_CUDA_ARITH_SIMD_FLOAT = 1 };
IQO_CUDA_API int iqo_cuda_set_arithmetic(iqo_cuda_resizer *r, int arithmetic);
/* name of the kernel the last launch used, e.g. "generic" */
IQO_CUDA_API const char *iqo_cuda_last_kernel(const iqo_cuda_resizer *r);
/* number of kernels this library has launched in this process (all handles) */
IQO_CUDA_API unsigned long long iqo_cuda_launch_count(void);
/* synchronise the handle's streams */
IQO_CUDA_API int iqo_cuda_sync(iqo_cuda_resizer *r);

/* Plans (coefficient tables on the device) are cached per (device, kind, degree, sizes, pxScale)
 * and stream/staging workspaces are pooled, so that create + resize + destroy in a loop -- what
 * the reference's benchmark does (benchmark/benchmark.cpp:214-227) -- is cheap.  This drops both. */
IQO_CUDA_API void iqo_cuda_clear_cache(void);

/* pinned host memory helpers (for callers that want overlap-capable host buffers) */
IQO_CUDA_API void *iqo_cuda_host_alloc(size_t bytes);
IQO_CUDA_API void iqo_cuda_host_free(void *p);

/* message of the last failure on the calling thread ("" if none) */
IQO_CUDA_API const char *iqo_cuda_last_error(void);
IQO_CUDA_API int iqo_cuda_device_count(void);
IQO_CUDA_API const char *iqo_cuda_version(void);

#ifdef __cplusplus
}
#endif

#endif /* IQO_CUDA_H_ */
