/* Test-vector conventions of SURVEY 8c in C (the LCG that fills source images, FNV-1a-64 of a result), for
 * bench.py and the tests at sizes where libiqo_b200/vectors.py's numpy forms are too slow (the 1 GiB source of
 * BASELINE config 5).  Not part of the resize path and not part of the C ABI in include/iqo_cuda.h; built as
 * libiqo_b200/lib/libiqo_vectors.so. */
#include <stddef.h>
#include <stdint.h>

#define VEC_API __attribute__((visibility("default")))

/* bytes [offset, offset + n) of the stream  x = seed; per byte: x = x * 1664525 + 1013904223 (mod 2^32), byte = x >> 24 */
VEC_API void iqo_vec_fill_lcg(uint8_t *buf, size_t n, uint32_t seed, uint64_t offset)
{
    /* jump ahead: x_k = A^k x_0 + C (A^k - 1) / (A - 1), as the composition of 2^i-step maps */
    uint32_t a = 1664525u, c = 1013904223u, x = seed;
    for (uint64_t k = offset; k; k >>= 1) {
        if (k & 1) x = a * x + c;
        c = (a + 1u) * c;
        a = a * a;
    }
    for (size_t i = 0; i < n; ++i) {
        x = x * 1664525u + 1013904223u;
        buf[i] = (uint8_t)(x >> 24);
    }
}

/* FNV-1a 64 continued from h (start with 0xcbf29ce484222325) over h rows of w bytes */
VEC_API uint64_t iqo_vec_fnv1a64(uint64_t h, const uint8_t *p, size_t w, size_t rows, size_t stride)
{
    for (size_t y = 0; y < rows; ++y) {
        const uint8_t *r = p + y * stride;
        for (size_t x = 0; x < w; ++x) {
            h ^= r[x];
            h *= 0x100000001b3ull;
        }
    }
    return h;
}
