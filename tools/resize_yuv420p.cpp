// resize_yuv420p -- command line twin of the reference's sample program
// (reference sample/resize_yuv420p.cpp:36-191): same flags, same planar YUV420 file layout
// (even-rounded strides; Y, U, V planes; chroma resized with pxScale 2), same printed summary,
// but running on the CUDA backend through the YUV420 batch entry points of include/iqo_cuda.h.
//
//   resize_yuv420p -m lanczos3 -i in.yuv -iw 3840 -ih 2160 -o out.yuv -ow 1920 -oh 1080 [-n frames]
//
// -n (extension): number of consecutive frames in the file (default 1).
#include <errno.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../include/iqo_cuda.h"

namespace {

struct Options {
    std::string method, in, out;
    long iw, ih, ow, oh, frames;
    Options() : iw(0), ih(0), ow(0), oh(0), frames(1) {}
};

bool parse(int argc, char **argv, Options &o)
{
    for (int i = 1; i + 1 < argc; i += 2) {
        const std::string key = argv[i];
        const char *val = argv[i + 1];
        if (key == "-m") o.method = val;
        else if (key == "-i") o.in = val;
        else if (key == "-o") o.out = val;
        else if (key == "-iw") o.iw = atol(val);
        else if (key == "-ih") o.ih = atol(val);
        else if (key == "-ow") o.ow = atol(val);
        else if (key == "-oh") o.oh = atol(val);
        else if (key == "-n") o.frames = atol(val);
        else return false;
    }
    return !o.in.empty() && !o.out.empty() && o.iw > 0 && o.ih > 0 && o.ow > 0 && o.oh > 0 && o.frames > 0;
}

}  // namespace

int main(int argc, char **argv)
{
    Options o;
    if (!parse(argc, argv, o)) {
        printf("usage: resize_yuv420 -m method -i input.yuv -iw in_width -ih in_height -o output.yuv -ow out_width -oh out_height [-n frames]\n");
        printf("method: linear, area or lanczos[1-9]\n");
        return EINVAL;
    }
    int kind;
    unsigned degree = 2;
    if (o.method == "linear") {
        kind = IQO_CUDA_LINEAR;
    } else if (o.method == "area") {
        kind = IQO_CUDA_AREA;
    } else if (o.method.size() == 8 && o.method.compare(0, 7, "lanczos") == 0 && o.method[7] >= '1' && o.method[7] <= '9') {
        kind = IQO_CUDA_LANCZOS;
        degree = unsigned(o.method[7] - '0');
    } else {
        printf("invalid method: %s\n", o.method.c_str());
        return EINVAL;
    }

    iqo_cuda_yuv420 *h = 0;
    if (iqo_cuda_yuv420_create(&h, kind, degree, size_t(o.iw), size_t(o.ih), size_t(o.ow), size_t(o.oh)) != IQO_CUDA_OK) {
        printf("cannot set up the resizer: %s\n", iqo_cuda_last_error());
        return EINVAL;
    }
    size_t srcFrame = 0, dstFrame = 0;
    iqo_cuda_yuv420_frame_bytes(h, &srcFrame, &dstFrame);

    printf("method: %s\n", kind == IQO_CUDA_LANCZOS ? "lanczos" : o.method.c_str());
    if (kind == IQO_CUDA_LANCZOS) printf("quality\n  degree: %u\n", degree);
    printf("input\n    path: %s\n    size: %ldx%ld\n  stride: %ldx%ld\n", o.in.c_str(), o.iw, o.ih, o.iw + o.iw % 2, o.ih + o.ih % 2);
    printf("output\n    path: %s\n    size: %ldx%ld\n  stride: %ldx%ld\n", o.out.c_str(), o.ow, o.oh, o.ow + o.ow % 2, o.oh + o.oh % 2);
    printf("backend: %s, frames: %ld\n", iqo_cuda_version(), o.frames);

    std::vector<uint8_t> src(srcFrame * size_t(o.frames)), dst(dstFrame * size_t(o.frames));
    FILE *fi = fopen(o.in.c_str(), "rb");
    if (!fi) {
        int e = errno;
        perror("fopen");
        printf("Could not open \"%s\".\n", o.in.c_str());
        return e;
    }
    if (fread(&src[0], 1, src.size(), fi) < src.size()) {
        int e = errno ? errno : EIO;
        printf("Could not read %zu bytes.\n", src.size());
        fclose(fi);
        return e;
    }
    fclose(fi);

    if (iqo_cuda_yuv420_resize(h, size_t(o.frames), &src[0], &dst[0], 0) != IQO_CUDA_OK) {
        printf("resize failed: %s\n", iqo_cuda_last_error());
        return EIO;
    }
    iqo_cuda_yuv420_destroy(h);

    FILE *fo = fopen(o.out.c_str(), "wb");
    if (!fo) {
        int e = errno;
        perror("fopen");
        printf("Could not open \"%s\".\n", o.out.c_str());
        return e;
    }
    if (fwrite(&dst[0], 1, dst.size(), fo) < dst.size()) {
        int e = errno ? errno : EIO;
        printf("Could not write %zu bytes.\n", dst.size());
        fclose(fo);
        return e;
    }
    fclose(fo);
    return 0;
}
