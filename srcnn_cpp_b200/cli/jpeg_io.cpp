// jpeg_io.cpp -- JPEG read/write for the CLI through nvJPEG (the image has no libjpeg headers; the reference
// gets JPEG from cv::imread / cv::imwrite, src/srcnn.cpp:462,670).  File I/O is outside the reference's timed
// region and outside the parity contract: decoded pixels may differ from libjpeg's by IDCT rounding.
// Quality 95 with 4:2:0 chroma subsampling = cv::imwrite's JPEG defaults.
#include <cuda_runtime.h>
#include <nvjpeg.h>

#include <cstdio>
#include <string>
#include <vector>

#include "image_io.h"

namespace {
struct Nvj {
    nvjpegHandle_t h = nullptr;
    nvjpegJpegState_t st = nullptr;
    bool ok = false;
    Nvj() { ok = nvjpegCreateSimple(&h) == NVJPEG_STATUS_SUCCESS && nvjpegJpegStateCreate(h, &st) == NVJPEG_STATUS_SUCCESS; }
    ~Nvj() {
        if (st) nvjpegJpegStateDestroy(st);
        if (h) nvjpegDestroy(h);
    }
};
}  // namespace

bool jpeg_decode(const std::vector<uint8_t>& file, ImageBGR* out, std::string* err) {
    Nvj nv;
    if (!nv.ok) { *err = "nvJPEG initialisation failed (no CUDA device?)"; return false; }
    int nc = 0, ws[NVJPEG_MAX_COMPONENT], hs[NVJPEG_MAX_COMPONENT];
    nvjpegChromaSubsampling_t ss;
    if (nvjpegGetImageInfo(nv.h, file.data(), file.size(), &nc, &ss, ws, hs) != NVJPEG_STATUS_SUCCESS) { *err = "not a decodable JPEG"; return false; }
    const int w = ws[0], h = hs[0];
    nvjpegImage_t img = {};
    unsigned char* d = nullptr;
    if (cudaMalloc(&d, (size_t)w * h * 3) != cudaSuccess) { *err = "cudaMalloc failed"; return false; }
    img.channel[0] = d;
    img.pitch[0] = (size_t)w * 3;
    bool ok = nvjpegDecode(nv.h, nv.st, file.data(), file.size(), NVJPEG_OUTPUT_BGRI, &img, 0) == NVJPEG_STATUS_SUCCESS &&
              cudaDeviceSynchronize() == cudaSuccess;
    if (ok) {
        out->w = w; out->h = h;
        out->px.resize((size_t)w * h * 3);
        ok = cudaMemcpy(out->px.data(), d, out->px.size(), cudaMemcpyDeviceToHost) == cudaSuccess;
    }
    cudaFree(d);
    if (!ok) *err = "nvJPEG decode failed";
    return ok;
}

bool jpeg_encode(const ImageBGR& src, std::vector<uint8_t>* file, std::string* err) {
    Nvj nv;
    if (!nv.ok) { *err = "nvJPEG initialisation failed (no CUDA device?)"; return false; }
    nvjpegEncoderState_t es = nullptr;
    nvjpegEncoderParams_t ep = nullptr;
    unsigned char* d = nullptr;
    bool ok = nvjpegEncoderStateCreate(nv.h, &es, 0) == NVJPEG_STATUS_SUCCESS && nvjpegEncoderParamsCreate(nv.h, &ep, 0) == NVJPEG_STATUS_SUCCESS;
    ok = ok && nvjpegEncoderParamsSetQuality(ep, 95, 0) == NVJPEG_STATUS_SUCCESS &&
         nvjpegEncoderParamsSetSamplingFactors(ep, NVJPEG_CSS_420, 0) == NVJPEG_STATUS_SUCCESS;
    ok = ok && cudaMalloc(&d, src.px.size()) == cudaSuccess &&
         cudaMemcpy(d, src.px.data(), src.px.size(), cudaMemcpyHostToDevice) == cudaSuccess;
    if (ok) {
        nvjpegImage_t img = {};
        img.channel[0] = d;
        img.pitch[0] = (size_t)src.w * 3;
        ok = nvjpegEncodeImage(nv.h, es, ep, &img, NVJPEG_INPUT_BGRI, src.w, src.h, 0) == NVJPEG_STATUS_SUCCESS;
        size_t len = 0;
        ok = ok && nvjpegEncodeRetrieveBitstream(nv.h, es, nullptr, &len, 0) == NVJPEG_STATUS_SUCCESS;
        if (ok) {
            file->resize(len);
            ok = nvjpegEncodeRetrieveBitstream(nv.h, es, file->data(), &len, 0) == NVJPEG_STATUS_SUCCESS && cudaDeviceSynchronize() == cudaSuccess;
            file->resize(len);
        }
    }
    if (d) cudaFree(d);
    if (ep) nvjpegEncoderParamsDestroy(ep);
    if (es) nvjpegEncoderStateDestroy(es);
    if (!ok) *err = "nvJPEG encode failed";
    return ok;
}
