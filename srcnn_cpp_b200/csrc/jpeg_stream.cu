// jpeg_stream.cu -- stream ingest for the hot path: JPEG frames in, JPEG frames out, pixels never visit host
// memory.  The reference's only I/O is cv::imread before and cv::imwrite after its timed region (src/srcnn.cpp:462,670), one
// image per process; a stream of frames (BASELINE configs[4]) needs the decode of frame i+1 and the encode of frame i-1 to
// overlap frame i's kernels.  Three stages on three streams with a ring of three device frames:
//     nvJPEG decode (host Huffman + device IDCT, BGR interleaved)  ->  srcnn_process_device  ->  nvJPEG encode (4:2:0, like cv::imwrite)
// Codecs are outside the parity contract (JPEG decoders differ by IDCT rounding); the pixels between them go through exactly
// the kernels of srcnn_process_device.
#include <nvjpeg.h>

#include <cstdlib>
#include <string>

#include "common.h"

namespace {
constexpr int kRing = 3;
}

struct srcnn_jpeg_stream {
    srcnn_ctx* ctx = nullptr;
    nvjpegHandle_t h = nullptr;
    nvjpegJpegState_t dec[kRing] = {};
    nvjpegEncoderState_t enc[kRing] = {};
    nvjpegEncoderParams_t ep = nullptr;
    cudaStream_t s_dec = nullptr, s_enc = nullptr;
    cudaEvent_t ev_dec[kRing] = {}, ev_cmp[kRing] = {}, ev_enc[kRing] = {};
    uint8_t* d_in[kRing] = {};
    uint8_t* d_out[kRing] = {};
    size_t in_cap = 0, out_cap = 0;
    std::string err;
    double last_ms = 0.0;
};

using namespace srcnn;

static int jfail(srcnn_jpeg_stream* s, int rc, const std::string& m) {
    s->err = m;
    return rc;
}

extern "C" {

int srcnn_jpeg_stream_create(srcnn_jpeg_stream** out, srcnn_ctx* ctx, int quality) {
    if (!out || !ctx) return SRCNN_E_ARG;
    *out = nullptr;
    if (quality < 1 || quality > 100) return SRCNN_E_ARG;
    DeviceScope scope(ctx->device);
    if (!scope.ok) return SRCNN_E_CUDA;
    srcnn_jpeg_stream* s = new (std::nothrow) srcnn_jpeg_stream();
    if (!s) return SRCNN_E_NOMEM;
    s->ctx = ctx;
    bool ok = nvjpegCreateSimple(&s->h) == NVJPEG_STATUS_SUCCESS;
    for (int i = 0; ok && i < kRing; i++) {
        ok = nvjpegJpegStateCreate(s->h, &s->dec[i]) == NVJPEG_STATUS_SUCCESS && nvjpegEncoderStateCreate(s->h, &s->enc[i], 0) == NVJPEG_STATUS_SUCCESS &&
             cudaEventCreateWithFlags(&s->ev_dec[i], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&s->ev_cmp[i], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&s->ev_enc[i], cudaEventDisableTiming) == cudaSuccess;
    }
    // quality / chroma subsampling like cv::imwrite's JPEG defaults when quality = 95
    ok = ok && nvjpegEncoderParamsCreate(s->h, &s->ep, 0) == NVJPEG_STATUS_SUCCESS &&
         nvjpegEncoderParamsSetQuality(s->ep, quality, 0) == NVJPEG_STATUS_SUCCESS &&
         nvjpegEncoderParamsSetSamplingFactors(s->ep, NVJPEG_CSS_420, 0) == NVJPEG_STATUS_SUCCESS;
    ok = ok && cudaStreamCreateWithFlags(&s->s_dec, cudaStreamNonBlocking) == cudaSuccess &&
         cudaStreamCreateWithFlags(&s->s_enc, cudaStreamNonBlocking) == cudaSuccess;
    if (!ok) {
        srcnn_jpeg_stream_destroy(s);
        cudaGetLastError();
        return SRCNN_E_CUDA;
    }
    *out = s;
    return SRCNN_OK;
}

int srcnn_jpeg_stream_destroy(srcnn_jpeg_stream* s) {
    if (!s) return SRCNN_E_ARG;
    DeviceScope scope(s->ctx->device);
    if (s->s_dec) cudaStreamSynchronize(s->s_dec);
    if (s->s_enc) cudaStreamSynchronize(s->s_enc);
    for (int i = 0; i < kRing; i++) {
        if (s->dec[i]) nvjpegJpegStateDestroy(s->dec[i]);
        if (s->enc[i]) nvjpegEncoderStateDestroy(s->enc[i]);
        if (s->ev_dec[i]) cudaEventDestroy(s->ev_dec[i]);
        if (s->ev_cmp[i]) cudaEventDestroy(s->ev_cmp[i]);
        if (s->ev_enc[i]) cudaEventDestroy(s->ev_enc[i]);
        if (s->d_in[i]) cudaFree(s->d_in[i]);
        if (s->d_out[i]) cudaFree(s->d_out[i]);
    }
    if (s->ep) nvjpegEncoderParamsDestroy(s->ep);
    if (s->h) nvjpegDestroy(s->h);
    if (s->s_dec) cudaStreamDestroy(s->s_dec);
    if (s->s_enc) cudaStreamDestroy(s->s_enc);
    cudaGetLastError();
    delete s;
    return SRCNN_OK;
}

const char* srcnn_jpeg_stream_last_error(srcnn_jpeg_stream* s) { return s ? s->err.c_str() : "null handle"; }
void srcnn_jpeg_free(uint8_t* p) { free(p); }

int srcnn_jpeg_stream_process(srcnn_jpeg_stream* s, const uint8_t* const* jpegs, const size_t* sizes, int n, float scale,
                              uint8_t** out, size_t* out_sizes, int* out_w, int* out_h) {
    if (!s) return SRCNN_E_ARG;
    if (n < 0 || (n > 0 && (!jpegs || !sizes || !out || !out_sizes))) return jfail(s, SRCNN_E_ARG, "null pointer");
    if (n == 0) return SRCNN_OK;
    srcnn_ctx* c = s->ctx;
    DeviceScope scope(c->device);
    if (!scope.ok) return jfail(s, SRCNN_E_CUDA, "cudaSetDevice failed");
    for (int i = 0; i < n; i++) { out[i] = nullptr; out_sizes[i] = 0; }
    // geometry: the first frame decides, every other frame must agree (a stream)
    int nc = 0, ws[NVJPEG_MAX_COMPONENT], hs[NVJPEG_MAX_COMPONENT];
    nvjpegChromaSubsampling_t ss;
    if (nvjpegGetImageInfo(s->h, jpegs[0], sizes[0], &nc, &ss, ws, hs) != NVJPEG_STATUS_SUCCESS) return jfail(s, SRCNN_E_ARG, "frame 0 is not a decodable JPEG");
    const int w = ws[0], h = hs[0];
    int ow, oh;
    int rc = srcnn_out_dims(w, h, scale, &ow, &oh);
    if (rc) return jfail(s, rc, srcnn_strerror(rc));
    if (ow > 65535 || oh > 65535) return jfail(s, SRCNN_E_ARG, "result exceeds JPEG's 65535-pixel limit");
    if (out_w) *out_w = ow;
    if (out_h) *out_h = oh;
    const size_t in_bytes = (size_t)w * h * 3, out_bytes = (size_t)ow * oh * 3;
    if (in_bytes > s->in_cap || out_bytes > s->out_cap) {
        cudaStreamSynchronize(s->s_dec);
        cudaStreamSynchronize(s->s_enc);
        cudaStreamSynchronize(c->stream);
        for (int i = 0; i < kRing; i++) {
            if (s->d_in[i]) cudaFree(s->d_in[i]);
            if (s->d_out[i]) cudaFree(s->d_out[i]);
            s->d_in[i] = s->d_out[i] = nullptr;
            if (cudaMalloc(&s->d_in[i], in_bytes) != cudaSuccess || cudaMalloc(&s->d_out[i], out_bytes) != cudaSuccess) {
                s->in_cap = s->out_cap = 0;
                cudaGetLastError();
                return jfail(s, SRCNN_E_NOMEM, "device frame ring");
            }
        }
        s->in_cap = in_bytes;
        s->out_cap = out_bytes;
    }
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0);
    cudaEventCreate(&t1);
    cudaEventRecord(t0, c->stream);
    int status = SRCNN_OK;
    // software pipeline: iteration i decodes frame i, runs the kernels of frame i-1, encodes frame i-2
    for (int i = 0; i < n + 2 && status == SRCNN_OK; i++) {
        if (i < n) {
            const int k = i % kRing;
            int wi[NVJPEG_MAX_COMPONENT], hi[NVJPEG_MAX_COMPONENT];
            if (i > 0 && (nvjpegGetImageInfo(s->h, jpegs[i], sizes[i], &nc, &ss, wi, hi) != NVJPEG_STATUS_SUCCESS || wi[0] != w || hi[0] != h)) {
                status = jfail(s, SRCNN_E_ARG, "frame " + std::to_string(i) + " is not a JPEG of the stream's size");
                break;
            }
            // slot k was last read by the kernels of frame i-3: its encode (iteration i-1) was waited for on the host
            nvjpegImage_t img = {};
            img.channel[0] = s->d_in[k];
            img.pitch[0] = (size_t)w * 3;
            if (nvjpegDecode(s->h, s->dec[k], jpegs[i], sizes[i], NVJPEG_OUTPUT_BGRI, &img, s->s_dec) != NVJPEG_STATUS_SUCCESS) {
                status = jfail(s, SRCNN_E_ARG, "nvJPEG could not decode frame " + std::to_string(i));
                break;
            }
            cudaEventRecord(s->ev_dec[k], s->s_dec);
        }
        if (i >= 1 && i - 1 < n) {
            const int j = i - 1, k = j % kRing;
            cudaStreamWaitEvent(c->stream, s->ev_dec[k], 0);
            rc = srcnn_process_device(c, s->d_in[k], w, h, (size_t)w * 3, SRCNN_ORDER_BGR, scale, s->d_out[k], (size_t)ow * 3);
            if (rc) { status = jfail(s, rc, srcnn_last_error(c)); break; }
            cudaEventRecord(s->ev_cmp[k], c->stream);
        }
        if (i >= 2) {
            const int j = i - 2, k = j % kRing;
            cudaStreamWaitEvent(s->s_enc, s->ev_cmp[k], 0);
            nvjpegImage_t img = {};
            img.channel[0] = s->d_out[k];
            img.pitch[0] = (size_t)ow * 3;
            size_t len = 0;
            bool ok = nvjpegEncodeImage(s->h, s->enc[k], s->ep, &img, NVJPEG_INPUT_BGRI, ow, oh, s->s_enc) == NVJPEG_STATUS_SUCCESS &&
                      nvjpegEncodeRetrieveBitstream(s->h, s->enc[k], nullptr, &len, s->s_enc) == NVJPEG_STATUS_SUCCESS;
            if (ok) {
                out[j] = (uint8_t*)malloc(len);
                ok = out[j] && nvjpegEncodeRetrieveBitstream(s->h, s->enc[k], out[j], &len, s->s_enc) == NVJPEG_STATUS_SUCCESS &&
                     cudaStreamSynchronize(s->s_enc) == cudaSuccess;
                out_sizes[j] = len;
            }
            if (!ok) { status = jfail(s, SRCNN_E_CUDA, "nvJPEG could not encode frame " + std::to_string(j)); break; }
        }
    }
    cudaEventRecord(t1, c->stream);
    cudaStreamSynchronize(s->s_dec);
    cudaStreamSynchronize(s->s_enc);
    cudaStreamSynchronize(c->stream);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, t0, t1);
    s->last_ms = ms;
    cudaEventDestroy(t0);
    cudaEventDestroy(t1);
    if (status != SRCNN_OK) {
        for (int i = 0; i < n; i++) { free(out[i]); out[i] = nullptr; out_sizes[i] = 0; }
        cudaGetLastError();
    }
    return status;
}

}  // extern "C"
