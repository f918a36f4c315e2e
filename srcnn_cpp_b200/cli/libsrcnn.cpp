// libsrcnn.cpp -- ProcessSRCNN (include/libsrcnn.h): the reference's implied library API
// (src/test.cpp:347-361) on top of the C ABI.  Host-side glue only; all arithmetic runs in the CUDA
// kernels behind srcnn_process_host().  No CPU fallback.
#include "../../include/libsrcnn.h"

#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "../../include/srcnn_b200.h"

namespace {
std::mutex g_mu;
srcnn_ctx* g_ctx = nullptr;  // one lazily created context on device 0, like the reference's process-wide state
}  // namespace

int ProcessSRCNN(const unsigned char* refbuff, unsigned w, unsigned h, unsigned d, float muliply,
                 unsigned char*& outbuff, unsigned& outbuffsz) {
    outbuff = nullptr;
    outbuffsz = 0;
    if (!refbuff || w == 0 || h == 0 || d < 1 || d > 4) return SRCNN_E_ARG;
    int ow = 0, oh = 0;
    int rc = srcnn_out_dims((int)w, (int)h, muliply, &ow, &oh);
    if (rc) return rc;
    std::lock_guard<std::mutex> lock(g_mu);
    if (!g_ctx) {
        rc = srcnn_create(&g_ctx, 0, SRCNN_VARIANT_TC);
        if (rc) return rc;
    }
    const size_t npx = (size_t)w * h, onpx = (size_t)ow * oh;
    std::vector<unsigned char> rgb(npx * 3), out_rgb(onpx * 3);
    const bool has_alpha = (d == 2 || d == 4);
    std::vector<unsigned char> alpha, alpha_out;
    if (has_alpha) { alpha.resize(npx); alpha_out.resize(onpx); }
    for (size_t i = 0; i < npx; i++) {
        const unsigned char* p = refbuff + i * d;
        if (d >= 3) { rgb[3 * i] = p[0]; rgb[3 * i + 1] = p[1]; rgb[3 * i + 2] = p[2]; }
        else { rgb[3 * i] = rgb[3 * i + 1] = rgb[3 * i + 2] = p[0]; }
        if (has_alpha) alpha[i] = p[d - 1];
    }
    rc = srcnn_process_host(g_ctx, rgb.data(), (int)w, (int)h, (size_t)w * 3, SRCNN_ORDER_RGB, muliply, out_rgb.data(), (size_t)ow * 3);
    if (rc) return rc;
    unsigned char* out = new (std::nothrow) unsigned char[onpx * d];
    if (!out) return SRCNN_E_NOMEM;
    if (has_alpha) {
        // alpha: the plain bicubic resize of the path (resize(..., CV_INTER_CUBIC), src/srcnn.cpp:577-582), no CNN
        rc = srcnn_resize_plane_host(g_ctx, alpha.data(), (int)w, (int)h, (size_t)w, muliply, alpha_out.data(), (size_t)ow);
        if (rc) { delete[] out; return rc; }
    }
    for (size_t i = 0; i < onpx; i++) {
        unsigned char* q = out + i * d;
        if (d >= 3) { q[0] = out_rgb[3 * i]; q[1] = out_rgb[3 * i + 1]; q[2] = out_rgb[3 * i + 2]; }
        else { q[0] = out_rgb[3 * i + 1]; }
        if (has_alpha) q[d - 1] = alpha_out[i];
    }
    outbuff = out;
    outbuffsz = (unsigned)(onpx * d);
    return 0;
}
