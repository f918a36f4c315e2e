/*
 * srcnn_b200.h -- C ABI of the B200-native SRCNN inference hot path (libsrcnn_b200.so).
 *
 * This is the drop-in boundary for shuwang127/SRCNN_Cpp's hot path.  The reference has no FFI or
 * plugin layer; what its own host code binds for this path is (citations into /root/reference):
 *   (1) the timed body of pthreadcall(), src/srcnn.cpp:505-659  -> srcnn_process_host()/_device()
 *   (2) the conv entry points declared at src/srcnn.cpp:60-73    -> srcnn_stage_cnn_device()
 *   (3) the OpenCV stage calls at src/srcnn.cpp:509,540,577-582  -> srcnn_stage_color_bicubic_device()
 *       and at src/srcnn.cpp:637-639,657                         -> srcnn_stage_merge_device()
 *   (4) the implied library call ProcessSRCNN(), src/test.cpp:347-353 -> include/libsrcnn.h (C++)
 * INTEGRATION.md shows the binding a maintainer of the reference would add.
 *
 * Plain C: pointers, sizes, ints.  No C++/torch/OpenCV types.  Every function returns an int status
 * (0 = OK, negative = error) and never throws.  There is NO CPU fallback: without a usable CUDA
 * device srcnn_create() fails with SRCNN_E_NODEVICE.
 *
 * Pixel format everywhere: 8-bit, 3 interleaved channels (HWC), row stride in BYTES.  `order` says
 * whether memory order is B,G,R (OpenCV / bin/srcnn, src/srcnn.cpp:462) or R,G,B (ProcessSRCNN,
 * src/test.cpp:334).  Output size is ow=(int)((float)w*scale), oh=(int)((float)h*scale) -- the
 * reference's truncation (src/srcnn.cpp:573-575, src/test.cpp:357-358).
 */
#ifndef SRCNN_B200_H
#define SRCNN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SRCNN_B200_ABI_VERSION 1

#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

typedef struct srcnn_ctx srcnn_ctx;

/* Arithmetic variant of the CNN stage (Convolution99x11 + Convolution55). */
enum {
    SRCNN_VARIANT_TC = 0,   /* fused tcgen05/TMEM kernel, FP16 operands, FP32 accumulate (<=1 LSB on >=99.9% px, max 2) */
    SRCNN_VARIANT_FP32 = 1  /* strict FP32 CUDA-core kernels in the reference's summation order (bit-exact) */
};

enum { SRCNN_ORDER_BGR = 0, SRCNN_ORDER_RGB = 1 };

/* Status codes.  -1 / -2 / -3 / -10 keep the reference CLI's meanings (src/srcnn.cpp:479,493,526,555,684). */
enum {
    SRCNN_OK = 0,
    SRCNN_E_RATIO = -1,     /* scale makes an empty image ("ratio too small", src/srcnn.cpp:485-495) */
    SRCNN_E_ARG = -20,      /* null pointer, non-positive size, stride too small, unknown enum */
    SRCNN_E_NODEVICE = -30, /* no CUDA device / device is not sm_100 */
    SRCNN_E_CUDA = -31,     /* a CUDA runtime call or kernel failed; see srcnn_last_error() */
    SRCNN_E_NOMEM = -32,    /* device or pinned-host allocation failed */
    SRCNN_E_KERNEL = -33    /* a device-side guard tripped (pipeline watchdog) */
};

int srcnn_abi_version(void);
const char* srcnn_strerror(int status);

/* Context = one device + one stream + workspace + packed weights.  Re-entrant: one context per host
 * thread (or serialise calls); may be created and used from a non-main thread like the reference's
 * worker pthread (src/srcnn.cpp:717-724). */
int srcnn_create(srcnn_ctx** out, int device, int variant);
int srcnn_destroy(srcnn_ctx* ctx);
const char* srcnn_last_error(srcnn_ctx* ctx);
int srcnn_set_variant(srcnn_ctx* ctx, int variant);
int srcnn_get_variant(srcnn_ctx* ctx);
/* Use an existing CUDA stream (cudaStream_t as void*) for all work of this context; NULL = own stream. */
int srcnn_set_stream(srcnn_ctx* ctx, void* cuda_stream);
void* srcnn_get_stream(srcnn_ctx* ctx);
int srcnn_sync(srcnn_ctx* ctx);
/* Number of kernels this context has launched since creation (for bench.py's gpu_launches). */
long long srcnn_launch_count(srcnn_ctx* ctx);
int srcnn_device_sm_count(srcnn_ctx* ctx);
/* Optional per-stage device timing with CUDA events on the context stream (used by bench.py for the
 * roofline numbers).  srcnn_profile_read synchronises, returns the summed milliseconds of
 * [0] colour+bicubic, [1] fused SRCNN, [2] merge+colour-back over the whole-path calls since the
 * previous read, and the number of such calls. */
int srcnn_profile_enable(srcnn_ctx* ctx, int on);
int srcnn_profile_read(srcnn_ctx* ctx, double* ms3, int* calls);

int srcnn_out_dims(int w, int h, float scale, int* ow, int* oh);

/* Pinned host memory for callers that want true async DMA (bench e2e, CLI). */
int srcnn_host_alloc(void** p, size_t bytes);
int srcnn_host_free(void* p);

/* ---- whole path: replaces src/srcnn.cpp:505-659 -------------------------------------------------- */
/* Host buffers in, host buffers out; H2D and D2H copies are part of the call; returns when dst is ready. */
int srcnn_process_host(srcnn_ctx* ctx, const uint8_t* src, int w, int h, size_t src_stride, int order,
                       float scale, uint8_t* dst, size_t dst_stride);
/* Device buffers; enqueued on the context stream, returns without synchronising. */
int srcnn_process_device(srcnn_ctx* ctx, const uint8_t* d_src, int w, int h, size_t src_stride, int order,
                         float scale, uint8_t* d_dst, size_t dst_stride);
/* n same-sized frames, frame f at base + f*frame_stride (bytes).  Device or host variants. */
int srcnn_process_batch_device(srcnn_ctx* ctx, const uint8_t* d_src, int n, int w, int h, size_t src_stride,
                               size_t src_frame_stride, int order, float scale, uint8_t* d_dst,
                               size_t dst_stride, size_t dst_frame_stride);
int srcnn_process_batch_host(srcnn_ctx* ctx, const uint8_t* src, int n, int w, int h, size_t src_stride,
                             size_t src_frame_stride, int order, float scale, uint8_t* dst,
                             size_t dst_stride, size_t dst_frame_stride);

/* ---- row bands (gigapixel images, multi-GPU sharding) -------------------------------- */
/* Source rows [*s0,*s1) that output rows [r0,r1) of the full (w x h)*scale image depend on (6-px halo
 * in the upscaled-Y domain: 4 for conv1 + 2 for conv3, src/srcnn.cpp:273,279,203,209). */
int srcnn_band_src_rows(int h, float scale, int r0, int r1, int* s0, int* s1);
/* Computes output rows [r0,r1).  d_src points at source row s0 (as returned above) and holds rows
 * [s0,s1); d_dst points at output row r0.  All coordinates, taps and border clamps are those of the
 * full image, so the union of bands is bit-identical to the unsplit result. */
int srcnn_process_band_device(srcnn_ctx* ctx, const uint8_t* d_src, int w, int h, size_t src_stride, int s0,
                              int s1, int order, float scale, int r0, int r1, uint8_t* d_dst,
                              size_t dst_stride);

/* ---- stages (device pointers; enqueued on the context stream) ------------------------------------ */
/* cvtColor + split + 3x resize (src/srcnn.cpp:509,540,570-583): BGR8 -> three u8 planes of ow x oh. */
int srcnn_stage_color_bicubic_device(srcnn_ctx* ctx, const uint8_t* d_src, int w, int h, size_t src_stride,
                                     int order, float scale, uint8_t* d_y, uint8_t* d_cr, uint8_t* d_cb,
                                     size_t plane_pitch);
/* Convolution99x11 + Convolution55 (src/srcnn.cpp:609,627): Y plane -> Y' plane, `variant` as above. */
int srcnn_stage_cnn_device(srcnn_ctx* ctx, int variant, const uint8_t* d_y, int w, int h, size_t pitch,
                           uint8_t* d_out, size_t out_pitch);
/* FP32 variant only: also returns conv2's activations, 32 planes of h*w float (src/srcnn.cpp:602-607). */
int srcnn_stage_conv99x11_fp32_device(srcnn_ctx* ctx, const uint8_t* d_y, int w, int h, size_t pitch,
                                      float* d_act2);
/* merge + cvtColor back (src/srcnn.cpp:637-639,657): three planes -> BGR8 HWC. */
int srcnn_stage_merge_device(srcnn_ctx* ctx, const uint8_t* d_y, const uint8_t* d_cr, const uint8_t* d_cb,
                             int w, int h, size_t plane_pitch, int order, uint8_t* d_dst, size_t dst_stride);

/* ---- optional: frawscale-compatible float-plane resize --------------------------- */
/* FRAWResizeEngine::scale (src/frawscale.h:160-161, src/frawscale.cpp:162-286) on device float planes (tight
 * rows).  filter: 0 = Box, 1 = Bilinear, 2 = Bicubic (Mitchell B = C = 1/3, the reference default).  Not used by
 * the bin/srcnn path (the reference never calls frawscale either); bit-identical to the compiled reference. */
enum { SRCNN_FRAW_BOX = 0, SRCNN_FRAW_BILINEAR = 1, SRCNN_FRAW_BICUBIC = 2 };
int srcnn_fraw_scale_device(srcnn_ctx* ctx, const float* d_src, unsigned src_w, unsigned src_h, unsigned dst_w,
                            unsigned dst_h, float* d_dst, int filter);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif

#ifdef __cplusplus
}
#endif
#endif /* SRCNN_B200_H */
