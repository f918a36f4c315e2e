// oracle/ref_wrap.cpp -- compiles the reference's src/srcnn.cpp UNMODIFIED, from where it lies under
// /root/reference, into oracle/_ref/libref.so, and exports its two conv functions behind a C ABI.
// Test infrastructure only (the strongest checker we have for the CNN stage: it IS the reference's
// code).  No reference source is copied into this repo; the path comes from the Makefile (-I).
//
// Calls exactly what the reference's pipeline calls at src/srcnn.cpp:609 and :627:
//   Convolution99x11(pImg[0], pImgConv2, weights_conv1_data, biases_conv1, weights_conv2_data, biases_conv2)
//   Convolution55(pImgConv2, pImgConv3, weights_conv3_data, biases_conv3)
#include <cstdint>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#define main ref_main
#include "srcnn.cpp"   // -I/root/reference/src ; its "srcnn.h" pulls <opencv2/...> from oracle/shim
#undef main

extern "C" {

__attribute__((visibility("default"))) int ref_max_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// conv1+conv2: y [h*w] u8 -> act2 [32][h*w] f32
__attribute__((visibility("default"))) void ref_conv99x11(const uint8_t* y, int h, int w, float* act2) {
    cv::Mat src(h, w, CV_8U, (void*)y);
    std::vector<cv::Mat> dst(CONV2_FILTERS);
    for (int k = 0; k < CONV2_FILTERS; k++) dst[k] = cv::Mat(h, w, CV_32F, act2 + (size_t)k * h * w);
    Convolution99x11(src, dst, weights_conv1_data, biases_conv1, weights_conv2_data, biases_conv2);
}

// conv3: act2 [32][h*w] f32 -> out [h*w] u8
__attribute__((visibility("default"))) void ref_conv55(const float* act2, int h, int w, uint8_t* out) {
    std::vector<cv::Mat> src(CONV2_FILTERS);
    for (int k = 0; k < CONV2_FILTERS; k++) src[k] = cv::Mat(h, w, CV_32F, (void*)(act2 + (size_t)k * h * w));
    cv::Mat dst(h, w, CV_8U, out);
    Convolution55(src, dst, weights_conv3_data, biases_conv3);
}

// Y -> CNN -> Y'; act2 may be null (then a scratch buffer is used)
__attribute__((visibility("default"))) int ref_cnn(const uint8_t* y, int h, int w, uint8_t* out, float* act2) {
    float* tmp = act2 ? act2 : (float*)malloc(sizeof(float) * 32 * (size_t)h * w);
    if (!tmp) return -1;
    ref_conv99x11(y, h, w, tmp);
    ref_conv55(tmp, h, w, out);
    if (!act2) free(tmp);
    return 0;
}

// the reference's parameter tables, for checking the packed blob against them
__attribute__((visibility("default"))) void ref_params(float* out /* 8129 */) {
    float* p = out;
    memcpy(p, weights_conv1_data, sizeof(float) * 64 * 81); p += 64 * 81;
    memcpy(p, biases_conv1, sizeof(float) * 64); p += 64;
    memcpy(p, weights_conv2_data, sizeof(float) * 32 * 64); p += 32 * 64;
    memcpy(p, biases_conv2, sizeof(float) * 32); p += 32;
    memcpy(p, weights_conv3_data, sizeof(float) * 32 * 25); p += 32 * 25;
    *p = biases_conv3;
}

}  // extern "C"
