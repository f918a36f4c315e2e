// oracle/shim/opencv2/core/core.hpp -- a minimal stand-in for the OpenCV C++ headers, written for
// this repo so that the reference's src/srcnn.cpp compiles UNMODIFIED here (no OpenCV C++ headers
// exist in the image).  Test infrastructure only.
//
// Only what src/srcnn.cpp touches is provided.  The conv functions (the code we want from the
// reference) use cv::Mat::{rows,cols,at<T>()}; the OpenCV *algorithms* (imread, cvtColor, split,
// resize, merge, imwrite) are stubbed out -- the oracle takes those stages from python cv2 /
// oracle/srcnn_oracle.c instead and never calls the reference's pthreadcall()/main().
#pragma once
#include <cstdlib>
#include <cstring>
#include <vector>
#include <string>

#define CV_VERSION "shim"
#define CV_8U 0
#define CV_32F 5
#define CV_BGR2YCrCb 36
#define CV_YCrCb2BGR 38
#define CV_INTER_CUBIC 2

typedef unsigned char uchar;

namespace cv {

struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
};

// Non-owning or calloc-owning dense 2-D single-channel matrix, row pitch = cols * elemsize.
struct Mat {
    int rows, cols, type;
    uchar* data;
    bool owned;
    Mat() : rows(0), cols(0), type(0), data(nullptr), owned(false) {}
    Mat(int r, int c, int t, void* p) : rows(r), cols(c), type(t), data((uchar*)p), owned(false) {}
    Mat(const Mat& o) : rows(o.rows), cols(o.cols), type(o.type), data(o.data), owned(false) {}
    Mat& operator=(const Mat& o) {
        if (this != &o) { release(); rows = o.rows; cols = o.cols; type = o.type; data = o.data; owned = false; }
        return *this;
    }
    ~Mat() { release(); }
    void release() { if (owned && data) free(data); data = nullptr; owned = false; }
    Size size() const { return Size(cols, rows); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    static size_t elem(int t) { return t == CV_32F ? 4 : 1; }
    void create(Size s, int t) {
        release();
        rows = s.height; cols = s.width; type = t;
        data = (uchar*)calloc((size_t)rows * cols, elem(t));
        owned = true;
    }
    template <typename T> T& at(int r, int c) { return ((T*)data)[(size_t)r * cols + c]; }
};

// Stubs: never executed by the oracle (ref_main / pthreadcall are not called).
inline Mat imread(const char*) { return Mat(); }
inline bool imwrite(const char*, const Mat&) { return false; }
inline void cvtColor(const Mat&, Mat&, int) {}
inline void split(const Mat&, std::vector<Mat>&) {}
inline void merge(const std::vector<Mat>&, Mat&) {}
inline void resize(const Mat&, Mat&, Size, double, double, int) {}

}  // namespace cv
