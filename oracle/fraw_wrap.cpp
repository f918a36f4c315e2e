// oracle/fraw_wrap.cpp -- compiles the reference's src/frawscale.cpp UNMODIFIED (from where it lies under
// /root/reference) into oracle/_ref/libfraw.so and exports FRAWResizeEngine::scale behind a C ABI.
// Test infrastructure only: the checker for the optional frawscale-compatible resize stage.
#include <cstdint>
#include <cstring>

#include "frawscale.cpp"   // -I/root/reference/src

extern "C" __attribute__((visibility("default")))
unsigned ref_fraw_scale(const float* src, unsigned sw, unsigned sh, unsigned dw, unsigned dh, float* dst, int filter) {
    FRAWBoxFilter box;
    FRAWBilinearFilter bil;
    FRAWBicubicFilter bic;   // default Mitchell B = C = 1/3 (src/frawscale.h:93)
    FRAWGenericFilter* f = filter == 0 ? (FRAWGenericFilter*)&box : (filter == 1 ? (FRAWGenericFilter*)&bil : (FRAWGenericFilter*)&bic);
    FRAWResizeEngine eng(f);
    float* out = nullptr;
    unsigned n = eng.scale(src, sw, sh, dw, dh, &out);
    if (out) {
        memcpy(dst, out, sizeof(float) * (size_t)dw * dh);
        delete[] out;
    }
    return n;
}
