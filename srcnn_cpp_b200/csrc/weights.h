// weights.h -- layout of the trained-parameter blob (srcnn_cpp_b200/data/srcnn_weights.bin,
// produced by tools/pack_weights.cpp from the reference's src/convdata.h; 8129 little-endian fp32).
#pragma once
namespace srcnn {
constexpr int kC1 = 64;   // CONV1_FILTERS, convdata.h:5
constexpr int kC2 = 32;   // CONV2_FILTERS, convdata.h:8
constexpr int kOffW1 = 0;                       // [64][9][9]  convdata.h:32-674
constexpr int kOffB1 = kOffW1 + kC1 * 81;       // [64]        convdata.h:19-29
constexpr int kOffW2 = kOffB1 + kC1;            // [32][64]    convdata.h:686-976
constexpr int kOffB2 = kOffW2 + kC2 * kC1;      // [32]        convdata.h:677-683
constexpr int kOffW3 = kOffB2 + kC2;            // [32][5][5]  convdata.h:982-1176
constexpr int kOffB3 = kOffW3 + kC2 * 25;       // scalar      convdata.h:979
constexpr int kNumParams = kOffB3 + 1;          // 8129
}  // namespace srcnn
extern "C" const unsigned char srcnn_weights_blob[];      // weights_blob.c (.incbin of the .bin)
extern "C" const unsigned int srcnn_weights_blob_size;
