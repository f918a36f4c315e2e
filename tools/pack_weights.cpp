// Dev-time tool: dumps the reference's trained SRCNN parameters into the flat binary blob the
// product and the oracle both load (srcnn_cpp_b200/data/srcnn_weights.bin).
//
// The parameters are DATA (8 129 fp32 values, reference src/convdata.h:19-29, 32-674, 677-683,
// 686-976, 979, 982-1176). They are taken by #including the reference header where it lies, so the
// values are exactly what the reference's compiler sees; no reference source is copied.
//
//   g++ -O0 -I/root/reference/src tools/pack_weights.cpp -o /tmp/pack_weights
//   /tmp/pack_weights srcnn_cpp_b200/data/srcnn_weights.bin
//
// Blob layout (little-endian fp32, no header), fixed by srcnn_cpp_b200/csrc/weights.h:
//   w1[64][9][9] | b1[64] | w2[32][64] | b2[32] | w3[32][5][5] | b3[1]
#include <cstdio>
#include "convdata.h"

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s out.bin\n", argv[0]); return 2; }
    FILE* f = fopen(argv[1], "wb");
    if (!f) { perror("fopen"); return 1; }
    size_t n = 0;
    n += fwrite(weights_conv1_data, sizeof(float), 64 * 81, f);
    n += fwrite(biases_conv1, sizeof(float), 64, f);
    n += fwrite(weights_conv2_data, sizeof(float), 32 * 64, f);
    n += fwrite(biases_conv2, sizeof(float), 32, f);
    n += fwrite(weights_conv3_data, sizeof(float), 32 * 25, f);
    n += fwrite(&biases_conv3, sizeof(float), 1, f);
    fclose(f);
    printf("wrote %zu floats (%zu bytes)\n", n, n * sizeof(float));
    return n == 8129 ? 0 : 1;
}
