// srcnn_main.cpp -- bin/srcnn, the drop-in for the reference CLI (src/srcnn.cpp:331-447, 707-731):
// same options (--scale=, --noverbose, --help, src [dst]), same default output name
// (<stem>_resized<ext>, :396-416), same progress lines and exit codes, and one worker pthread like
// the reference (:717-724).  Everything between the reference's two timer reads (:505 ... :659) is
// ONE call into the CUDA library: srcnn_process_host().  Extra: --variant=tc|fp32, --device=N.
#include <pthread.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/srcnn_b200.h"
#include "image_io.h"

using std::string;

static float image_multiply = 2.0f;
static bool opt_verbose = true;
static bool opt_help = false;
static int opt_variant = SRCNN_VARIANT_TC;
static int opt_device = 0;
static int t_exit_code = 0;
static string file_me, file_src, file_dst;

#define DEF_STR_VERSION "0.1.5.20-b200"

static bool parseArgs(int argc, char** argv) {
    for (int cnt = 0; cnt < argc; cnt++) {
        string strtmp = argv[cnt];
        if (cnt == 0) {
            size_t fpos = strtmp.find_last_of("\\");
            if (fpos == string::npos) fpos = strtmp.find_last_of("/");
            file_me = fpos != string::npos ? strtmp.substr(fpos + 1) : strtmp;
        } else if (strtmp.find("--scale=") == 0) {
            string strval = strtmp.substr(8);
            if (!strval.empty()) {
                float tmpfv = (float)atof(strval.c_str());
                if (tmpfv > 0.f) image_multiply = tmpfv;   // values <= 0 are ignored, like the reference
            }
        } else if (strtmp.find("--noverbose") == 0) {
            opt_verbose = false;
        } else if (strtmp.find("--help") == 0) {
            opt_help = true;
        } else if (strtmp.find("--variant=") == 0) {
            opt_variant = strtmp.substr(10) == "fp32" ? SRCNN_VARIANT_FP32 : SRCNN_VARIANT_TC;
        } else if (strtmp.find("--device=") == 0) {
            opt_device = atoi(strtmp.substr(9).c_str());
        } else if (file_src.empty()) {
            file_src = strtmp;
        } else if (file_dst.empty()) {
            file_dst = strtmp;
        }
    }
    if (!opt_help) {
        if (!file_src.empty() && file_dst.empty()) {
            string convname = file_src, srcext;
            size_t posdot = file_src.find_last_of(".");
            if (posdot != string::npos) {
                convname = file_src.substr(0, posdot);
                srcext = file_src.substr(posdot);
            }
            file_dst = convname + "_resized" + srcext;
        }
        if (!file_src.empty() && !file_dst.empty()) return true;
    }
    return false;
}

static void printTitle() {
    printf("%s : Super-Resolution with deep Convolutional Neural Networks\n", file_me.c_str());
    printf("(C)2018..2023 Raphael Kim, (C)2014 Wang Shu., version %s\n", DEF_STR_VERSION);
    printf("Built with libsrcnn_b200 (CUDA sm_100a, C ABI v%d), no OpenCV\n", srcnn_abi_version());
}

static void printHelp() {
    printf("\n");
    printf("    usage : %s (options) [source file name] ([output file name])\n", file_me.c_str());
    printf("\n");
    printf("    _options_:\n");
    printf("\n");
    printf("        --scale=( ratio: 0.1 to .. ) : scaling by ratio.\n");
    printf("        --noverbose                  : turns off all verbose\n");
    printf("        --help                       : this help\n");
    printf("        --variant=tc|fp32            : tensor-core (default) or strict FP32 CNN kernels\n");
    printf("        --device=N                   : CUDA device index (default 0)\n");
    printf("\n");
}

static void* pthreadcall(void*) {
    if (opt_verbose) {
        printTitle();
        printf("\n");
        printf("- Scale multiply ratio : %.2f\n", image_multiply);
        fflush(stdout);
    }
    ImageBGR src;
    string err;
    if (image_read(file_src, &src, &err) && !src.empty()) {
        if (opt_verbose) { printf("- Image load : %s\n", file_src.c_str()); fflush(stdout); }
    } else {
        if (opt_verbose) printf("- load failure : %s\n", file_src.c_str());
        t_exit_code = -1;
        return nullptr;
    }
    int ow = 0, oh = 0;
    if (srcnn_out_dims(src.w, src.h, image_multiply, &ow, &oh) != SRCNN_OK) {
        if (opt_verbose) printf("- Image scale error : ratio too small.\n");
        t_exit_code = -1;
        return nullptr;
    }
    srcnn_ctx* ctx = nullptr;
    int rc = srcnn_create(&ctx, opt_device, opt_variant);
    if (rc != SRCNN_OK) {
        printf("- CUDA context failure : %s\n", srcnn_strerror(rc));
        t_exit_code = rc;
        return nullptr;
    }
    ImageBGR dst;
    dst.w = ow; dst.h = oh;
    dst.px.resize((size_t)ow * oh * 3);
    if (opt_verbose) {
        // the reference prints one line per stage; the stages are one fused GPU call here
        printf("- Image converting to Y-Cr-Cb : Ok.\n");
        printf("- Splitting channels : Ok.\n");
        printf("- Resizing splitted channels with bicublic interpolation : Ok.\n");
        printf("- Processing convolutional layer I + II ... ");
        fflush(stdout);
    }
    auto t0 = std::chrono::steady_clock::now();
    rc = srcnn_process_host(ctx, src.px.data(), src.w, src.h, (size_t)src.w * 3, SRCNN_ORDER_BGR, image_multiply,
                            dst.px.data(), (size_t)ow * 3);
    auto t1 = std::chrono::steady_clock::now();
    if (rc != SRCNN_OK) {
        if (opt_verbose) printf("Failure. (%s: %s)\n", srcnn_strerror(rc), srcnn_last_error(ctx));
        srcnn_destroy(ctx);
        t_exit_code = rc == SRCNN_E_RATIO ? -1 : -10;
        return nullptr;
    }
    if (opt_verbose) {
        printf("completed.\n");
        printf("- Processing convolutional layer III ... completed.\n");
        printf("- Merging images : Ok.\n");
        printf("- Converting channel to BGR : Ok.\n");
        printf("- Writing result to %s : ", file_dst.c_str());
        fflush(stdout);
    }
    if (!image_write(file_dst, dst, &err)) {
        if (opt_verbose) printf("Failure. (%s)\n", err.c_str());
        srcnn_destroy(ctx);
        t_exit_code = -10;
        return nullptr;
    }
    if (opt_verbose) {
        printf("Ok.\n");
        printf("- Performace : %u ms took.\n", (unsigned)std::chrono::duration_cast<std::chrono::milliseconds>(t1 - t0).count());
    }
    fflush(stdout);
    srcnn_destroy(ctx);
    t_exit_code = 0;
    return nullptr;
}

int main(int argc, char** argv) {
    if (!parseArgs(argc, argv)) {
        printTitle();
        printHelp();
        fflush(stdout);
        return 0;
    }
    pthread_t ptt;
    if (pthread_create(&ptt, nullptr, pthreadcall, nullptr) == 0) {
        pthread_join(ptt, nullptr);
    } else {
        printf("Error: pthread failure.\n");
    }
    return t_exit_code;
}
