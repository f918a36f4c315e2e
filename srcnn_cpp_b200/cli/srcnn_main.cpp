// srcnn_main.cpp -- bin/srcnn, the drop-in for the reference CLI (src/srcnn.cpp:331-447, 707-731):
// same options (--scale=, --noverbose, --help, src [dst]), same default output name
// (<stem>_resized<ext>, :396-416), same progress lines and exit codes, and one worker pthread like
// the reference (:717-724).  Everything between the reference's two timer reads (:505 ... :659) is
// ONE call into the CUDA library: srcnn_process_host(), or srcnn_mgpu_process_banded_host() when several devices are
// named.  Extra options: --variant=tc|fp32, --device=N, --devices=0,1,..|all.
#include <pthread.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <string>
#include <vector>

#include "../../include/srcnn_b200.h"
#include "image_io.h"

using std::string;

// run-time options (the reference keeps the same things in file-scope globals, src/srcnn.cpp:40-51)
struct Options {
    float scale = 2.0f;
    bool verbose = true;
    bool help = false;
    int variant = SRCNN_VARIANT_TC;
    std::vector<int> devices = {0};
    string program, src, dst;
};
static Options opt;
static int t_exit_code = 0;

#define DEF_STR_VERSION "0.1.5.20-b200"

// One row per option.  An argument selects a row when it BEGINS with the row's text (so "--helpme" is --help and
// "--scale=" needs its value glued on, as in the reference); `apply` gets whatever follows the matched text.
struct OptionRow {
    const char* text;
    void (*apply)(Options&, const string& rest);
};
static const OptionRow kOptionTable[] = {
    {"--scale=", [](Options& o, const string& v) {
         const float f = v.empty() ? 0.f : (float)atof(v.c_str());
         if (f > 0.f) o.scale = f;                       // empty, zero and negative ratios leave the default in place
     }},
    {"--noverbose", [](Options& o, const string&) { o.verbose = false; }},
    {"--help", [](Options& o, const string&) { o.help = true; }},
    {"--variant=", [](Options& o, const string& v) { o.variant = v == "fp32" ? SRCNN_VARIANT_FP32 : SRCNN_VARIANT_TC; }},
    {"--device=", [](Options& o, const string& v) { o.devices.assign(1, atoi(v.c_str())); }},
    {"--devices=", [](Options& o, const string& v) {      // comma-separated list, or "all"
         o.devices.clear();
         if (v == "all") return;                          // empty list = every visible device
         for (size_t p = 0; p <= v.size();) {
             const size_t q = std::min(v.find(',', p), v.size());
             if (q > p) o.devices.push_back(atoi(v.substr(p, q - p).c_str()));
             p = q + 1;
         }
     }},
};

static string base_name(const string& path) {
    const size_t cut = path.find_last_of("/\\");
    return cut == string::npos ? path : path.substr(cut + 1);
}

// "<stem>_resized<ext>" next to the source (src/srcnn.cpp:396-416); a name without a dot just gets the suffix
static string default_output_name(const string& src) {
    const size_t dot = src.find_last_of('.');
    return dot == string::npos ? src + "_resized" : src.substr(0, dot) + "_resized" + src.substr(dot);
}

// true = there is something to process; false = print the title and the help (also for --help), exit code 0
static bool parseArgs(int argc, char** argv) {
    if (argc > 0) opt.program = base_name(argv[0]);
    for (int i = 1; i < argc; i++) {
        const string arg = argv[i];
        const OptionRow* hit = nullptr;
        for (const OptionRow& row : kOptionTable)
            if (arg.compare(0, strlen(row.text), row.text) == 0) {
                if (!hit || strlen(row.text) > strlen(hit->text)) hit = &row;   // "--devices=" must not stop at "--device="
            }
        if (hit) hit->apply(opt, arg.substr(strlen(hit->text)));
        else if (opt.src.empty()) opt.src = arg;
        else if (opt.dst.empty()) opt.dst = arg;          // further positional arguments are ignored
    }
    if (opt.help || opt.src.empty()) return false;
    if (opt.dst.empty()) opt.dst = default_output_name(opt.src);
    return true;
}

static void printTitle() {
    printf("%s : Super-Resolution with deep Convolutional Neural Networks\n", opt.program.c_str());
    printf("(C)2018..2023 Raphael Kim, (C)2014 Wang Shu., version %s\n", DEF_STR_VERSION);
    printf("Built with libsrcnn_b200 (CUDA sm_100a, C ABI v%d), no OpenCV\n", srcnn_abi_version());
}

static void printHelp() {
    printf("\n");
    printf("    usage : %s (options) [source file name] ([output file name])\n", opt.program.c_str());
    printf("\n");
    printf("    _options_:\n");
    printf("\n");
    printf("        --scale=( ratio: 0.1 to .. ) : scaling by ratio.\n");
    printf("        --noverbose                  : turns off all verbose\n");
    printf("        --help                       : this help\n");
    printf("        --variant=tc|fp32            : tensor-core (default) or strict FP32 CNN kernels\n");
    printf("        --device=N                   : CUDA device index (default 0)\n");
    printf("        --devices=0,1,..|all         : split the image into row bands over several GPUs\n");
    printf("\n");
}

static void* pthreadcall(void*) {
    if (opt.verbose) {
        printTitle();
        printf("\n");
        printf("- Scale multiply ratio : %.2f\n", opt.scale);
        fflush(stdout);
    }
    ImageBGR src;
    string err;
    if (image_read(opt.src, &src, &err) && !src.empty()) {
        if (opt.verbose) { printf("- Image load : %s\n", opt.src.c_str()); fflush(stdout); }
    } else {
        if (opt.verbose) printf("- load failure : %s\n", opt.src.c_str());
        t_exit_code = -1;
        return nullptr;
    }
    int ow = 0, oh = 0;
    if (srcnn_out_dims(src.w, src.h, opt.scale, &ow, &oh) != SRCNN_OK) {
        if (opt.verbose) printf("- Image scale error : ratio too small.\n");
        t_exit_code = -1;
        return nullptr;
    }
    // one GPU: a context; several: the multi-GPU driver, which cuts the image into one row band per device
    srcnn_ctx* ctx = nullptr;
    srcnn_mgpu* mg = nullptr;
    const bool multi = opt.devices.size() != 1;
    int rc = multi ? srcnn_mgpu_create(&mg, opt.devices.empty() ? nullptr : opt.devices.data(), (int)opt.devices.size(), opt.variant)
                   : srcnn_create(&ctx, opt.devices[0], opt.variant);
    if (rc != SRCNN_OK) {
        printf("- CUDA context failure : %s\n", srcnn_strerror(rc));
        t_exit_code = rc;
        return nullptr;
    }
    auto release = [&] {
        if (ctx) srcnn_destroy(ctx);
        if (mg) srcnn_mgpu_destroy(mg);
    };
    ImageBGR dst;
    dst.w = ow; dst.h = oh;
    dst.px.resize((size_t)ow * oh * 3);
    if (opt.verbose) {
        // the reference prints one line per stage; the stages are one fused GPU call here
        printf("- Image converting to Y-Cr-Cb : Ok.\n");
        printf("- Splitting channels : Ok.\n");
        printf("- Resizing splitted channels with bicublic interpolation : Ok.\n");
        printf("- Processing convolutional layer I + II ... ");
        fflush(stdout);
    }
    auto t0 = std::chrono::steady_clock::now();
    rc = multi ? srcnn_mgpu_process_banded_host(mg, src.px.data(), src.w, src.h, (size_t)src.w * 3, SRCNN_ORDER_BGR, opt.scale,
                                                dst.px.data(), (size_t)ow * 3)
               : srcnn_process_host(ctx, src.px.data(), src.w, src.h, (size_t)src.w * 3, SRCNN_ORDER_BGR, opt.scale,
                                    dst.px.data(), (size_t)ow * 3);
    auto t1 = std::chrono::steady_clock::now();
    if (rc != SRCNN_OK) {
        if (opt.verbose) printf("Failure. (%s: %s)\n", srcnn_strerror(rc), multi ? srcnn_mgpu_last_error(mg) : srcnn_last_error(ctx));
        // the reference's exit codes: -1 ratio (src/srcnn.cpp:493), -2 colour conversion (:526), -3 split (:555), -10 the rest (:684)
        int stage = SRCNN_STAGE_NONE;
        if (ctx) stage = srcnn_last_failed_stage(ctx);
        for (int i = 0; mg && i < srcnn_mgpu_device_count(mg) && stage == SRCNN_STAGE_NONE; i++)
            stage = srcnn_last_failed_stage(srcnn_mgpu_context(mg, i));
        t_exit_code = rc == SRCNN_E_RATIO ? -1 : stage == SRCNN_STAGE_COLOR_BICUBIC ? -2 : stage == SRCNN_STAGE_PLANES ? -3 : -10;
        release();
        return nullptr;
    }
    if (opt.verbose) {
        printf("completed.\n");
        printf("- Processing convolutional layer III ... completed.\n");
        printf("- Merging images : Ok.\n");
        printf("- Converting channel to BGR : Ok.\n");
        printf("- Writing result to %s : ", opt.dst.c_str());
        fflush(stdout);
    }
    if (!image_write(opt.dst, dst, &err)) {
        if (opt.verbose) printf("Failure. (%s)\n", err.c_str());
        release();
        t_exit_code = -10;
        return nullptr;
    }
    if (opt.verbose) {
        printf("Ok.\n");
        printf("- Performace : %u ms took.\n", (unsigned)std::chrono::duration_cast<std::chrono::milliseconds>(t1 - t0).count());
    }
    fflush(stdout);
    release();
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
