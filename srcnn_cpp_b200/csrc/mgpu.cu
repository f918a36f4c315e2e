// mgpu.cu -- all GPUs of one box behind one C-ABI call (include/srcnn_b200.h, srcnn_mgpu_*).
//
// The reference parallelises ONE call over the host's cores with an OpenMP row loop (src/srcnn.cpp:283,213) from one worker
// pthread (:717-724).  The drop-in equivalent on a B200 box: one call that fans frames (frame f -> worker f mod n) or
// output-row bands (6-px halo in the upscaled-Y domain: srcnn_band_src_rows) out over a device list.  The path needs no
// exchange step: inputs are scattered with overlapping halo rows, outputs are disjoint, so there is no NCCL and
// no peer traffic here -- one persistent host thread per device, each with its own srcnn_ctx (stream set, workspace,
// weights), started once and fed through a mailbox.  Results are bit-identical to the one-GPU call because every tap and
// clamp uses full-image coordinates (tests/test_mgpu.py).
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <new>
#include <string>
#include <thread>

#include "common.h"

namespace srcnn {

struct Worker {
    int device = 0;
    srcnn_ctx* ctx = nullptr;
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::function<int(Worker&)> job;   // set by the caller, cleared by the worker
    bool has_job = false, done = false, quit = false;
    int rc = 0;
    double ms = 0.0;                   // time of the worker's share of the last call
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;

    void loop() {
        cudaSetDevice(device);         // this thread never runs anything else
        for (;;) {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return has_job || quit; });
            if (quit) return;
            lk.unlock();
            const int r = job(*this);
            lk.lock();
            rc = r;
            has_job = false;
            done = true;
            cv.notify_all();
        }
    }
};

}  // namespace srcnn

struct srcnn_mgpu {
    std::vector<srcnn::Worker*> w;
    std::mutex call_mu;                // calls are serialised
    std::string err;
    double wall_ms = 0.0;
};

namespace srcnn {

static int mfail(srcnn_mgpu* m, int rc, const std::string& msg) {
    m->err = msg;
    return rc;
}

// hands job(i) to every worker and waits for all of them; returns the first failure
static int run_all(srcnn_mgpu* m, const std::function<int(Worker&, int)>& job) {
    const auto t0 = std::chrono::steady_clock::now();
    for (size_t i = 0; i < m->w.size(); i++) {
        Worker* k = m->w[i];
        std::lock_guard<std::mutex> lk(k->mu);
        k->job = [&job, i](Worker& me) { return job(me, (int)i); };
        k->done = false;
        k->has_job = true;
        k->cv.notify_all();
    }
    int rc = SRCNN_OK;
    for (size_t i = 0; i < m->w.size(); i++) {
        Worker* k = m->w[i];
        std::unique_lock<std::mutex> lk(k->mu);
        k->cv.wait(lk, [&] { return k->done; });
        if (k->rc != SRCNN_OK && rc == SRCNN_OK) {
            rc = k->rc;
            m->err = "worker " + std::to_string(i) + " (device " + std::to_string(k->device) + "): " + srcnn_last_error(k->ctx);
        }
    }
    m->wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return rc;
}

// a worker's share on device buffers: enqueue, bracket with events on its stream, wait
template <class F>
static int timed_device(Worker& me, F&& enqueue) {
    cudaStream_t st = (cudaStream_t)srcnn_get_stream(me.ctx);
    if (cudaEventRecord(me.ev0, st) != cudaSuccess) return SRCNN_E_CUDA;
    int rc = enqueue();
    if (rc) return rc;
    if (cudaEventRecord(me.ev1, st) != cudaSuccess) return SRCNN_E_CUDA;
    rc = srcnn_sync(me.ctx);
    if (rc) return rc;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, me.ev0, me.ev1) != cudaSuccess) return SRCNN_E_CUDA;
    me.ms = ms;
    return SRCNN_OK;
}
template <class F>
static int timed_host(Worker& me, F&& run) {
    const auto t0 = std::chrono::steady_clock::now();
    const int rc = run();
    me.ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return rc;
}

static int band_rows(int n, int oh, int i, int* r0, int* r1) {
    *r0 = (int)((long long)oh * i / n);
    *r1 = (int)((long long)oh * (i + 1) / n);
    return *r1 > *r0;
}

}  // namespace srcnn

using namespace srcnn;

extern "C" {

int srcnn_mgpu_create(srcnn_mgpu** out, const int* devices, int n, int variant) {
    if (!out) return SRCNN_E_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        return SRCNN_E_NODEVICE;
    }
    if (n <= 0) { n = ndev; devices = nullptr; }
    if (n > 64) return SRCNN_E_ARG;
    srcnn_mgpu* m = new (std::nothrow) srcnn_mgpu();
    if (!m) return SRCNN_E_NOMEM;
    int rc = SRCNN_OK;
    for (int i = 0; i < n && rc == SRCNN_OK; i++) {
        const int dev = devices ? devices[i] : i;
        Worker* k = new (std::nothrow) Worker();
        if (!k) { rc = SRCNN_E_NOMEM; break; }
        k->device = dev;
        m->w.push_back(k);
        rc = srcnn_create(&k->ctx, dev, variant);   // also validates the device index and its architecture
        if (rc) break;
        DeviceScope scope(dev);
        if (cudaEventCreate(&k->ev0) != cudaSuccess || cudaEventCreate(&k->ev1) != cudaSuccess) { rc = SRCNN_E_CUDA; break; }
    }
    if (rc == SRCNN_OK)
        for (Worker* k : m->w) k->th = std::thread([k] { k->loop(); });
    if (rc) {
        srcnn_mgpu_destroy(m);
        return rc;
    }
    *out = m;
    return SRCNN_OK;
}

int srcnn_mgpu_destroy(srcnn_mgpu* m) {
    if (!m) return SRCNN_E_ARG;
    for (Worker* k : m->w) {
        if (k->th.joinable()) {
            {
                std::lock_guard<std::mutex> lk(k->mu);
                k->quit = true;
                k->cv.notify_all();
            }
            k->th.join();
        }
        {
            DeviceScope scope(k->device);
            if (k->ev0) cudaEventDestroy(k->ev0);
            if (k->ev1) cudaEventDestroy(k->ev1);
        }
        if (k->ctx) srcnn_destroy(k->ctx);
        delete k;
    }
    delete m;
    return SRCNN_OK;
}

int srcnn_mgpu_device_count(srcnn_mgpu* m) { return m ? (int)m->w.size() : SRCNN_E_ARG; }
srcnn_ctx* srcnn_mgpu_context(srcnn_mgpu* m, int i) { return (m && i >= 0 && i < (int)m->w.size()) ? m->w[i]->ctx : nullptr; }
const char* srcnn_mgpu_last_error(srcnn_mgpu* m) { return m ? m->err.c_str() : "null handle"; }

int srcnn_mgpu_last_timing(srcnn_mgpu* m, double* ms_per_worker, double* wall_ms) {
    if (!m) return SRCNN_E_ARG;
    if (ms_per_worker)
        for (size_t i = 0; i < m->w.size(); i++) ms_per_worker[i] = m->w[i]->ms;
    if (wall_ms) *wall_ms = m->wall_ms;
    return SRCNN_OK;
}

int srcnn_mgpu_band_plan(int n, int h, float scale, int i, int* r0, int* r1, int* s0, int* s1) {
    if (n <= 0 || i < 0 || i >= n || !r0 || !r1 || !s0 || !s1 || h <= 0) return SRCNN_E_ARG;
    int ow, oh;
    int rc = srcnn_out_dims(1, h, scale, &ow, &oh);
    if (rc == SRCNN_E_RATIO) {   // a one-pixel-wide probe may vanish where the real image does not: height alone decides here
        if (!(((float)h * scale) > 0.f)) return SRCNN_E_RATIO;
        oh = (int)((float)h * scale);
        if (oh <= 0) return SRCNN_E_RATIO;
    } else if (rc) return rc;
    if (!band_rows(n, oh, i, r0, r1)) { *s0 = *s1 = 0; return SRCNN_OK; }   // more workers than rows: an empty share
    return srcnn_band_src_rows(h, scale, *r0, *r1, s0, s1);
}

int srcnn_mgpu_process_batch_host(srcnn_mgpu* m, const uint8_t* src, int nframes, int w, int h, size_t src_stride,
                                  size_t src_frame_stride, int order, float scale, uint8_t* dst, size_t dst_stride,
                                  size_t dst_frame_stride) {
    if (!m) return SRCNN_E_ARG;
    std::lock_guard<std::mutex> call(m->call_mu);
    if (nframes < 0) return mfail(m, SRCNN_E_ARG, "negative frame count");
    const int n = (int)m->w.size();
    // worker i takes frames i, i+n, i+2n, ...: to its context that is a batch with n times the frame stride
    return run_all(m, [&](Worker& me, int i) {
        const int mine = nframes > i ? (nframes - i + n - 1) / n : 0;
        me.ms = 0.0;
        if (mine == 0) return (int)SRCNN_OK;
        return timed_host(me, [&] {
            return srcnn_process_batch_host(me.ctx, src + (size_t)i * src_frame_stride, mine, w, h, src_stride,
                                            src_frame_stride * (size_t)n, order, scale, dst + (size_t)i * dst_frame_stride,
                                            dst_stride, dst_frame_stride * (size_t)n);
        });
    });
}

int srcnn_mgpu_process_banded_host(srcnn_mgpu* m, const uint8_t* src, int w, int h, size_t src_stride, int order,
                                   float scale, uint8_t* dst, size_t dst_stride) {
    if (!m) return SRCNN_E_ARG;
    std::lock_guard<std::mutex> call(m->call_mu);
    int ow, oh;
    int rc = srcnn_out_dims(w, h, scale, &ow, &oh);
    if (rc) return mfail(m, rc, srcnn_strerror(rc));
    const int n = (int)m->w.size();
    return run_all(m, [&](Worker& me, int i) {
        int r0, r1;
        me.ms = 0.0;
        if (!band_rows(n, oh, i, &r0, &r1)) return (int)SRCNN_OK;
        return timed_host(me, [&] {
            return srcnn_process_band_host(me.ctx, src, w, h, src_stride, order, scale, r0, r1, dst + (size_t)r0 * dst_stride, dst_stride);
        });
    });
}

int srcnn_mgpu_process_batch_device(srcnn_mgpu* m, const uint8_t* const* d_src, const int* counts, int w, int h,
                                    size_t src_stride, size_t src_frame_stride, int order, float scale,
                                    uint8_t* const* d_dst, size_t dst_stride, size_t dst_frame_stride) {
    if (!m) return SRCNN_E_ARG;
    std::lock_guard<std::mutex> call(m->call_mu);
    if (!d_src || !d_dst || !counts) return mfail(m, SRCNN_E_ARG, "null pointer table");
    return run_all(m, [&](Worker& me, int i) {
        me.ms = 0.0;
        if (counts[i] <= 0) return (int)SRCNN_OK;
        return timed_device(me, [&] {
            return srcnn_process_batch_device(me.ctx, d_src[i], counts[i], w, h, src_stride, src_frame_stride, order, scale, d_dst[i],
                                              dst_stride, dst_frame_stride);
        });
    });
}

int srcnn_mgpu_process_banded_device(srcnn_mgpu* m, const uint8_t* const* d_src, int w, int h, size_t src_stride, int order,
                                     float scale, uint8_t* const* d_dst, size_t dst_stride) {
    if (!m) return SRCNN_E_ARG;
    std::lock_guard<std::mutex> call(m->call_mu);
    if (!d_src || !d_dst) return mfail(m, SRCNN_E_ARG, "null pointer table");
    const int n = (int)m->w.size();
    return run_all(m, [&](Worker& me, int i) {
        int r0, r1, s0, s1;
        me.ms = 0.0;
        int rc = srcnn_mgpu_band_plan(n, h, scale, i, &r0, &r1, &s0, &s1);
        if (rc) return rc;
        if (r1 <= r0) return (int)SRCNN_OK;
        return timed_device(me, [&] {
            return srcnn_process_band_device(me.ctx, d_src[i], w, h, src_stride, s0, s1, order, scale, r0, r1, d_dst[i], dst_stride);
        });
    });
}

}  // extern "C"
