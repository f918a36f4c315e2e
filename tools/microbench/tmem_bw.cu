// Microbenchmark: tcgen05.ld / tcgen05.st throughput per SM as a function of warp count and shape.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define TM_R(v, o) "=r"(v[o + 0]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]), "=r"(v[o + 6]), "=r"(v[o + 7])
#define TM_W(v, o) "r"(v[o + 0]), "r"(v[o + 1]), "r"(v[o + 2]), "r"(v[o + 3]), "r"(v[o + 4]), "r"(v[o + 5]), "r"(v[o + 6]), "r"(v[o + 7])
__device__ __forceinline__ void ld32(uint32_t a, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : TM_R(v, 0), TM_R(v, 8), TM_R(v, 16), TM_R(v, 24) : "r"(a) : "memory");
}
__device__ __forceinline__ void ld16(uint32_t a, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : TM_R(v, 0), TM_R(v, 8) : "r"(a) : "memory");
}
__device__ __forceinline__ void st16(uint32_t a, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(a), TM_W(v, 0), TM_W(v, 8) : "memory");
}
__global__ void k(int mode, int iters, long long* out, uint32_t* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = slot + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t v[32], acc = 0;
    for (int i = 0; i < 32; i++) v[i] = threadIdx.x + i;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        const uint32_t col = ((it * 32) + (warp >> 2) * 64) & 480;
        if (mode == 0) {            // ld x32, wait every load
            ld32(tm + col, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc += v[0] + v[31];
        } else if (mode == 1) {     // 4 x ld x32 in flight... registers limit: reuse v (values discarded)
            ld32(tm + col, v); ld32(tm + ((col + 32) & 480), v); ld32(tm + ((col + 64) & 480), v); ld32(tm + ((col + 96) & 480), v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc += v[0] + v[31];
        } else if (mode == 2) {     // ld x16, wait every load
            ld16(tm + col, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc += v[0] + v[15];
        } else if (mode == 3) {     // st x16, wait every store
            st16(tm + col, v);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        } else {                    // 4 x st x16 then wait
            st16(tm + col, v); st16(tm + ((col + 16) & 496), v); st16(tm + ((col + 32) & 496), v); st16(tm + ((col + 48) & 496), v);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
    }
    long long t1 = clock64();
    if (threadIdx.x % 32 == 0) out[blockIdx.x * 32 + warp] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
}
int main() {
    long long* d; uint32_t* s;
    cudaMalloc(&d, 148 * 32 * 8); cudaMalloc(&s, 148 * 1024 * 4);
    const int iters = 2000;
    const char* names[] = {"ld.x32 wait-each", "4x ld.x32 per wait", "ld.x16 wait-each", "st.x16 wait-each", "4x st.x16 per wait"};
    const int ops[] = {1, 4, 1, 1, 4};
    const int bytes[] = {4096, 4096, 2048, 2048, 2048};
    for (int mode = 0; mode < 5; mode++)
        for (int warps : {1, 4, 8, 16}) {
            k<<<1, warps * 32, 0>>>(mode, iters, d, s);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
            long long h[32]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            long long mx = 0; for (int w = 0; w < warps; w++) mx = h[w] > mx ? h[w] : mx;
            double per_op = (double)mx / (iters * ops[mode]);
            printf("%-22s warps=%2d  cycles/op/warp=%7.1f  SM bytes/clk=%7.1f\n", names[mode], warps, per_op,
                   (double)warps * iters * ops[mode] * bytes[mode] / mx);
        }
    return 0;
}
