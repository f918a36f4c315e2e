// Microbenchmark: tcgen05.mma issue/execute time per instruction vs N, for SS (A in smem) and TS (A in TMEM).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t idesc(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | (8u << 24); }
__global__ void k(int n, int ts, int reps, long long* out, int nissuers, int alternate) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint32_t slot;
    __shared__ __align__(8) uint64_t bars[4];
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) ((uint32_t*)sm)[i] = 0x3c003c00u;  // fp16 1.0
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[i])), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = slot;
    if ((threadIdx.x & 31) == 0 && warp < nissuers) {
        const uint32_t sb = smem_u32(sm);
        const uint32_t tmw = tm + warp * 128;
        const uint64_t ad = smem_desc(sb, 2176, 128), bd = smem_desc(sb + 8192, (uint32_t)n * 16, 128);
        const uint32_t id = idesc(n);
        long long t0 = clock64();
        for (int r = 0; r < reps; r++) {
            if (ts)
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmw + (alternate ? (r & 1) * 32 : 0)), "r"(tm + 384 + (r & 7) * 8), "l"(bd), "r"(id), "r"(r > 0 ? 1 : 0) : "memory");
            else
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmw + (alternate ? (r & 1) * 32 : 0)), "l"(ad), "l"(bd), "r"(id), "r"(r > 1 ? 1 : 0) : "memory");
        }
        long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[warp])) : "memory");
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bars[warp])), "r"(0) : "memory");
        long long t2 = clock64();
        out[2 * warp] = t1 - t0;
        out[2 * warp + 1] = t2 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
}
int main() {
    long long* d; cudaMalloc(&d, 64);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    struct Cfg { int n, ts, ni; } cfgs[] = {{128,0,1},{128,0,2},{128,0,4},{64,0,1},{64,0,2},{64,0,4},{64,1,4},{32,1,4},{32,0,4},{16,1,4},{128,1,1},{128,1,2}};
    for (auto c : cfgs) {
        const int reps = 64;
        k<<<1, 128, 65536>>>(c.n, c.ts, reps, d, c.ni, 0);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
        long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
        long long mx = 0; for (int w = 0; w < c.ni; w++) mx = h[2*w+1] > mx ? h[2*w+1] : mx;
        printf("%s N=%3d issuers=%d: total %6lld clk -> %.1f clk per MMA overall (pipe ideal %.0f)\n", c.ts ? "TS" : "SS", c.n, c.ni, mx, (double)mx / (reps * c.ni), 128.0*c.n/256);
    }
    return 0;
}
