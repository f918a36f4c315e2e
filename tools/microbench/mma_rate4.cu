// (issue pattern: warp-uniform branch + elect.sync, so ptxas emits back-to-back UTCHMMA with uniform-register
// operands instead of the ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall loop that a (tid & 31) == 0 branch gets)
// Microbenchmark: is the ~120-cycle "cost per tcgen05.mma" an ISSUE cost of the thread or the latency of a
// dependent accumulate chain?  One or several issuing threads rotate over `nacc` independent accumulators.
// Second part: the MMA mix of one image row of the row-walking SRCNN mapping (6 x TS N64 chain, 5 x TS N32
// chain, 2 x TS N32 chain) rotated over U independent units, issued by 1 or 3 threads.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t idesc(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | (8u << 24); }
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t id, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit_wait(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(0) : "memory");
}
// mode 0: uniform stream.  each issuer owns nacc accumulators of n columns and rotates over them; chain = consecutive
//         MMAs into the same accumulator before moving on (chain=1: round robin)
// mode 1: SRCNN row mix over U units (unit = 64 TMEM columns), nissuers 1 or 3
__global__ void k(int mode, int n, int ts, int reps, long long* out, int nissuers, int nacc, int chain) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint32_t slot;
    __shared__ __align__(8) uint64_t bars[4];
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // provably warp-uniform: operands stay in uniform registers
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
    if (warp < nissuers) {
      uint32_t leader;
      asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.b32 %0, 1, 0, P1;\n\t}" : "=r"(leader));
      if (leader) {
        const uint32_t sb = smem_u32(sm);
        long long t0 = clock64(), t1;
        if (mode == 0) {
            const uint64_t ad = smem_desc(sb, 2176, 128), bd = smem_desc(sb + 8192, (uint32_t)n * 16, 128);
            const uint32_t id = idesc(n);
            const uint32_t d0 = tm + warp * (nacc * n);
            for (int r = 0; r < reps; r++) {
                const uint32_t d = d0 + ((r / chain) % nacc) * n;
                if (ts) mma_ts(d, tm + 448 + (r & 7) * 8, bd, id, r >= nacc * chain);
                else mma_ss(d, ad, bd, id, r >= nacc * chain);
            }
            t1 = clock64();
        } else {
            const int U = nacc;
            const uint64_t b64 = smem_desc(sb + 8192, 64 * 16, 128), b32 = smem_desc(sb + 16384, 32 * 16, 128);
            const uint32_t id64 = idesc(64), id32 = idesc(32);
            for (int r = 0; r < reps; r++) {
                const uint32_t u = tm + (r % U) * 64;
                if (nissuers == 1 || warp == 0)
                    for (int i = 0; i < n; i++) mma_ts(u, tm + 448 + i * 8, b64, id64, i > 0);          // conv1: ring (48+ cols) -> D1
                if (nissuers == 1 || warp == 1)
                    for (int i = 0; i < 5; i++) mma_ts(u + 32, tm + 448 + i * 8, b32, id32, i > 0);     // conv2
                if (nissuers == 1 || warp == 2)
                    for (int i = 0; i < 2; i++) mma_ts(u, tm + 448 + i * 8, b32, id32, i > 0);          // conv3
            }
            t1 = clock64();
        }
        commit_wait(smem_u32(&bars[warp]));
        long long t2 = clock64();
        out[2 * warp] = t1 - t0;
        out[2 * warp + 1] = t2 - t0;
      }
      __syncwarp();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
}
int main() {
    long long* d; cudaMalloc(&d, 64);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    auto run = [&](int mode, int n, int ts, int reps, int ni, int nacc, int chain) {
        k<<<1, 128, 65536>>>(mode, n, ts, reps, d, ni, nacc, chain);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); exit(1); }
        long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
        long long mx = 0, is = 0; for (int w = 0; w < ni; w++) { mx = h[2*w+1] > mx ? h[2*w+1] : mx; is = h[2*w] > is ? h[2*w] : is; }
        if (mode == 0)
            printf("%s N=%3d issuers=%d nacc=%d chain=%d: issue loop %6lld clk, total %6lld clk -> %.1f clk per MMA (pipe ideal %.0f)\n",
                   ts ? "TS" : "SS", n, ni, nacc, chain, is, mx, (double)mx / (reps * ni), 128.0 * n / 256);
        else
            printf("ROWMIX conv1 chunks=%d issuers=%d units=%d: issue loop %6lld clk, total %6lld clk -> %.1f clk per row (pipe ideal %.0f)\n",
                   n, ni, nacc, is, mx, (double)mx / reps, n * 32.0 + 7 * 16.0);
    };
    for (int rep = 0; rep < 2; rep++) {   // second pass = warm
        for (int n : {32, 64}) for (int nacc : {1, 2, 4, 8}) if (nacc * n <= 448) run(0, n, 1, 128, 1, nacc, 1);
        run(0, 32, 1, 128, 1, 4, 2);
        run(0, 32, 1, 128, 1, 4, 4);
        run(0, 64, 1, 126, 1, 4, 6);
        run(0, 64, 1, 126, 1, 6, 6);
        for (int nacc : {1, 2, 3}) run(0, 128, 0, 126, 1, nacc, 1);
        run(0, 32, 1, 128, 2, 4, 1);
        run(0, 32, 1, 128, 4, 2, 1);
        run(0, 32, 1, 128, 4, 3, 1);
        run(0, 64, 1, 128, 3, 2, 1);
        for (int U : {1, 2, 4, 7}) { run(1, 6, 1, 70, 1, U, 0); run(1, 6, 1, 70, 3, U, 0); }
        run(1, 7, 1, 70, 1, 7, 0); run(1, 7, 1, 70, 3, 7, 0);
        printf("----\n");
    }
    return 0;
}
