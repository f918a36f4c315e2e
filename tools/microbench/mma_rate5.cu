// Clean issue-rate microbenchmark: warp-uniform control flow + elect.sync, operands in uniform registers, no per-iteration
// integer division; the issue loop is unrolled x8 over power-of-two accumulator counts.  Verify with
//   cuobjdump -sass mma_rate5 | grep -E "UTCHMMA|BRA"   (back-to-back UTCHMMA, no BRA.U.ANY waterfall)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__host__ __device__ constexpr uint32_t idesc(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | (8u << 24); }
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
// MODE 0: TS stream, MODE 1: SS stream, MODE 2: SRCNN row mix (C1 conv1 chunks N64 + 5 N32 + 2 N32 per row, U units)
template <int MODE, int N, int NACC, int C1>
__global__ void k(int reps, long long* out, int nissuers) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint32_t slot;
    __shared__ __align__(8) uint64_t bars[4];
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
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
            if (MODE < 2) {
                const uint64_t ad = smem_desc(sb, 2176, 128), bd = smem_desc(sb + 8192, (uint32_t)N * 16, 128);
                const uint32_t d0 = tm + warp * (NACC * N);
                for (int r = 0; r < reps; r += 8) {
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        const uint32_t d = d0 + (q % NACC) * N;
                        if (MODE == 0) mma_ts(d, tm + 448 + q * 8, bd, idesc(N), 1);
                        else mma_ss(d, ad, bd, idesc(N), 1);
                    }
                }
                t1 = clock64();
            } else {
                const uint64_t b64 = smem_desc(sb + 8192, 64 * 16, 128), b32 = smem_desc(sb + 16384, 32 * 16, 128);
                for (int r = 0; r < reps; r += NACC) {
#pragma unroll
                    for (int q = 0; q < NACC; q++) {
                        const uint32_t u = tm + q * 64;
                        if (nissuers == 1 || warp == 0)
#pragma unroll
                            for (int i = 0; i < C1; i++) mma_ts(u, tm + 448 + i * 8, b64, idesc(64), i > 0);
                        if (nissuers == 1 || warp == 1)
#pragma unroll
                            for (int i = 0; i < 5; i++) mma_ts(u + 32, tm + 448 + i * 8, b32, idesc(32), i > 0);
                        if (nissuers == 1 || warp == 2)
#pragma unroll
                            for (int i = 0; i < 2; i++) mma_ts(u, tm + 448 + i * 8, b32, idesc(32), i > 0);
                    }
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
long long* d;
template <int MODE, int N, int NACC, int C1>
void run(int reps, int ni) {
    cudaFuncSetAttribute(k<MODE, N, NACC, C1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    for (int pass = 0; pass < 2; pass++) {
        k<MODE, N, NACC, C1><<<1, 128, 65536>>>(reps, d, ni);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); exit(1); }
    }
    long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    long long mx = 0, is = 0; for (int w = 0; w < ni; w++) { mx = h[2*w+1] > mx ? h[2*w+1] : mx; is = h[2*w] > is ? h[2*w] : is; }
    if (MODE < 2)
        printf("%s N=%3d issuers=%d nacc=%d: issue loop %6lld clk, total %6lld clk -> %.1f clk per MMA (pipe ideal %.0f)\n",
               MODE == 0 ? "TS" : "SS", N, ni, NACC, is, mx, (double)mx / (reps * ni), 128.0 * N / 256);
    else
        printf("ROWMIX conv1 chunks=%d issuers=%d units=%d: issue loop %6lld clk, total %6lld clk -> %.1f clk per row (pipe ideal %.0f)\n",
               C1, ni, NACC, is, mx, (double)mx / reps, C1 * 32.0 + 7 * 16.0);
}
int main() {
    cudaMalloc(&d, 64);
    run<0, 8, 1, 0>(256, 1);  run<0, 8, 4, 0>(256, 1);
    run<0, 16, 1, 0>(256, 1); run<0, 16, 4, 0>(256, 1);
    run<0, 32, 1, 0>(256, 1); run<0, 32, 2, 0>(256, 1); run<0, 32, 4, 0>(256, 1); run<0, 32, 8, 0>(256, 1);
    run<0, 32, 2, 0>(256, 2); run<0, 32, 2, 0>(256, 4);
    run<0, 64, 1, 0>(256, 1); run<0, 64, 2, 0>(256, 1); run<0, 64, 4, 0>(256, 1); run<0, 64, 2, 0>(256, 2);
    run<0, 128, 1, 0>(256, 1); run<0, 128, 2, 0>(256, 1);
    run<0, 256, 1, 0>(256, 1);
    run<1, 32, 1, 0>(256, 1); run<1, 32, 4, 0>(256, 1);
    run<1, 64, 1, 0>(256, 1); run<1, 64, 4, 0>(256, 1);
    run<1, 128, 1, 0>(256, 1); run<1, 128, 2, 0>(256, 1); run<1, 128, 2, 0>(256, 2);
    run<1, 256, 1, 0>(256, 1);
    run<2, 0, 1, 6>(64, 1); run<2, 0, 2, 6>(64, 1); run<2, 0, 4, 6>(64, 1); run<2, 0, 4, 6>(64, 3);
    run<2, 0, 4, 7>(64, 1); run<2, 0, 4, 7>(64, 3);
    return 0;
}
