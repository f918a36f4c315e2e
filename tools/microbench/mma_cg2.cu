// Microbenchmark + semantics check for tcgen05.mma.cta_group::2 (one instruction feeds an SM pair).
//   - D[cta][lane][n] for A = (rank+1) everywhere, B = 1.0 in rank 0's half-tile and 2.0 in rank 1's:
//     tells which CTA supplies which N half and that each CTA's A comes from its own shared memory.
//   - cycles per MMA (leader issues) for SS N=128/256 and TS N=32 with 1/2/4 issuing threads.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t idesc(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | (16u << 24); }  // M = 256
#define TM_R(v, o) "=r"(v[o + 0]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]), "=r"(v[o + 6]), "=r"(v[o + 7])
__global__ void __cluster_dims__(2, 1, 1) k(int n, int ts, int reps, int nissuers, long long* out, float* dump) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint32_t slot;
    __shared__ __align__(8) uint64_t bars[4];
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t rank = cluster.block_rank();
    const int warp = threadIdx.x >> 5;
    // A tile at sm[0..8192): fp16 (rank+1); B half-tile at sm[8192..): fp16 (rank ? 2 : 1)
    const __half av = __float2half((float)(rank + 1)), bv = __float2half(rank ? 2.f : 1.f);
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) ((__half*)sm)[i] = av;
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) ((__half*)(sm + 8192))[i] = bv;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[i])), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster.sync();
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster.sync();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = slot;
    if (rank == 0 && (threadIdx.x & 31) == 0 && warp < nissuers) {
        const uint32_t sb = smem_u32(sm);
        const uint64_t ad = smem_desc(sb, 2048, 128), bd = smem_desc(sb + 8192, (uint32_t)(n / 2) * 16, 128);
        const uint32_t id = idesc(n);
        const uint32_t tmw = tm + warp * 128 * (n <= 128 ? 1 : 0);
        long long t0 = clock64();
        for (int r = 0; r < reps; r++) {
            if (ts)
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmw), "r"(tm + 384 + (r & 7) * 8), "l"(bd), "r"(id), "r"(r > 0 ? 1 : 0) : "memory");
            else
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmw), "l"(ad), "l"(bd), "r"(id), "r"(r > 0 ? 1 : 0) : "memory");
        }
        long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bars[warp])), "h"((uint16_t)3) : "memory");
        out[2 * warp] = t1 - t0;
    }
    // every CTA waits on ITS copy of each issuer's barrier
    if ((threadIdx.x & 31) == 0 && warp < nissuers) {
        uint32_t ok = 0;
        long long t0 = clock64();
        while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bars[warp])), "r"(0) : "memory");
        if (rank == 0) out[2 * warp + 1] = clock64() - t0 + out[2 * warp];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (dump && warp < 4) {   // dump D[lane][0..31] and D[lane][n-32..n-1] of this CTA
        uint32_t v[8];
        const uint32_t tml = tm + ((uint32_t)(warp * 32) << 16);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : TM_R(v, 0) : "r"(tml) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        dump[(rank * 128 + threadIdx.x) * 2 + 0] = __uint_as_float(v[0]);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : TM_R(v, 0) : "r"(tml + n - 8) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        dump[(rank * 128 + threadIdx.x) * 2 + 1] = __uint_as_float(v[7]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster.sync();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
}
int main() {
    long long* d; float* dump;
    cudaMalloc(&d, 64); cudaMalloc(&dump, 256 * 2 * 4);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    // semantics: one SS MMA, N = 128
    k<<<2, 128, 65536>>>(128, 0, 1, 1, d, dump);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
    float h[512]; cudaMemcpy(h, dump, sizeof(h), cudaMemcpyDeviceToHost);
    printf("semantics (A=rank+1, B=1 in rank0's half, 2 in rank1's; K=16): cta0 lane0 D[0]=%g D[N-1]=%g | cta0 lane127 D[0]=%g D[N-1]=%g | cta1 lane0 D[0]=%g D[N-1]=%g\n",
           h[0], h[1], h[127 * 2], h[127 * 2 + 1], h[256], h[257]);
    struct Cfg { int n, ts, ni; } cfgs[] = {{128,0,1},{128,0,2},{128,0,4},{256,0,1},{256,0,2},{32,1,1},{32,1,2},{32,1,4},{64,0,4},{32,0,4}};
    for (auto c : cfgs) {
        const int reps = 64;
        k<<<2, 128, 65536>>>(c.n, c.ts, reps, c.ni, d, nullptr);
        e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
        long long hh[8]; cudaMemcpy(hh, d, 64, cudaMemcpyDeviceToHost);
        long long mx = 0; for (int w = 0; w < c.ni; w++) mx = hh[2*w+1] > mx ? hh[2*w+1] : mx;
        printf("cta_group::2 %s N=%3d issuers=%d: total %6lld clk -> %.1f clk per MMA (per-SM pipe ideal %.0f)\n", c.ts ? "TS" : "SS", c.n, c.ni, mx, (double)mx / (reps * c.ni), 128.0*c.n/256);
    }
    return 0;
}
