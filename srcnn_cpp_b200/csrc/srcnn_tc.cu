// srcnn_tc.cu -- K-B-tc: the fused SRCNN kernel on tcgen05 tensor cores with TMEM accumulators.
//
// Replaces Convolution99x11 (src/srcnn.cpp:254-325) and Convolution55 (src/srcnn.cpp:189-243) of the
// reference with ONE persistent kernel; the 64- and 32-channel activations never leave the SM.
// Operands are FP16 (Y is exact in FP16; SURVEY Appendix C: all-FP16 operands + FP32 accumulation
// stay within 1 LSB of the reference), accumulators FP32 in TMEM, biases FP32.
//
// Mapping (B200-first, not a translation of the CPU loops):
//   * GEMM M dimension = 128 consecutive IMAGE ROWS (one TMEM lane per row); a work item walks a band
//     of 124 output rows from left to right in groups of 4 output columns.
//   * conv1 is a Toeplitz-weight GEMM straight off the staged Y tile: for kernel row i the A operand
//     is the FP16 Y tile itself (K = 16 contiguous pixels of image row r+i-4, no im2col at all), the
//     B operand is a pre-built [N = 4 cols x 64 ch][K = 16] Toeplitz image of w1[.][i][.]; 9 MMAs
//     (M128 N256 K16) accumulate one group.  The Y tile lives in shared memory in the canonical
//     no-swizzle K-major layout (8-pixel column chunks, 16 B per row), so shifting by a kernel row is
//     a +16 B change of the descriptor start address.
//   * ReLU+bias+FP16 pack run on CUDA cores TMEM->registers->TMEM (tcgen05.ld / tcgen05.st), and the
//     packed activations are fed back as the A operand FROM TMEM: conv2 = 4 x (M128 N32 K16) per
//     column, conv3 = "tap GEMM" T[p][tap] = sum_c act2[p][c]*w3[c][tap] (2 x M128 N32 K16), followed
//     by 25 shifted adds per pixel: horizontal taps accumulate in registers while the band is walked,
//     vertical taps cross lanes through a small shared-memory exchange.
//   * The reference's two border clamps are reproduced exactly: conv1 reads the replicate-clamped Y
//     (applied when the tile is staged), conv3 reads act2 AT THE CLAMPED PIXEL (src/srcnn.cpp:203,209)
//     -- implemented by folding the out-of-image taps onto the edge column/row of T, never by padding.
//   * Two independent warpgroups per CTA (each owns 256 TMEM columns, its own Y ring and barriers)
//     ping-pong on the tensor pipe: while one runs its CUDA-core epilogue the other's MMAs execute.
//     The 150 KB of packed weights are fetched once per CTA with cp.async.bulk (TMA) and shared.
//
// Executed tensor work per pixel: conv1 2*16*9*64 = 18 432 FLOP (K efficiency 9/16), conv2 4 096,
// conv3 2 048; algorithmic 16 064 FLOP/px is what bench.py reports against the roofline.
#include <cuda_fp16.h>

#include "common.h"

namespace srcnn {
namespace tc {

// ---------------------------------------------------------------------------------------------
// geometry
// ---------------------------------------------------------------------------------------------
constexpr int kBandRows = 124;                    // valid output rows per band = TMEM lanes 2..125
constexpr int kTileRows = 136;                    // staged Y rows: 128 lanes + 8 (conv1 reach)
constexpr int kChunkBytes = kTileRows * 16;       // one 8-pixel column chunk, 16 B per row
constexpr int kRingSlots = 4;                     // + slot 4 = duplicate of slot 0 (window wrap)
constexpr int kRingBytes = (kRingSlots + 1) * kChunkBytes;
constexpr int kB1Tile = 256 * 16 * 2;             // one Toeplitz B tile [256][16] fp16
constexpr int kB1Bytes = 2 * 9 * kB1Tile;         // [half][kernel row]
constexpr int kB2Bytes = 4 * 32 * 16 * 2;         // [k step][32][16]
constexpr int kB3Bytes = 2 * 32 * 16 * 2;         // [k step][32 taps (25 used)][16]
constexpr int kWeightBytes = kB1Bytes + kB2Bytes + kB3Bytes;  // 153 600
constexpr int kHxBytes = 5 * 4 * 128 * 4;         // vertical-tap exchange: [m][col][lane] fp32

constexpr int kOffW = 0;
constexpr int kOffRing = kOffW + kWeightBytes;
constexpr int kOffHx = kOffRing + 2 * kRingBytes;
constexpr int kOffBar = kOffHx + 2 * kHxBytes;    // 1 weight barrier + 2 x 3 pipeline barriers
constexpr int kOffTmem = kOffBar + 8 * 8;
constexpr int kSmemBytes = kOffTmem + 64;

__constant__ float c_b1[kC1];
__constant__ float c_b2[kC2];
__constant__ float c_b3;

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok;
}
// bounded wait: a protocol bug must surface as an error code, never as a hung GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* guard, int code) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 2000000000ll) {
            *guard = code;
            __threadfence_system();
            __trap();
        }
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// shared-memory matrix descriptor, SWIZZLE_NONE, K-major canonical layout
//   ((8,m),(8,2)) : ((16 B, SBO), (2 B, LBO))   -- 8 rows x 16 B core matrices
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
           (1ull << 46);
}
// instruction descriptor: kind::f16, A=B=FP16, D=FP32, both K-major, M=128
__host__ __device__ constexpr uint32_t idesc_f16(int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
        "l"(a), "l"(b), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
        "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

#define TM_R(v, o) "=r"(v[o + 0]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]), "=r"(v[o + 6]), "=r"(v[o + 7])
#define TM_W(v, o) "r"(v[o + 0]), "r"(v[o + 1]), "r"(v[o + 2]), "r"(v[o + 3]), "r"(v[o + 4]), "r"(v[o + 5]), "r"(v[o + 6]), "r"(v[o + 7])

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : TM_R(v, 0), TM_R(v, 8), TM_R(v, 16), TM_R(v, 24)
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
        "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
        TM_W(v, 0), TM_W(v, 8), TM_W(v, 16), TM_W(v, 24)
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        TM_W(v, 0), TM_W(v, 8)
        : "memory");
}

// ReLU + round-to-nearest FP16 + pack: low half = first element (even K index)
__device__ __forceinline__ uint32_t relu_pack_f16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

// ---------------------------------------------------------------------------------------------
// kernel parameters
// ---------------------------------------------------------------------------------------------
struct Params {
    const uint8_t* y;      // plane row 0 = image row `row0`; rows [row0, row0+rows) are present
    size_t pitch;
    int W, H;
    int row0, rows;
    int out_begin, out_end;  // image rows to produce
    uint8_t* out;          // same row0 convention as y
    size_t out_pitch;
    int out_aligned4;
    const uint8_t* wimg;   // packed FP16 operand image (kWeightBytes)
    int gpb;               // column groups per band = ceil(W/4)
    long long total_groups;
    int* guard;
};

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1) k_srcnn_tc(const Params p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x;
    const int wg = tid >> 7;        // warpgroup: an independent pipeline
    const int t = tid & 127;        // TMEM lane of this thread = image row R0 + t
    const int warp = tid >> 5;
    const uint32_t sbase = smem_u32(smem);
    const uint32_t wbar = sbase + kOffBar;
    const uint32_t mb0 = sbase + kOffBar + 8 + wg * 24, mb1 = mb0 + 8, mb2 = mb0 + 16;
    volatile uint32_t* tmem_slot = (volatile uint32_t*)(smem + kOffTmem);

    if (tid == 0) {
        mbar_init(wbar, 1);
        for (int i = 0; i < 6; i++) mbar_init(sbase + kOffBar + 8 + i * 8, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32((const void*)tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {  // weights: one TMA bulk fetch per CTA, shared by both warpgroups
        mbar_expect_tx(wbar, kWeightBytes);
        constexpr int kPiece = kWeightBytes / 4;
        for (int i = 0; i < 4; i++) bulk_g2s(sbase + kOffW + i * kPiece, p.wimg + i * kPiece, kPiece, wbar);
    }
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tm = tmem_base + wg * 256;                           // this warpgroup's 256 columns
    const uint32_t tml = tm + ((uint32_t)((warp & 3) * 32) << 16);      // + this warp's lane quarter
    const uint32_t ring = sbase + kOffRing + wg * kRingBytes;
    uint8_t* ring_p = smem + kOffRing + wg * kRingBytes;
    float* hx = (float*)(smem + kOffHx + wg * kHxBytes);
    const int bar_id = 1 + wg;
    uint32_t ph0 = 0, ph1 = 0, ph2 = 0;

    mbar_wait(wbar, 0, p.guard, 1);

    const int W = p.W, H = p.H;
    const long long nworkers = (long long)gridDim.x * 2;
    const long long wk = (long long)blockIdx.x * 2 + wg;
    long long lin = p.total_groups * wk / nworkers;
    const long long lin_end = p.total_groups * (wk + 1) / nworkers;

    while (lin < lin_end) {
        // ---- one segment: band `band`, output columns [s, e) ----
        const int band = (int)(lin / p.gpb);
        const int gfirst = (int)(lin - (long long)band * p.gpb);
        const int ng = (int)min((long long)(p.gpb - gfirst), lin_end - lin);
        lin += ng;
        const int s = gfirst * 4;
        const int e = min(W, s + ng * 4);
        const int band_begin = p.out_begin + band * kBandRows;
        const int band_end = min(p.out_end, band_begin + kBandRows);
        const int R0 = band_begin - 2;            // image row of TMEM lane 0
        const int G = (e - s + 3) / 4 + 1;        // T groups: group g covers T columns s-2+4g .. s+1+4g
        const int jlast = (G - 1) >> 1;           // last step; needs Y chunks jlast, jlast+1

        // stage one 8-pixel Y chunk (FP16, replicate-clamped) into the ring
        auto load_chunk = [&](int q) {
            const int slot = q & (kRingSlots - 1);
            const int c0 = s - 6 + 8 * q;
            for (int tr = t; tr < kTileRows; tr += 128) {
                int r = min(max(R0 - 4 + tr, 0), H - 1);
                r = min(max(r - p.row0, 0), p.rows - 1);
                const uint8_t* src = p.y + (size_t)r * p.pitch;
                uint32_t h[4];
#pragma unroll
                for (int x = 0; x < 4; x++) {
                    const int ca = min(max(c0 + 2 * x, 0), W - 1), cb = min(max(c0 + 2 * x + 1, 0), W - 1);
                    const __half2 v = __halves2half2(__int2half_rn((int)src[ca]), __int2half_rn((int)src[cb]));
                    h[x] = *reinterpret_cast<const uint32_t*>(&v);
                }
                const uint4 v4 = make_uint4(h[0], h[1], h[2], h[3]);
                *reinterpret_cast<uint4*>(ring_p + slot * kChunkBytes + tr * 16) = v4;
                if (slot == 0) *reinterpret_cast<uint4*>(ring_p + kRingSlots * kChunkBytes + tr * 16) = v4;
            }
        };
        // conv1 of group g: 9 Toeplitz MMAs, one per kernel row, accumulate into D1 = columns [0,256)
        auto issue_conv1 = [&](int g) {
            const int j = g >> 1, half = g & 1;
            const uint32_t a0 = ring + (j & (kRingSlots - 1)) * kChunkBytes;
            const uint32_t b0 = sbase + kOffW + half * 9 * kB1Tile;
#pragma unroll
            for (int i = 0; i < 9; i++)
                mma_ss(tm, smem_desc(a0 + i * 16, kChunkBytes, 128), smem_desc(b0 + i * kB1Tile, 4096, 128), idesc_f16(256), i > 0);
            mma_commit(mb0);
        };

        load_chunk(0);
        load_chunk(1);
        fence_proxy_async();
        named_bar(bar_id, 128);
        if (t == 0) {
            tc_fence_after();
            issue_conv1(0);
        }

        float carry[4][5];
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int m = 0; m < 5; m++) carry[a][m] = 0.f;

        for (int g = 0; g < G; g++) {
            const int j = g >> 1;
            if ((g & 1) == 0 && j + 2 <= jlast + 1) {  // prefetch the chunk the next step needs
                load_chunk(j + 2);
                fence_proxy_async();
            }
            // ---------------- E1: D1 -> +b1, ReLU, FP16 -> A1 (in place, columns [0,128)) ----------------
            mbar_wait(mb0, ph0, p.guard, 2);
            ph0 ^= 1;
            tc_fence_after();
#pragma unroll
            for (int d = 0; d < 4; d++) {
                uint32_t v[64], r[32];
                tmem_ld32(tml + d * 64, v);
                tmem_ld32(tml + d * 64 + 32, v + 32);
                tc_wait_ld();
#pragma unroll
                for (int c = 0; c < 32; c++)
                    r[c] = relu_pack_f16x2(__uint_as_float(v[2 * c]) + c_b1[2 * c], __uint_as_float(v[2 * c + 1]) + c_b1[2 * c + 1]);
                tmem_st32(tml + d * 32, r);
            }
            tc_wait_st();
            tc_fence_before();
            named_bar(bar_id, 128);
            if (t == 0) {  // conv2: per column d, D2[d] = A1[d] (TMEM) x W2, K = 64 in 4 steps
                tc_fence_after();
                const uint32_t b2 = sbase + kOffW + kB1Bytes;
#pragma unroll
                for (int d = 0; d < 4; d++)
#pragma unroll
                    for (int ks = 0; ks < 4; ks++)
                        mma_ts(tm + 128 + d * 32, tm + d * 32 + ks * 8, smem_desc(b2 + ks * 1024, 512, 128), idesc_f16(32), ks > 0);
                mma_commit(mb1);
            }
            // ---------------- E2: D2 -> +b2, ReLU, FP16 -> A2 (in place, columns [128,192)) ----------------
            mbar_wait(mb1, ph1, p.guard, 3);
            ph1 ^= 1;
            tc_fence_after();
#pragma unroll
            for (int d = 0; d < 4; d++) {
                uint32_t v[32], r[16];
                tmem_ld32(tml + 128 + d * 32, v);
                tc_wait_ld();
#pragma unroll
                for (int c = 0; c < 16; c++)
                    r[c] = relu_pack_f16x2(__uint_as_float(v[2 * c]) + c_b2[2 * c], __uint_as_float(v[2 * c + 1]) + c_b2[2 * c + 1]);
                tmem_st16(tml + 128 + d * 16, r);
            }
            tc_wait_st();
            tc_fence_before();
            named_bar(bar_id, 128);
            if (t == 0) {  // conv3 tap GEMM: T[d][tap] = A2[d] (TMEM) x W3, K = 32 in 2 steps -> columns [0,128)
                tc_fence_after();
                const uint32_t b3 = sbase + kOffW + kB1Bytes + kB2Bytes;
#pragma unroll
                for (int d = 0; d < 4; d++)
#pragma unroll
                    for (int ks = 0; ks < 2; ks++)
                        mma_ts(tm + d * 32, tm + 128 + d * 16 + ks * 8, smem_desc(b3 + ks * 1024, 512, 128), idesc_f16(32), ks > 0);
                mma_commit(mb2);
            }
            // ---------------- E3a: horizontal taps in registers ----------------
            mbar_wait(mb2, ph2, p.guard, 4);
            ph2 ^= 1;
            tc_fence_after();
            const int t0 = s - 2 + 4 * g;  // first T column of this group
            float acc[8][5];               // window column cw <-> image column t0-2+cw; [vertical tap m]
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int m = 0; m < 5; m++) {
                    acc[a][m] = carry[a][m];
                    acc[a + 4][m] = 0.f;
                }
#pragma unroll
            for (int d = 0; d < 4; d++) {
                uint32_t tv[32];
                tmem_ld32(tml + d * 32, tv);
                tc_wait_ld();
                const int cp = t0 + d;  // image column of this T column (warp-uniform)
                if (cp >= 0 && cp < W) {
#pragma unroll
                    for (int m = 0; m < 5; m++)
#pragma unroll
                        for (int n = 0; n < 5; n++) acc[d - n + 4][m] += __uint_as_float(tv[m * 5 + n]);
                    if (cp == 0) {  // columns -1 and -2 read act2 at column 0 (src/srcnn.cpp:209)
#pragma unroll
                        for (int m = 0; m < 5; m++) {
                            acc[d + 3][m] += __uint_as_float(tv[m * 5]);
                            acc[d + 2][m] += __uint_as_float(tv[m * 5]) + __uint_as_float(tv[m * 5 + 1]);
                        }
                    }
                    if (cp == W - 1) {  // columns W and W+1 read act2 at column W-1
#pragma unroll
                        for (int m = 0; m < 5; m++) {
                            acc[d + 2][m] += __uint_as_float(tv[m * 5 + 3]) + __uint_as_float(tv[m * 5 + 4]);
                            acc[d + 1][m] += __uint_as_float(tv[m * 5 + 4]);
                        }
                    }
                }
            }
            // vertical taps cross lanes: publish the 4 finished columns
#pragma unroll
            for (int m = 0; m < 5; m++)
#pragma unroll
                for (int a = 0; a < 4; a++) hx[(m * 4 + a) * 128 + t] = acc[a][m];
            tc_fence_before();
            named_bar(bar_id, 128);  // T fully read (D1 region reusable) and hx visible
            if (t == 0 && g + 1 < G) {
                tc_fence_after();
                issue_conv1(g + 1);  // the tensor pipe starts the next group while we finish this one
            }
            // ---------------- E3b: vertical taps, bias, truncate, clamp, store ----------------
            {
                const int row = R0 + t;
                int ln[5];
#pragma unroll
                for (int m = 0; m < 5; m++) ln[m] = min(max(min(max(row + m - 2, 0), H - 1) - R0, 0), 127);  // src/srcnn.cpp:203
                const int c0 = t0 - 2;
                const bool row_ok = (t >= 2) && (t <= 125) && (row >= band_begin) && (row < band_end);
                uint32_t pk = 0;
                int px[4];
#pragma unroll
                for (int a = 0; a < 4; a++) {
                    float sum = hx[(0 * 4 + a) * 128 + ln[0]];
#pragma unroll
                    for (int m = 1; m < 5; m++) sum += hx[(m * 4 + a) * 128 + ln[m]];
                    sum += c_b3;                                // src/srcnn.cpp:235
                    int q = (int)sum;                           // :238 truncation toward zero
                    q = min(max(q, 0), 255);
                    px[a] = q;
                    pk |= (uint32_t)q << (8 * a);
                }
                if (row_ok) {
                    uint8_t* o = p.out + (size_t)(row - p.row0) * p.out_pitch + c0;
                    if (p.out_aligned4 && c0 >= s && c0 + 3 < e) {
                        *reinterpret_cast<uint32_t*>(o) = pk;
                    } else {
#pragma unroll
                        for (int a = 0; a < 4; a++)
                            if (c0 + a >= s && c0 + a < e) o[a] = (uint8_t)px[a];
                    }
                }
            }
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int m = 0; m < 5; m++) carry[a][m] = acc[a + 4][m];
        }
        // segment done: every MMA of this warpgroup has been waited for; hx reads of the last group
        // must finish before the next segment's first hx write -> covered by its first named barriers
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------
// primitive self-test (tests/test_tc_primitives.py): one SS MMA off the Y-tile layout with a
// kernel-row offset, then ReLU/pack/tcgen05.st and one TS MMA.  Lets a wrong descriptor or TMEM
// layout assumption be told apart from a pipeline bug.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) k_tc_selftest(const uint8_t* a_tile /* 2 chunks */, const uint8_t* b1 /* 8192 B */,
                                                        const uint8_t* b2 /* 4 x 1024 B */, int row_off, float* d1_out,
                                                        float* d2_out, int* guard) {
    __shared__ __align__(1024) uint8_t sm[2 * kChunkBytes + kB1Tile + kB2Bytes + 64];
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t sb = smem_u32(sm);
    const uint32_t bar = sb + 2 * kChunkBytes + kB1Tile + kB2Bytes;
    volatile uint32_t* slot = (volatile uint32_t*)(sm + 2 * kChunkBytes + kB1Tile + kB2Bytes + 16);
    for (int i = tid; i < (2 * kChunkBytes) / 16; i += 128) ((uint4*)sm)[i] = ((const uint4*)a_tile)[i];
    for (int i = tid; i < kB1Tile / 16; i += 128) ((uint4*)(sm + 2 * kChunkBytes))[i] = ((const uint4*)b1)[i];
    for (int i = tid; i < kB2Bytes / 16; i += 128) ((uint4*)(sm + 2 * kChunkBytes + kB1Tile))[i] = ((const uint4*)b2)[i];
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(smem_u32((const void*)slot), 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = *slot;
    const uint32_t tml = tm + ((uint32_t)(warp * 32) << 16);
    if (tid == 0) {
        mma_ss(tm, smem_desc(sb + row_off * 16, kChunkBytes, 128), smem_desc(sb + 2 * kChunkBytes, 4096, 128), idesc_f16(256), 0);
        mma_commit(bar);
    }
    mbar_wait(bar, 0, guard, 10);
    tc_fence_after();
    for (int c = 0; c < 256; c += 32) {
        uint32_t v[32];
        tmem_ld32(tml + c, v);
        tc_wait_ld();
        for (int k = 0; k < 32; k++) d1_out[(size_t)tid * 256 + c + k] = __uint_as_float(v[k]);
    }
    {   // pack D1[:, 0:64] (ReLU, no bias) into A columns [256, 288)
        uint32_t v[64], r[32];
        tmem_ld32(tml, v);
        tmem_ld32(tml + 32, v + 32);
        tc_wait_ld();
        for (int c = 0; c < 32; c++) r[c] = relu_pack_f16x2(__uint_as_float(v[2 * c]), __uint_as_float(v[2 * c + 1]));
        tmem_st32(tml + 256, r);
        tc_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
        tc_fence_after();
        for (int ks = 0; ks < 4; ks++)
            mma_ts(tm + 320, tm + 256 + ks * 8, smem_desc(sb + 2 * kChunkBytes + kB1Tile + ks * 1024, 512, 128), idesc_f16(32), ks > 0);
        mma_commit(bar);
    }
    mbar_wait(bar, 1, guard, 11);
    tc_fence_after();
    {
        uint32_t v[32];
        tmem_ld32(tml + 320, v);
        tc_wait_ld();
        for (int k = 0; k < 32; k++) d2_out[(size_t)tid * 32 + k] = __uint_as_float(v[k]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 512);
}

}  // namespace tc

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static inline void put_h(uint8_t* img, size_t byte_off, float v) {
    const __half h = __float2half_rn(v);
    memcpy(img + byte_off, &h, 2);
}

// Packs the FP32 parameters into the FP16 operand images the kernel's UMMA descriptors expect
// (SWIZZLE_NONE, K-major: element (n,k) of a [N][16] tile at (k/8)*(N*16) + n*16 + (k%8)*2 bytes).
int tc_prepare_weights(Ctx* c, const float* P) {
    using namespace tc;
    std::vector<uint8_t> img(kWeightBytes, 0);
    const float* w1 = P + kOffW1;
    const float* w2 = P + kOffW2;
    const float* w3 = P + kOffW3;
    // conv1 Toeplitz tiles: n = d*64 + ch (d = output column within the group), K index k = pixel of
    // the 16-pixel window; group half h has its outputs at window pixel 4 + 4h + d.
    for (int h = 0; h < 2; h++)
        for (int i = 0; i < 9; i++) {
            const size_t base = (size_t)(h * 9 + i) * kB1Tile;
            for (int d = 0; d < 4; d++)
                for (int ch = 0; ch < 64; ch++)
                    for (int k = 0; k < 16; k++) {
                        const int tt = k - d - 4 * h;  // horizontal tap index j of w1[ch][i][j]
                        const float v = (tt >= 0 && tt <= 8) ? w1[(ch * 9 + i) * 9 + tt] : 0.f;
                        const int n = d * 64 + ch;
                        put_h(img.data(), base + (size_t)(k / 8) * 4096 + (size_t)n * 16 + (k % 8) * 2, v);
                    }
        }
    for (int ks = 0; ks < 4; ks++)  // conv2: B[n = out ch][k = in ch ks*16+k]
        for (int n = 0; n < 32; n++)
            for (int k = 0; k < 16; k++)
                put_h(img.data(), kB1Bytes + (size_t)ks * 1024 + (size_t)(k / 8) * 512 + (size_t)n * 16 + (k % 8) * 2,
                      w2[n * 64 + ks * 16 + k]);
    for (int ks = 0; ks < 2; ks++)  // conv3 tap GEMM: B[n = tap m*5+n][k = in ch]
        for (int n = 0; n < 32; n++)
            for (int k = 0; k < 16; k++)
                put_h(img.data(), kB1Bytes + kB2Bytes + (size_t)ks * 1024 + (size_t)(k / 8) * 512 + (size_t)n * 16 + (k % 8) * 2,
                      n < 25 ? w3[(ks * 16 + k) * 25 + n] : 0.f);
    SRCNN_CUDA(c, cudaMalloc(&c->d_tc_weights, kWeightBytes));
    c->tc_weights_bytes = kWeightBytes;
    SRCNN_CUDA(c, cudaMemcpy(c->d_tc_weights, img.data(), kWeightBytes, cudaMemcpyHostToDevice));
    SRCNN_CUDA(c, cudaMemcpyToSymbol(tc::c_b1, P + kOffB1, sizeof(float) * kC1));
    SRCNN_CUDA(c, cudaMemcpyToSymbol(tc::c_b2, P + kOffB2, sizeof(float) * kC2));
    SRCNN_CUDA(c, cudaMemcpyToSymbol(tc::c_b3, P + kOffB3, sizeof(float)));
    SRCNN_CUDA(c, cudaFuncSetAttribute(tc::k_srcnn_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    return SRCNN_OK;
}

void tc_release(Ctx* c) {
    if (c->d_tc_weights) cudaFree(c->d_tc_weights);
    c->d_tc_weights = nullptr;
}

int launch_cnn_tc(Ctx* c, const CnnArgs& a) {
    using namespace tc;
    if (a.out_end <= a.out_begin) return SRCNN_OK;
    Params p;
    p.y = a.y; p.pitch = a.pitch;
    p.W = a.W; p.H = a.H;
    p.row0 = a.row0; p.rows = a.rows;
    p.out_begin = a.out_begin; p.out_end = a.out_end;
    p.out = a.out; p.out_pitch = a.out_pitch;
    p.out_aligned4 = ((((uintptr_t)a.out) | a.out_pitch) & 3) == 0;
    p.wimg = (const uint8_t*)c->d_tc_weights;
    p.gpb = (a.W + 3) / 4;
    const int nbands = (a.out_end - a.out_begin + kBandRows - 1) / kBandRows;
    p.total_groups = (long long)nbands * p.gpb;
    p.guard = c->d_guard;
    // one persistent CTA per SM; fewer when the image is too small to give every warpgroup ~8 groups
    long long want = (p.total_groups + 15) / 16;
    int grid = (int)std::min<long long>(c->sm_count, std::max<long long>(1, want));
    k_srcnn_tc<<<grid, 256, kSmemBytes, c->stream>>>(p);
    c->launches++;
    SRCNN_CUDA(c, cudaGetLastError());
    return SRCNN_OK;
}

}  // namespace srcnn

// test hook (not part of the stable ABI; exported for tests/test_tc_primitives.py)
extern "C" __attribute__((visibility("default"))) int srcnn_debug_tc_selftest(srcnn_ctx* c, const void* d_a_tile,
                                                                              const void* d_b1, const void* d_b2,
                                                                              int row_off, float* d_d1, float* d_d2) {
    if (!c) return SRCNN_E_ARG;
    cudaSetDevice(c->device);
    srcnn::tc::k_tc_selftest<<<1, 128, 0, c->stream>>>((const uint8_t*)d_a_tile, (const uint8_t*)d_b1, (const uint8_t*)d_b2,
                                                      row_off, d_d1, d_d2, c->d_guard);
    c->launches++;
    SRCNN_CUDA(c, cudaGetLastError());
    return SRCNN_OK;
}
