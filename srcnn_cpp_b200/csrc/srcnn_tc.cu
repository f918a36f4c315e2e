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
//   * Biases ride on the tensor pipe too: one extra K=16 MMA whose A operand is a constant "ones" tile
//     and whose B operand holds hi+lo FP16 halves of the FP32 bias (error < 1e-4), so the epilogues
//     are pure ReLU+pack (cvt.rn.relu.f16x2.f32).
//   * Two independent pipelines per CTA (one epilogue warpgroup each, one thread per TMEM lane; each
//     owns 256 TMEM columns, its own Y ring and barriers) share the tensor pipe.  Inside a pipeline a
//     group is two independently committed halves (2 columns each, 128 TMEM columns each) that the
//     epilogue threads process interleaved, so the MMA -> epilogue -> MMA round trips of one half hide
//     behind the CUDA-core work of the other.  Twelve dedicated issuer warps (a tcgen05.mma costs its
//     issuing thread ~100+ cycles) keep the queue fed; epilogue TMEM loads are software-pipelined
//     (tcgen05.ld round trips are ~150 cycles).  The 163 KB of packed operands are fetched once per
//     CTA with cp.async.bulk (TMA) and shared by both pipelines.
//
// Executed tensor work per pixel: conv1 2*16*(9+1)*64 = 20 480 FLOP (K efficiency 9/16, +1 bias MMA),
// conv2 5 120, conv3 2 048; algorithmic 16 064 FLOP/px is what bench.py reports against the roofline.
#include <cuda_fp16.h>
#include <cstdlib>

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
constexpr int kBias1Bytes = kB1Tile;              // [256][16]: k=0 hi(b1), k=1 lo(b1)  (bias as one more MMA)
constexpr int kBias2Bytes = 32 * 16 * 2;          // [32][16]:  k=0 hi(b2), k=1 lo(b2)
constexpr int kOnesBytes = 2 * 128 * 16;          // A operand of the bias MMAs: chunk0 = [1,1,0..0] per row, chunk1 = 0
constexpr int kWeightBytes = kB1Bytes + kB2Bytes + kB3Bytes + kBias1Bytes + kBias2Bytes + kOnesBytes;  // 166 912
constexpr int kNV = 6;                            // vertical-tap partial planes: m0,m1,m2,(m3,n0),(m3,n1..4),m4
constexpr int kHxBytes = kNV * 4 * 128 * 4;       // vertical-tap exchange: [plane][col][lane] fp32

constexpr int kOffW = 0;
constexpr int kImgB2 = kOffW + kB1Bytes;
constexpr int kImgB3 = kImgB2 + kB2Bytes;
constexpr int kOffBias1 = kImgB3 + kB3Bytes;
constexpr int kOffBias2 = kOffBias1 + kBias1Bytes;
constexpr int kOffOnes = kOffBias2 + kBias2Bytes;
constexpr int kOffRing = kOffW + kWeightBytes;
constexpr int kOffHx = kOffRing + 2 * kRingBytes;
constexpr int kOffBar = kOffHx + 2 * kHxBytes;    // 1 weight barrier + 2 x 3 pipeline barriers
constexpr int kOffTmem = kOffBar + 32 * 8;
constexpr int kThreads = 640;                    // 2 epilogue warpgroups + 3 warpgroups of MMA-issuer warps
constexpr int kSmemBytes = kOffTmem + 64;
static_assert(kWeightBytes % 64 == 0 && kSmemBytes <= 227 * 1024, "shared memory budget");

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
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
// try_wait with a suspend-time hint: the hardware parks the thread until the phase completes or ~hint_ns pass,
// instead of returning at once -- a spinning waiter steals issue slots from the epilogue warps of its SM sub-partition
// (ncu: the twelve issuer warps alone executed more instructions spinning than the epilogues did working)
__device__ __forceinline__ uint32_t mbar_try_wait_hint(uint32_t bar, uint32_t parity, uint32_t hint_ns) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(hint_ns)
        : "memory");
    return ok;
}
// bounded wait: a protocol bug must surface as an error code, never as a hung GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* guard, int code) {
    if (mbar_try_wait(bar, parity)) return;
    uint32_t tries = 0;
    while (!mbar_try_wait_hint(bar, parity, 20000u)) {
        if (++tries > 200000u) {   // >> any legitimate wait (each try parks for up to 20 us)
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
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.b32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
    return pred;
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

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : TM_R(v, 0), TM_R(v, 8)
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), TM_W(v, 0) : "memory");
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
    int y_aligned2;
    const uint8_t* wimg;   // packed FP16 operand image (kWeightBytes)
    int gpb;               // column groups per band = ceil(W/4)
    long long total_groups;
    int* guard;
    long long* dbg;        // optional timeline (SRCNN_TC_DEBUG=1): clock64 stamps of CTA 0
};
#define TL(slot) do { if (p.dbg && blockIdx.x == 0 && tp == 0 && g < 24) p.dbg[((pipe * 24 + g) * 16) + (slot)] = clock64(); } while (0)

// One horizontal tap of conv3 (compile-time column D of the group, horizontal tap N, vertical tap M):
// adds T to the output-window column it belongs to, plus the reference's clamped reads at the image's
// left/right edge (act2[clamp(c+n-2)], src/srcnn.cpp:209): an edge column stands in for its
// out-of-image neighbours.
template <int D, int N, int M>
__device__ __forceinline__ void tap_add(float (&acc)[8][5], float val, bool left, bool right) {
    acc[D - N + 4][M] += val;
    if (left) {
        if (N == 0) { acc[D + 3][M] += val; acc[D + 2][M] += val; }
        if (N == 1) { acc[D + 2][M] += val; }
    }
    if (right) {
        if (N == 3) { acc[D + 2][M] += val; }
        if (N == 4) { acc[D + 2][M] += val; acc[D + 1][M] += val; }
    }
}
template <int D>
__device__ __forceinline__ void taps_col(float (&acc)[8][5], const uint32_t* tv, bool l, bool r) {
#define TAPROW(M_)                                                   \
    tap_add<D, 0, M_>(acc, __uint_as_float(tv[M_ * 5 + 0]), l, r);   \
    tap_add<D, 1, M_>(acc, __uint_as_float(tv[M_ * 5 + 1]), l, r);   \
    tap_add<D, 2, M_>(acc, __uint_as_float(tv[M_ * 5 + 2]), l, r);   \
    tap_add<D, 3, M_>(acc, __uint_as_float(tv[M_ * 5 + 3]), l, r);   \
    tap_add<D, 4, M_>(acc, __uint_as_float(tv[M_ * 5 + 4]), l, r);
    TAPROW(0) TAPROW(1) TAPROW(2) TAPROW(3) TAPROW(4)
#undef TAPROW
}

// ---------------------------------------------------------------------------------------------
// the kernel: 20 warps.
//   warps 0-3 / 4-7   : epilogue warpgroups of pipeline 0 / 1 (one thread per TMEM lane = image row)
//   warps 8-19        : MMA issuers, lane 0 of each: per pipeline two conv1 issuers (one per half of the
//                       group) and four conv2/conv3 issuers (one per output column).  A tcgen05.mma costs
//                       its issuing thread ~100-120 cycles whatever its size, so issue is spread over
//                       threads that do nothing else.
// A group of 4 output columns is processed as two independently committed HALVES (columns 0,1 and 2,3),
// each living in its own 128 TMEM columns.  The epilogue threads run the halves interleaved
// (E1a E1b E2a E2b E3a E3b), so the MMA -> epilogue -> MMA round trips of one half hide behind the
// CUDA-core work of the other; two such pipelines per SM share the tensor pipe.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1) k_srcnn_tc(const Params p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x;
    // warp index through a shuffle: provably warp-uniform, so ptxas keeps everything derived from it (roles, TMEM
    // addresses, descriptors) in uniform registers and emits back-to-back UTCHMMA.  With a (tid & 31) == 0 branch it
    // wraps EVERY tcgen05.mma in an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall loop (~100 cycles per MMA).
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const bool issuer = warp >= 8;
    const int irole = issuer ? (warp - 8) % 6 : 0;        // 0,1: conv1 half a/b; 2..5: conv2+conv3 of column irole-2
    const int pipe = issuer ? (warp - 8) / 6 : (tid >> 7);
    const int tp = tid & 127;        // thread within the pipeline = TMEM lane = image row R0 + tp
    const int quarter = warp & 3;    // TMEM lane quarter this warp may access
    const uint32_t sbase = smem_u32(smem);
    const uint32_t wbar = sbase + kOffBar;
    // per pipeline, per half h: mb0[h], mb1[h], mb2[h] = "conv1/2/3 MMAs of the half complete" (tcgen05.commit);
    // rq0[h], rq1[h], rq2[h] = "operands of conv1/2/3 of the half are ready" (128 epilogue arrivals)
    const uint32_t bars = sbase + kOffBar + 8 + pipe * 96;
    auto MB = [&](int stage, int h) { return bars + (uint32_t)(stage * 2 + h) * 8; };
    auto RQ = [&](int stage, int h) { return bars + 48 + (uint32_t)(stage * 2 + h) * 8; };
    volatile uint32_t* tmem_slot = (volatile uint32_t*)(smem + kOffTmem);

    if (tid == 0) {
        mbar_init(wbar, 1);
        for (int q = 0; q < 2; q++)
            for (int i = 0; i < 12; i++)
                mbar_init(sbase + kOffBar + 8 + q * 96 + i * 8, i < 2 ? 1 : (i < 6 ? 2 : 128));
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32((const void*)tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {  // weights: one TMA bulk fetch per CTA, shared by both pipelines
        mbar_expect_tx(wbar, kWeightBytes);
        constexpr int kPiece = kWeightBytes / 4;
        static_assert(kPiece % 16 == 0, "bulk copy granularity");
        for (int i = 0; i < 4; i++) bulk_g2s(sbase + kOffW + i * kPiece, p.wimg + i * kPiece, kPiece, wbar);
    }
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tm = tmem_base + pipe * 256;                         // this pipeline's 256 columns
    const uint32_t ring = sbase + kOffRing + pipe * kRingBytes;

    const int W = p.W, H = p.H;
    const long long nworkers = (long long)gridDim.x * 2;
    const long long wk = (long long)blockIdx.x * 2 + pipe;
    long long lin = p.total_groups * wk / nworkers;
    const long long lin_end = p.total_groups * (wk + 1) / nworkers;

    if (issuer) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");   // the three issuer warpgroups hand registers to the epilogues
        const uint32_t leader = elect_one();   // the whole warp walks the loop (uniform control flow); one lane issues
        {
            mbar_wait(wbar, 0, p.guard, 1);
            uint32_t ph = 0;   // every request barrier completes exactly once per group
            while (lin < lin_end) {
                const int band = (int)(lin / p.gpb);
                const int gfirst = (int)(lin - (long long)band * p.gpb);
                const int ng = (int)min((long long)(p.gpb - gfirst), lin_end - lin);
                lin += ng;
                const int s = gfirst * 4;
                const int e = min(W, s + ng * 4);
                const int G = (e - s + 3) / 4 + 1;
                if (irole < 2) {
                    // ---- conv1 of one half: bias MMA + 9 Toeplitz MMAs (M128 N128 K16) into D1 of the half ----
                    const int h = irole;
                    const uint32_t dcol = tm + h * 128;
                    const uint32_t bofs = h * 2048;   // rows 128..255 of a [256][16] no-swizzle tile
                    for (int g = 0; g < G; g++) {
                        mbar_wait(RQ(0, h), ph, p.guard, 5);
                        ph ^= 1;
                        tc_fence_after();
                        const int j = g >> 1, half = g & 1;
                        const uint32_t a0 = ring + (j & (kRingSlots - 1)) * kChunkBytes;
                        const uint32_t b0 = sbase + kOffW + half * 9 * kB1Tile + bofs;
                        if (leader) {
                            mma_ss(dcol, smem_desc(sbase + kOffOnes, 2048, 128), smem_desc(sbase + kOffBias1 + bofs, 4096, 128), idesc_f16(128), 0);
#pragma unroll
                            for (int i = 0; i < 9; i++)
                                mma_ss(dcol, smem_desc(a0 + i * 16, kChunkBytes, 128), smem_desc(b0 + i * kB1Tile, 4096, 128), idesc_f16(128), 1);
                            mma_commit(MB(0, h));
                        }
                        __syncwarp();
                    }
                } else {
                    // ---- conv2 and conv3 of one output column d ----
                    const int d = irole - 2, h = d >> 1, dl = d & 1;
                    const uint32_t hb = tm + h * 128;
                    for (int g = 0; g < G; g++) {
                        // conv2: D2[d] = b2 + A1[d] (TMEM) x W2, K = 64 in 4 steps
                        mbar_wait(RQ(1, h), ph, p.guard, 6);
                        tc_fence_after();
                        if (leader) {
                            mma_ss(hb + 64 + dl * 32, smem_desc(sbase + kOffOnes, 2048, 128), smem_desc(sbase + kOffBias2, 512, 128), idesc_f16(32), 0);
#pragma unroll
                            for (int ks = 0; ks < 4; ks++)
                                mma_ts(hb + 64 + dl * 32, hb + dl * 32 + ks * 8, smem_desc(sbase + kImgB2 + ks * 1024, 512, 128), idesc_f16(32), 1);
                            mma_commit(MB(1, h));
                        }
                        __syncwarp();
                        // conv3 tap GEMM: T[d][tap] = A2[d] (TMEM) x W3, K = 32 in 2 steps
                        mbar_wait(RQ(2, h), ph, p.guard, 7);
                        ph ^= 1;
                        tc_fence_after();
                        if (leader) {
#pragma unroll
                            for (int ks = 0; ks < 2; ks++)
                                mma_ts(hb + dl * 32, hb + 64 + dl * 16 + ks * 8, smem_desc(sbase + kImgB3 + ks * 1024, 512, 128), idesc_f16(32), ks > 0);
                            mma_commit(MB(2, h));
                        }
                        __syncwarp();
                    }
                }
            }
        }
    } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 176;");
    const uint32_t tml = tm + ((uint32_t)(quarter * 32) << 16);         // + this warp's lane quarter
    uint8_t* ring_p = smem + kOffRing + pipe * kRingBytes;
    float* hx = (float*)(smem + kOffHx + pipe * kHxBytes);
    const int pbar = 1 + pipe;                    // named barrier of the pipeline (128 threads)
    mbar_wait(wbar, 0, p.guard, 1);
    uint32_t ph = 0;   // every completion barrier fires exactly once per group
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

        // staged Y rows of this thread: tile row tp, and tile row 128+tp for the first 8 threads
        const uint8_t* yrow0;
        const uint8_t* yrow1;
        {
            int r = min(max(R0 - 4 + tp, 0), H - 1);
            r = min(max(r - p.row0, 0), p.rows - 1);
            yrow0 = p.y + (size_t)r * p.pitch;
            r = min(max(R0 - 4 + 128 + (tp & 7), 0), H - 1);
            r = min(max(r - p.row0, 0), p.rows - 1);
            yrow1 = p.y + (size_t)r * p.pitch;
        }
        // 8 pixels of chunk q of one row -> two packed words (global loads; conversion happens at store time)
        auto fetch_row = [&](const uint8_t* yrow, int c0, uint32_t& lo, uint32_t& hi) {
            if (c0 >= 0 && c0 + 7 < W && p.y_aligned2) {
                const unsigned short* q16 = reinterpret_cast<const unsigned short*>(yrow + c0);
                lo = (uint32_t)q16[0] | ((uint32_t)q16[1] << 16);
                hi = (uint32_t)q16[2] | ((uint32_t)q16[3] << 16);
            } else {
                lo = hi = 0;
#pragma unroll
                for (int x = 0; x < 4; x++) {
                    lo |= (uint32_t)yrow[min(max(c0 + x, 0), W - 1)] << (8 * x);
                    hi |= (uint32_t)yrow[min(max(c0 + 4 + x, 0), W - 1)] << (8 * x);
                }
            }
        };
        // bytes -> FP16 via the 0x6400 (=1024.0) exponent trick, exact for 0..255; 16 B store into the ring
        auto store_row = [&](int q, int tr, uint32_t lo, uint32_t hi) {
            const int slot = q & (kRingSlots - 1);
            const __half2 k1024 = __half2half2(__ushort_as_half((unsigned short)0x6400));
            uint32_t h[4];
            const uint32_t w[4] = {__byte_perm(lo, 0x64646464u, 0x4140), __byte_perm(lo, 0x64646464u, 0x4342),
                                   __byte_perm(hi, 0x64646464u, 0x4140), __byte_perm(hi, 0x64646464u, 0x4342)};
#pragma unroll
            for (int x = 0; x < 4; x++) {
                const __half2 v = __hsub2(*reinterpret_cast<const __half2*>(&w[x]), k1024);
                h[x] = *reinterpret_cast<const uint32_t*>(&v);
            }
            const uint4 v4 = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(ring_p + slot * kChunkBytes + tr * 16) = v4;
            if (slot == 0) *reinterpret_cast<uint4*>(ring_p + kRingSlots * kChunkBytes + tr * 16) = v4;
        };
        auto stage_chunk_now = [&](int q) {
            const int c0 = s - 6 + 8 * q;
            uint32_t lo, hi;
            fetch_row(yrow0, c0, lo, hi);
            store_row(q, tp, lo, hi);
            if (tp < 8) {
                fetch_row(yrow1, c0, lo, hi);
                store_row(q, 128 + tp, lo, hi);
            }
        };
        // E1 of one half: D1 (2 columns x 64 ch fp32 = 128 TMEM columns) -> ReLU, FP16 -> A1 (64 columns, in place).
        // Four 32-column chunks, software-pipelined: chunk k+1 is in flight while chunk k is packed.
        auto e1_half = [&](uint32_t hb) {
            uint32_t va[32], vb[32], r[16];
            tmem_ld32(hb, va);
            tc_wait_ld();
            tmem_ld32(hb + 32, vb);
#pragma unroll
            for (int c = 0; c < 16; c++) r[c] = relu_pack_f16x2(__uint_as_float(va[2 * c]), __uint_as_float(va[2 * c + 1]));
            tmem_st16(hb, r);
            tc_wait_ld();
            tmem_ld32(hb + 64, va);
#pragma unroll
            for (int c = 0; c < 16; c++) r[c] = relu_pack_f16x2(__uint_as_float(vb[2 * c]), __uint_as_float(vb[2 * c + 1]));
            tmem_st16(hb + 16, r);
            tc_wait_ld();
            tmem_ld32(hb + 96, vb);
#pragma unroll
            for (int c = 0; c < 16; c++) r[c] = relu_pack_f16x2(__uint_as_float(va[2 * c]), __uint_as_float(va[2 * c + 1]));
            tmem_st16(hb + 32, r);
            tc_wait_ld();
#pragma unroll
            for (int c = 0; c < 16; c++) r[c] = relu_pack_f16x2(__uint_as_float(vb[2 * c]), __uint_as_float(vb[2 * c + 1]));
            tmem_st16(hb + 48, r);
        };
        // E2 of one half: D2 (2 columns x 32 ch at [hb+64, hb+128)) -> ReLU, FP16 -> A2 (in place at [hb+64, hb+96))
        auto e2_half = [&](uint32_t hb) {
            uint32_t va[32], vb[32], r[16];
            tmem_ld32(hb + 64, va);
            tc_wait_ld();
            tmem_ld32(hb + 96, vb);
#pragma unroll
            for (int c = 0; c < 16; c++) r[c] = relu_pack_f16x2(__uint_as_float(va[2 * c]), __uint_as_float(va[2 * c + 1]));
            tmem_st16(hb + 64, r);
            tc_wait_ld();
#pragma unroll
            for (int c = 0; c < 16; c++) r[c] = relu_pack_f16x2(__uint_as_float(vb[2 * c]), __uint_as_float(vb[2 * c + 1]));
            tmem_st16(hb + 80, r);
        };

        stage_chunk_now(0);
        stage_chunk_now(1);
        fence_proxy_async();
        mbar_arrive(RQ(0, 0));   // Y chunks 0,1 staged (and the previous segment is fully drained): conv1(0) may issue
        mbar_arrive(RQ(0, 1));

        // horizontal-tap accumulators: window column cw <-> image column t0-2+cw; columns 0..3 finish in the
        // current group, 4..7 are partial and slide down afterwards.  [.][m]: vertical tap
        float acc[8][5];
#pragma unroll
        for (int a = 0; a < 8; a++)
#pragma unroll
            for (int m = 0; m < 5; m++) acc[a][m] = 0.f;

        for (int g = 0; g < G; g++) {
            const int j = g >> 1;
            const bool prefetch = ((g & 1) == 0) && (j + 2 <= jlast + 1);
            uint32_t pre0 = 0, pre1 = 0, pre2 = 0, pre3 = 0;
            TL(0);
            if (prefetch) {   // global-load latency hides behind conv1 + E1
                fetch_row(yrow0, s - 6 + 8 * (j + 2), pre0, pre1);
                if (tp < 8) fetch_row(yrow1, s - 6 + 8 * (j + 2), pre2, pre3);
            }
            const int t0 = s - 2 + 4 * g;  // first T column of this group
            // ---------------- E1, half a then half b ----------------
            mbar_wait(MB(0, 0), ph, p.guard, 2);
            TL(1);
            tc_fence_after();
            e1_half(tml);
            if (prefetch) {
                store_row(j + 2, tp, pre0, pre1);
                if (tp < 8) store_row(j + 2, 128 + tp, pre2, pre3);
                fence_proxy_async();
            }
            tc_wait_st();
            tc_fence_before();
            mbar_arrive(RQ(1, 0));   // A1 of half a complete for my lane -> its two conv2 issuers start when all 128 arrived
            TL(2);
            mbar_wait(MB(0, 1), ph, p.guard, 2);
            TL(3);
            tc_fence_after();
            e1_half(tml + 128);
            tc_wait_st();
            tc_fence_before();
            mbar_arrive(RQ(1, 1));
            TL(4);
            // ---------------- E2, half a then half b ----------------
            mbar_wait(MB(1, 0), ph, p.guard, 3);
            TL(5);
            tc_fence_after();
            e2_half(tml);
            tc_wait_st();
            tc_fence_before();
            mbar_arrive(RQ(2, 0));
            TL(6);
            mbar_wait(MB(1, 1), ph, p.guard, 3);
            TL(7);
            tc_fence_after();
            e2_half(tml + 128);
            tc_wait_st();
            tc_fence_before();
            mbar_arrive(RQ(2, 1));
            TL(8);
            // ---------------- E3a: horizontal taps in registers, half a (T columns 0,1) then half b (2,3) ----------------
            mbar_wait(MB(2, 0), ph, p.guard, 4);
            TL(9);
            tc_fence_after();
            {
                uint32_t ta[32], tb[32];
                tmem_ld32(tml, ta);
                tc_wait_ld();
                tmem_ld32(tml + 32, tb);
                if (t0 >= 0 && t0 < W) taps_col<0>(acc, ta, t0 == 0, t0 == W - 1);
                tc_wait_ld();
                if (t0 + 1 >= 0 && t0 + 1 < W) taps_col<1>(acc, tb, t0 + 1 == 0, t0 + 1 == W - 1);
            }
            tc_fence_before();
            if (g + 1 < G) mbar_arrive(RQ(0, 0));   // T of half a fully read: its D1 region is reusable, next Y chunk is staged
            TL(10);
            mbar_wait(MB(2, 1), ph, p.guard, 4);
            TL(11);
            tc_fence_after();
            {
                uint32_t ta[32], tb[32];
                tmem_ld32(tml + 128, ta);
                tc_wait_ld();
                tmem_ld32(tml + 128 + 32, tb);
                if (t0 + 2 >= 0 && t0 + 2 < W) taps_col<2>(acc, ta, t0 + 2 == 0, t0 + 2 == W - 1);
                tc_wait_ld();
                if (t0 + 3 >= 0 && t0 + 3 < W) taps_col<3>(acc, tb, t0 + 3 == 0, t0 + 3 == W - 1);
            }
            tc_fence_before();
            if (g + 1 < G) mbar_arrive(RQ(0, 1));
            ph ^= 1;
            // vertical taps cross lanes: publish the 4 finished columns
#pragma unroll
            for (int m = 0; m < 5; m++)
#pragma unroll
                for (int a = 0; a < 4; a++) hx[(m * 4 + a) * 128 + tp] = acc[a][m];
#pragma unroll
            for (int a = 0; a < 4; a++)   // slide the window: the 4 partial columns become the next group's first 4
#pragma unroll
                for (int m = 0; m < 5; m++) { acc[a][m] = acc[a + 4][m]; acc[a + 4][m] = 0.f; }
            TL(12);
            named_bar(pbar, 128);              // hx visible to the whole pipeline
            TL(13);
            // ---------------- E3b: vertical taps, bias, truncate, clamp, store (4 columns per thread) ----------------
            {
                const int row = R0 + tp;
                int ln[5];
#pragma unroll
                for (int m = 0; m < 5; m++) ln[m] = min(max(min(max(row + m - 2, 0), H - 1) - R0, 0), 127);  // src/srcnn.cpp:203
                const int c0 = t0 - 2;
                const bool row_ok = (tp >= 2) && (tp <= 125) && (row >= band_begin) && (row < band_end);
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
            TL(14);
        }
        // segment done: every MMA of this pipeline has been waited for; the hx reads of the last group
        // finish before the next segment's first hx write (conv3 of the next group needs all 128 arrivals)
    }
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
    // bias tiles: FP32 bias = hi + lo in FP16, multiplied by the two 1.0 columns of the "ones" A tile
    auto hi_lo = [](float b, float& hi, float& lo) {
        hi = __half2float(__float2half_rn(b));
        lo = b - hi;
    };
    for (int d = 0; d < 4; d++)
        for (int ch = 0; ch < 64; ch++) {
            float hi, lo;
            hi_lo(P[kOffB1 + ch], hi, lo);
            const size_t n = (size_t)d * 64 + ch;
            put_h(img.data(), kOffBias1 + n * 16 + 0, hi);
            put_h(img.data(), kOffBias1 + n * 16 + 2, lo);
        }
    for (int n = 0; n < 32; n++) {
        float hi, lo;
        hi_lo(P[kOffB2 + n], hi, lo);
        put_h(img.data(), kOffBias2 + (size_t)n * 16 + 0, hi);
        put_h(img.data(), kOffBias2 + (size_t)n * 16 + 2, lo);
    }
    for (int r = 0; r < 128; r++) {  // ones tile: K chunk 0 = [1,1,0,0,0,0,0,0] per row, K chunk 1 = 0
        put_h(img.data(), kOffOnes + (size_t)r * 16 + 0, 1.0f);
        put_h(img.data(), kOffOnes + (size_t)r * 16 + 2, 1.0f);
    }
    SRCNN_CUDA(c, cudaMalloc(&c->d_tc_weights, kWeightBytes));
    c->tc_weights_bytes = kWeightBytes;
    SRCNN_CUDA(c, cudaMemcpy(c->d_tc_weights, img.data(), kWeightBytes, cudaMemcpyHostToDevice));
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
    p.y_aligned2 = ((((uintptr_t)a.y) | a.pitch) & 1) == 0;
    p.wimg = (const uint8_t*)c->d_tc_weights;
    p.gpb = (a.W + 3) / 4;
    const int nbands = (a.out_end - a.out_begin + kBandRows - 1) / kBandRows;
    p.total_groups = (long long)nbands * p.gpb;
    p.guard = c->d_guard;
    p.dbg = nullptr;
    if (getenv("SRCNN_TC_DEBUG")) {
        if (!c->work_buf.p) {
            int rc = ensure(c, c->work_buf, 2 * 24 * 16 * sizeof(long long));
            if (rc) return rc;
        }
        cudaMemsetAsync(c->work_buf.p, 0, 2 * 24 * 16 * sizeof(long long), c->stream);
        p.dbg = (long long*)c->work_buf.p;
    }
    // one persistent CTA per SM; fewer when the image is too small to give every warpgroup ~8 groups
    long long want = (p.total_groups + 15) / 16;
    int grid = (int)std::min<long long>(c->sm_count, std::max<long long>(1, want));
    k_srcnn_tc<<<grid, kThreads, kSmemBytes, c->stream>>>(p);
    c->launches++;
    SRCNN_CUDA(c, cudaGetLastError());
    return SRCNN_OK;
}

}  // namespace srcnn

// debug hook: copies the last launch's timeline (2 pipelines x 24 groups x 8 stamps) to the host
extern "C" __attribute__((visibility("default"))) int srcnn_debug_tc_timeline(srcnn_ctx* c, long long* out) {
    if (!c || !c->work_buf.p) return SRCNN_E_ARG;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    return cudaMemcpy(out, c->work_buf.p, 2 * 24 * 16 * sizeof(long long), cudaMemcpyDeviceToHost) == cudaSuccess ? SRCNN_OK : SRCNN_E_CUDA;
}

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
