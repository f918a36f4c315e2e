// color_bicubic_int.cu -- kernel A of the hot path for INTEGER up-scales (x2, x4): the shapes every BASELINE configuration but the
// 384x384 golden pair uses.  Same arithmetic, bit for bit, as the generic kernel in color_bicubic.cu -- cvtColor(BGR2YCrCb)
// src/srcnn.cpp:509, split :540, three resize(..., CV_INTER_CUBIC) :570-583 -- restated for the case where the tap tables are
// periodic: output sample d = S*j + k reads source columns j-2+(k >= S/2) .. +3 with the taps of phase k, whatever j is.  So there
// are S constant tap sets (kernel parameters: constant-bank operands, no table loads), and a window of four source rows
// feeds S output rows.
//
// A tile = 256 output columns x isr*S output rows (isr = 8 or 16 source rows, chosen per launch); one CTA of three warps per
// tile, warp w owns plane w (Y, Cr, Cb), a lane owns 8 adjacent columns.
//   (1) the tile's source footprint (isr+3 rows) is colour-converted into shared memory once: four pixels per step from three
//       aligned 32-bit loads (all of a thread's loads are issued before its first conversion), dp2a for the dot products, one
//       32-bit shared store per plane; replicate border applied here;
//   (2) each lane walks down the footprint rows: per row 8 horizontal sums (16 dp2a on funnel-shifted windows; the int -> float
//       conversion is the dp2a's accumulator starting at the bit pattern of 1.5 * 2^23 and ONE packed subtraction per pair), kept
//       as packed FP32 pairs in a register window of four rows; per row S output rows of
//       v = H0*b0 + (H1*b1 + (H2*b2 + H3*b3)), every product and sum rounded on its own (f2_mul_rn / f2_add_rn), round-half-even,
//       saturate, one 8- or 16-byte store per lane and row.  No shared-memory round trip of the sums, no per-sample addressing.
// Y goes to the padded FP16 plane of the tcgen05 kernel when the context has one (color_bicubic.cu explains the layout).
// Measured and not kept (tools/experiments/color_bicubic_int_persistent_tma.cu.txt): fetching a tile's raw rows by TMA
// (neutral), persistent CTAs with the next tile's rows prefetched by TMA -- as one block-wide pipeline, with a producer warp, or
// as three per-warp pipelines (all slower: the kernel is bound by instruction issue while a wave of tiles computes, and every
// persistent form added instructions or took warps away from the walk).
#include "color_bicubic.h"

namespace srcnn {

namespace {

// x2 instance: minimum CTAs per SM the register allocation is held to.  Same-box A/B inside the bench step (tools/ab_bench_libs.sh):
// 9 (72 registers, 96 B of spills in the walk) 49.1 GPix/s, 8 (80 registers, 32 B) 49.5-49.8, 6 (91 registers, none) 49.9 -- but
// 1 % slower than 8 in many-wave launches (1024 720p frames: 8.33 vs 8.24 ms); 10 (64 registers) 48.5.
#ifndef SRCNN_KA_MINB
#define SRCNN_KA_MINB 8
#endif
#ifndef SRCNN_KA_MINB4
#define SRCNN_KA_MINB4 5         // x4 instance: 126 registers, no spills -- 168.7 vs 174.6 us per 4K -> 16K frame with 6 (96 registers, 64 B of spills)
#endif
constexpr int kISRMax = 16;         // most source rows a tile advances by (IntTaps::isr; plus 3 rows of apron staged with them)
constexpr int kIRows = kISRMax + 3;
constexpr int kIPitch = 144;       // bytes per converted row: 34 four-pixel groups + one spare word for the 3-word window read
constexpr int kITW = 256;          // output columns per tile
constexpr float kMagic = 12582912.0f;   // 1.5 * 2^23: int bits 0x4B400000 + k is the float 12582912 + k for |k| < 2^22

struct IntTaps {
    uint32_t hk01[4], hk23[4];          // horizontal taps of phase k as signed 16-bit pairs (taps 0,1 / 2,3)
    unsigned long long vb[4][4];        // vertical taps of phase k as FP32 pairs (b, b), b = tap * 2^-22 (exact)
    uint32_t sel[4];                    // byte selectors that turn three source words into four (B, G, R, x) pixel words
    int a_first, i_last;                // first / last source row index i whose output rows S*i + S/2 .. + S-1 are wanted
    int vec_ok;                         // source rows are 4-byte aligned: footprint groups inside the image take 32-bit loads
    int isr;                            // source rows a tile advances by (4 .. kISRMax)
    int ntx, nty, ntiles;               // tiles per row of tiles, rows of tiles per frame, tiles in the launch (frames included)
};


__device__ __forceinline__ int dp2a_lo_su(uint32_t k, uint32_t px, int c) {
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(k), "r"(px), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_hi_su(uint32_t k, uint32_t px, int c) {
    int d;
    asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(k), "r"(px), "r"(c));
    return d;
}
// sat_u8(lo) | sat_u8(hi) << 8 | (upper & 0xFFFF) << 16
__device__ __forceinline__ uint32_t pack2_sat_u8(int lo, int hi, uint32_t upper) {
    uint32_t d;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(hi), "r"(lo), "r"(upper));
    return d;
}

// which 4-byte window (in bytes past the lane's first footprint column) output column c of a lane's 8 reads
template <int S>
__host__ __device__ constexpr int win_ofs(int c) {
    // floor((2c + 1 - S) / (2S)) + 1, written for non-negative operands
    return (2 * c + 1 - S + 2 * S) / (2 * S);
}

template <int S>
struct Geo {
    static constexpr int G = kITW / S / 4 + 2;             // four-pixel groups per footprint row (the first starts 4 columns left of the tile)
};

// (1) twelve source bytes = four pixels -> one 32-bit word per plane of the converted footprint.  OpenCV RGB2YCrCb_i<uchar>, yuv_shift 14.
__device__ __forceinline__ void convert_group(const IntTaps& t, uint32_t w0, uint32_t w1, uint32_t w2, uint8_t* y4, uint8_t* cr4, uint8_t* cb4) {
    const uint32_t kYlo = 1868u | (9617u << 16), kYhi = 4899u, kR = 11682u, kB = 9241u;
    const int cadd = (128 << 14) + 8192;
    uint32_t pw[4];     // (B, G, R, x) per pixel
    pw[0] = __byte_perm(w0, w0, t.sel[0]);
    pw[1] = __byte_perm(w0, w1, t.sel[1]);
    pw[2] = __byte_perm(w1, w2, t.sel[2]);
    pw[3] = __byte_perm(w2, w2, t.sel[3]);
    int Y[4], Cr[4], Cb[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        Y[j] = dp2a_hi_su(kYhi, pw[j], dp2a_lo_su(kYlo, pw[j], 8192)) >> 14;
        Cr[j] = (dp2a_hi_su(kR, pw[j], cadd) - Y[j] * 11682) >> 14;
        Cb[j] = (dp2a_lo_su(kB, pw[j], cadd) - Y[j] * 9241) >> 14;
    }
    *reinterpret_cast<uint32_t*>(y4) = __byte_perm(__byte_perm(Y[0], Y[1], 0x0040), __byte_perm(Y[2], Y[3], 0x0040), 0x5410);
    *reinterpret_cast<uint32_t*>(cr4) = pack2_sat_u8(Cr[0], Cr[1], pack2_sat_u8(Cr[2], Cr[3], 0u));
    *reinterpret_cast<uint32_t*>(cb4) = pack2_sat_u8(Cb[0], Cb[1], pack2_sat_u8(Cb[2], Cb[3], 0u));
}
// the same twelve bytes for a group that crosses the left / right image edge: from clamped pixels of a row that starts at `row`
// (the byte of pixel 0, channel 0)
__device__ __forceinline__ void clamped_group(const uint8_t* row, int gx, int sw, uint32_t (&w)[3]) {
    w[0] = w[1] = w[2] = 0u;
#pragma unroll
    for (int j = 0; j < 12; j++) {
        const int x = clampi(gx + j / 3, 0, sw - 1);
        w[j >> 2] |= (uint32_t)row[3 * (long long)x + (j % 3)] << (8 * (j & 3));
    }
}

// (2) the walk of one warp (one plane) down a tile's converted footprint `plane` (rows kIPitch bytes apart): lane's output columns
// x0 .. x0+7 read footprint bytes cb .. cb+7 of a row
template <int S>
__device__ __forceinline__ void walk_tile(const ResizeDev& p, const IntTaps& t, const uint8_t* plane, int pl, int lane, int X0, int a,
                                          int nit, size_t fz) {
    const int x0 = X0 + 8 * lane;
    if (x0 >= p.ow) return;                        // ow is a multiple of 8 (launch condition)
    const int cb = (8 * lane) / S + 2;
    const uint8_t* const srow = plane + (cb & ~3);
    const int wsh = (cb & 3) * 8;                  // x2: always 16
    const unsigned long long nz = p.negzero2;
    const unsigned long long mm = f2_pack(-kMagic, -kMagic);
    const bool y16 = pl == 0 && p.y16 != nullptr;
    const size_t ostep = y16 ? p.pitch16 : p.pitch;

                auto hsum = [&](int r, unsigned long long (&h)[4]) {
                    const uint32_t* w = reinterpret_cast<const uint32_t*>(srow + r * kIPitch);
                    const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
                    uint32_t lo = 0, hi = 0;
                    if (S != 2) { lo = __funnelshift_r(w0, w1, wsh); hi = __funnelshift_r(w1, w2, wsh); }
                    float f[8];
    #pragma unroll
                    for (int c = 0; c < 8; c++) {
                        const int o = win_ofs<S>(c);
                        uint32_t px4;
                        if (S == 2) {                  // bytes cb + o .. + 3 straight from the three words (cb & 3 == 2)
                            const int sh = 16 + 8 * o;
                            px4 = sh < 32 ? __funnelshift_r(w0, w1, sh) : (sh == 32 ? w1 : __funnelshift_r(w1, w2, sh - 32));
                        } else {
                            px4 = o == 0 ? lo : (o == 4 ? hi : __funnelshift_r(lo, hi, 8 * o));
                        }
                        const int acc = dp2a_hi_su(t.hk23[c % S], px4, dp2a_lo_su(t.hk01[c % S], px4, 0x4B400000));
                        f[c] = __int_as_float(acc);    // = 12582912 + sum, exactly
                    }
    #pragma unroll
                    for (int q = 0; q < 4; q++) h[q] = f2_add_rn(f2_pack(f[2 * q], f[2 * q + 1]), mm);   // = (float)sum, exactly
                };

                // running output pointer: row S*a + S/2 first, one row further per emitted row (dereferenced only for wanted rows)
                int dy = S * a + S / 2;
                uint8_t* outp;
                if (y16) outp = p.y16 + fz * p.y16_frame + ((long long)dy - p.plane_row0) * (long long)p.pitch16 + 2 * (size_t)(x0 + kY16Pad);
                else outp = (pl == 0 ? p.y : (pl == 1 ? p.cr : p.cb)) + fz * p.plane_frame + ((long long)dy - p.plane_row0) * (long long)p.pitch + x0;
                uint8_t* const outp0 = outp;

                auto emit = [&](const unsigned long long (&h0)[4], const unsigned long long (&h1)[4], const unsigned long long (&h2)[4],
                                const unsigned long long (&h3)[4]) {
    #pragma unroll
                    for (int m = 0; m < S; m++) {
                        const int ph = (S / 2 + m) % S;
                        const bool wanted = dy >= p.row_begin && dy < p.row_end;     // uniform; only the first / last row of tiles has unwanted rows
                        uint32_t r[8];
    #pragma unroll
                        for (int q = 0; q < 4; q++) {
                            unsigned long long v = f2_mul_rn(h3[q], t.vb[ph][3], nz);
                            v = f2_add_rn(f2_mul_rn(h2[q], t.vb[ph][2], nz), v);
                            v = f2_add_rn(f2_mul_rn(h1[q], t.vb[ph][1], nz), v);
                            v = f2_add_rn(f2_mul_rn(h0[q], t.vb[ph][0], nz), v);
                            float va, vb;
                            f2_unpack(v, va, vb);
                            r[2 * q] = sat_u8_rn(va);
                            r[2 * q + 1] = sat_u8_rn(vb);
                        }
                        if (y16) {   // exact u8 -> FP16: 0x6400 | v is 1024 + v, minus 1024
                            const __half2 k1024 = __float2half2_rn(1024.f);
                            uint32_t hw[4];
    #pragma unroll
                            for (int q = 0; q < 4; q++) {
                                uint32_t u = __byte_perm(r[2 * q], r[2 * q + 1], 0x5410) | 0x64006400u;   // bytes (r0, 0x64, r1, 0x64)
                                __half2 hh = __hsub2(*reinterpret_cast<__half2*>(&u), k1024);
                                hw[q] = *reinterpret_cast<uint32_t*>(&hh);
                            }
                            if (wanted) *reinterpret_cast<uint4*>(outp) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                        } else {
                            const uint32_t lo4 = __byte_perm(__byte_perm(r[0], r[1], 0x0040), __byte_perm(r[2], r[3], 0x0040), 0x5410);
                            const uint32_t hi4 = __byte_perm(__byte_perm(r[4], r[5], 0x0040), __byte_perm(r[6], r[7], 0x0040), 0x5410);
                            if (wanted) *reinterpret_cast<uint2*>(outp) = make_uint2(lo4, hi4);
                        }
                        dy++;
                        outp += ostep;
                    }
                };

                unsigned long long hA[4], hB[4], hC[4], hD[4];
                hsum(0, hA);
                hsum(1, hB);
                hsum(2, hC);
                for (int tt = 0; tt < nit; tt += 4) {  // window rotation by unrolling, not by register moves
                    hsum(tt + 3, hD);
                    emit(hA, hB, hC, hD);
                    if (tt + 1 >= nit) break;
                    hsum(tt + 4, hA);
                    emit(hB, hC, hD, hA);
                    if (tt + 2 >= nit) break;
                    hsum(tt + 5, hB);
                    emit(hC, hD, hA, hB);
                    if (tt + 3 >= nit) break;
                    hsum(tt + 6, hC);
                    emit(hD, hA, hB, hC);
                }

                // FP16 plane: the replicated columns -8..-1 and W..W+7 (conv1 reads Y[clamp(c-4)], src/srcnn.cpp:279) -- by the two
                // lanes that own the first / last eight columns of the image, from what they have just stored
                if (y16 && (x0 == 0 || x0 + 8 == p.ow)) {
                    int d = S * a + S / 2;
                    uint8_t* o = outp0;
                    for (int j = 0; j < nit * S; j++, d++, o += ostep) {
                        if (d < p.row_begin || d >= p.row_end) continue;
                        const uint4 v = *reinterpret_cast<const uint4*>(o);
                        if (x0 == 0) {
                            const uint32_t e = __byte_perm(v.x, 0, 0x1010);
                            *reinterpret_cast<uint4*>(o - 2 * kY16Pad) = make_uint4(e, e, e, e);
                        }
                        if (x0 + 8 == p.ow) {
                            const uint32_t e = __byte_perm(v.w, 0, 0x3232);
                            *reinterpret_cast<uint4*>(o + 16) = make_uint4(e, e, e, e);
                        }
                    }
                }
}

// ------------------------------------------------------------------------------------------------
// The kernel: one CTA of 3 warps per tile.
// ------------------------------------------------------------------------------------------------
template <int S>
__global__ void __launch_bounds__(96, S == 2 ? SRCNN_KA_MINB : SRCNN_KA_MINB4) k_color_bicubic_int(const ResizeDev p, const IntTaps t) {
    constexpr int G = Geo<S>::G;
    __shared__ __align__(16) uint8_t sP[3][kIRows][kIPitch];
    const int tid = threadIdx.x;
    const int tile = blockIdx.x;
    const int bx = tile % t.ntx, rest = tile / t.ntx;
    const int by = rest % t.nty;
    const size_t fz = (size_t)(rest / t.nty);
    const int X0 = bx * kITW;
    const int a = t.a_first + by * t.isr;          // footprint row r is source row a - 1 + r
    const int F0 = X0 / S - 4;                     // first footprint column (a multiple of 4: aligned 12-byte groups)
    const int nit = min(t.isr, t.i_last - a + 1);
    const int nsr = nit + 3;
    const uint8_t* const fsrc = p.src + fz * p.src_frame;

    // (1) colour-convert the footprint; replicate border applied here
    if (t.vec_ok && F0 >= 0 && F0 + 4 * G <= p.sw) {
        // every group lies inside the image and is word-aligned: all of a thread's loads go out before its first conversion
        constexpr int NG = (kIRows * G + 95) / 96;
        uint32_t raw[NG][3];
#pragma unroll
        for (int it = 0; it < NG; it++) {
            const int g = tid + 96 * it;
            if (g < nsr * G) {
                const int r = g / G, q = g - r * G;
                const int gy = clampi(a - 1 + r, 0, p.sh - 1) - p.src_row0;
                const uint32_t* w = reinterpret_cast<const uint32_t*>(fsrc + (size_t)gy * p.src_stride + 3 * (size_t)(F0 + 4 * q));
                raw[it][0] = w[0]; raw[it][1] = w[1]; raw[it][2] = w[2];
            }
        }
#pragma unroll
        for (int it = 0; it < NG; it++) {
            const int g = tid + 96 * it;
            if (g < nsr * G) {
                const int r = g / G, q = g - r * G;
                convert_group(t, raw[it][0], raw[it][1], raw[it][2], &sP[0][r][4 * q], &sP[1][r][4 * q], &sP[2][r][4 * q]);
            }
        }
    } else {
        for (int g = tid; g < nsr * G; g += 96) {
            const int r = g / G, q = g - r * G;
            const int gy = clampi(a - 1 + r, 0, p.sh - 1) - p.src_row0;
            uint32_t w[3];
            clamped_group(fsrc + (size_t)gy * p.src_stride, F0 + 4 * q, p.sw, w);
            convert_group(t, w[0], w[1], w[2], &sP[0][r][4 * q], &sP[1][r][4 * q], &sP[2][r][4 * q]);
        }
    }
    __syncthreads();
    // (2) the walk
    walk_tile<S>(p, t, &sP[tid >> 5][0][0], tid >> 5, tid & 31, X0, a, nit, fz);
    // Launched with programmatic stream serialisation (api.cu, may_start_early) this grid may have started while the merge kernel
    // of the previous whole-path call was still running; nothing here depends on it.  The last tile waits for that kernel at its
    // very end, so that "this grid has finished" still implies "everything before it in the stream has finished" for whatever
    // follows.  A no-op after an ordinary launch.
    if (tile == t.ntiles - 1) asm volatile("griddepcontrol.wait;" ::: "memory");
}

inline int floordiv(int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }

template <int S>
int launch_variant(Ctx* c, const ResizeDev& p, IntTaps& t, int nframes, bool early) {
    static int ctas_per_sm[64] = {0};              // per device; a benign race: every writer stores the same value
    int& per_sm = ctas_per_sm[c->device & 63];
    if (per_sm == 0) {
        int n = 0;
        SRCNN_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_color_bicubic_int<S>, 96, 0));
        per_sm = std::max(1, n);
    }
    // Tile height (same-box A/B, B200): a single 4K frame runs fastest as ~1.5 waves of 8-row tiles (23.6 us; 6 rows: 23.7,
    // 4 rows: 25.1, 12 rows: 25.5, 16 rows = one wave: 27.6 -- the CTAs of one wave wait for their source rows together, and the
    // last wave's tiles run at half occupancy for a whole tile life); launches of many waves (batches, 16K frames, gigapixel
    // bands) take 16-row tiles for the smaller apron (7.5 vs 7.8 ms for 1024 720p frames).
    const int resident = per_sm * std::max(1, c->sm_count);
    const int nrows = t.i_last - t.a_first + 1;
    auto tiles_for = [&](int isr) { return (long long)t.ntx * ((nrows + isr - 1) / isr) * nframes; };
    t.isr = tiles_for(kISRMax) > 2LL * resident ? kISRMax : 8;
    if (c->ka_int_isr >= 4 && c->ka_int_isr <= kISRMax) t.isr = c->ka_int_isr;
    const long long ntiles = tiles_for(t.isr);
    if (ntiles > 0x3fffffffLL) return SRCNN_E_ARG;
    t.nty = (nrows + t.isr - 1) / t.isr;
    t.ntiles = (int)ntiles;
    if (early) {   // programmatic dependent launch: may begin before the stream's previous kernel (a merge of ours) has finished
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)t.ntiles);
        cfg.blockDim = dim3(96);
        cfg.dynamicSmemBytes = 0;
        cfg.stream = c->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        SRCNN_CUDA(c, cudaLaunchKernelEx(&cfg, k_color_bicubic_int<S>, p, t));
    } else {
        k_color_bicubic_int<S><<<t.ntiles, 96, 0, c->stream>>>(p, t);
    }
    c->launches++;
    SRCNN_CUDA(c, cudaGetLastError());
    return SRCNN_OK;
}

}  // namespace

int launch_color_bicubic_int(Ctx* c, const ResizeDev& p, const ResizeArgs& a, bool* done) {
    *done = false;
    const int S = a.tx->int_scale;
    if (S == 0 || a.ty->int_scale != S || (a.ow & 7) || a.row_end <= a.row_begin) return SRCNN_OK;
    // stores: 8 bytes per lane and row into the u8 planes, 16 into the FP16 plane
    if ((((uintptr_t)p.y | (uintptr_t)p.cr | (uintptr_t)p.cb | p.pitch | p.plane_frame) & 7) != 0) return SRCNN_OK;
    if (p.y16 && (((uintptr_t)p.y16 | p.pitch16 | p.y16_frame) & 15) != 0) return SRCNN_OK;

    IntTaps t;
    const float sc = 1.0f / 4194304.0f;   // 2^-22, exact
    for (int k = 0; k < 4; k++) {
        const short4 cx = a.tx->h_coef[k % S], cy = a.ty->h_coef[k % S];
        t.hk01[k] = (uint32_t)(unsigned short)cx.x | ((uint32_t)(unsigned short)cx.y << 16);
        t.hk23[k] = (uint32_t)(unsigned short)cx.z | ((uint32_t)(unsigned short)cx.w << 16);
        const float b[4] = {(float)cy.x * sc, (float)cy.y * sc, (float)cy.z * sc, (float)cy.w * sc};
        for (int j = 0; j < 4; j++) {
            uint32_t u;
            memcpy(&u, &b[j], 4);
            t.vb[k][j] = (unsigned long long)u | ((unsigned long long)u << 32);
        }
    }
    // three words = bytes c0 c1 c2 of pixels 0..3; pixel word = (B, G, R, x): B is c0 for BGR input, c2 for RGB input
    if (p.swapRB) { t.sel[0] = 0x3012; t.sel[1] = 0x6345; t.sel[2] = 0x5234; t.sel[3] = 0x0123; }
    else          { t.sel[0] = 0x3210; t.sel[1] = 0x6543; t.sel[2] = 0x5432; t.sel[3] = 0x0321; }
    t.a_first = floordiv(a.row_begin - S / 2, S);
    t.i_last = floordiv(a.row_end - 1 - S / 2, S);
    t.vec_ok = (((uintptr_t)p.src | p.src_stride | p.src_frame) & 3) == 0;
    const int nframes = std::max(1, a.nframes);
    t.ntx = (a.ow + kITW - 1) / kITW;
    if ((long long)t.ntx * (t.i_last - t.a_first + 1) * nframes > 0x3fffffffLL) return SRCNN_OK;
    const int rc = S == 2 ? launch_variant<2>(c, p, t, nframes, a.early) : launch_variant<4>(c, p, t, nframes, a.early);
    if (rc) return rc;
    *done = true;
    return SRCNN_OK;
}

}  // namespace srcnn

