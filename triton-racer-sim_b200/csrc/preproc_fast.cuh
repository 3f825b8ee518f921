// preproc_fast.cuh — frame-resident fast path of the fused observation kernel (sm_100a).
//
// Same contract as k_preprocess (preproc_kernel.cuh) for frames that (a) fit one CTA's shared memory whole,
// (b) have a width that is a multiple of 32 and (c) are 16-byte aligned.  Design:
//
//  * one CTA per frame, persistent over frames, two CTAs per SM; the frame arrives by a TMA bulk copy
//    (cp.async.bulk + mbarrier) and the next frame's copy is issued as soon as the pixels are dead, so it
//    overlaps non-maximum suppression, hysteresis and the output stores of the current frame
//  * a thread owns a 4-pixel-wide column strip and walks down SEG rows with a rolling 3-row window held in
//    registers; two adjacent lanes make one byte (8 pixels) of a bit plane with a single shuffle
//  * the Sobel arithmetic runs two pixels per instruction on the FMA pipe: a u8 value zero-extended to 16 bits
//    is a valid fp16 subnormal (n * 2^-24) and every intermediate stays below 2048, so HADD2/HFMA2 on the raw bit
//    patterns are exact integer add/sub/scale with free |x| and -x operand modifiers (results are sign-magnitude);
//    HSET2 on the same patterns gives packed compares for the range tests and the non-maximum suppression
//  * HSV: max/min/delta and the hue numerator are computed packed (VIMNMX3.U16x2, IADD3), the two fixed-point
//    multiplies per pixel stay scalar (IMAD) — bit-exact with OpenCV's integer path
//  * colour masks, NMS candidates and strong pixels live as bit planes; hysteresis is a word-parallel flood fill
//  * outputs are written once: u8 bytes expanded from plane nibbles by multiply-spread, f32 pixels fetched from an
//    8-entry {0,1}^3 table (masks are exactly 0.0/1.0 after /255)
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "pixel_math.cuh"
#include "preproc_kernel.cuh"

namespace trs {

struct FastGeom {
    int seg_rows, nseg, nsg, threads;
    int mag_stride;                  // u16 elements per magnitude row (w + 8; pixel x at index x + 4)
    int off_pix, off_mag, off_cand, off_edge, off_mask, off_tab, off_lut, off_f32lut, off_bar, off_red, total;
};

__host__ __device__ inline FastGeom fast_geometry(int h, int w, int n_ranges, int seg_rows)
{
    FastGeom g;
    g.seg_rows = seg_rows;
    g.nseg = (h + seg_rows - 1) / seg_rows;
    g.nsg = w / 32;
    g.threads = 32 * g.nsg * ((g.nseg + 3) / 4);
    g.mag_stride = w + 8;
    int o = 0;
    g.off_pix = o;    o += ((h * w * 3 + 15) & ~15) + 16;
    g.off_mag = o;    o += (((h + 2) * g.mag_stride * 2) + 15) & ~15;
    const int plane = (((h + 2) * g.nsg * 4) + 15) & ~15;          // one zero row above and below (hysteresis reads y-1 / y+1)
    g.off_cand = o;   o += plane;
    g.off_edge = o;   o += plane;
    g.off_mask = o;   o += plane * n_ranges;
    g.off_tab = o;    o += 1024 + 2048;                            // sdiv[256] int32, hue table[256] int2
    g.off_lut = o;    o += 256;
    g.off_f32lut = o; o += 128;
    g.off_bar = o;    o += 16;
    g.off_red = o;    o += 64;
    g.total = o;
    return g;
}

// packed-compare form of one colour range: fp16-subnormal bit patterns duplicated in both halves
struct FastRange {
    uint32_t lo[3], hi[3];
    uint32_t flags;                  // bit 2c: the lower bound of channel c can fail; bit 2c+1: the upper bound can fail
};

struct FastParams {
    PreKParams k;
    FastGeom g;
    FastRange fr[3];
    uint32_t low2, high2;            // edge thresholds as packed patterns
    int need_hue;                    // some range has a hue bound that can fail
};

// ---- small PTX helpers --------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) { return __byte_perm(a, b, sel); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// fp16x2 views of 32-bit registers (bit patterns are integers n < 2048 == fp16 subnormals n * 2^-24)
__device__ __forceinline__ __half2 h2(uint32_t x) { return *reinterpret_cast<__half2*>(&x); }
__device__ __forceinline__ uint32_t u32(__half2 x) { return *reinterpret_cast<uint32_t*>(&x); }
__device__ __forceinline__ uint32_t hsub(uint32_t a, uint32_t b) { return u32(__hsub2(h2(a), h2(b))); }
__device__ __forceinline__ uint32_t hadd(uint32_t a, uint32_t b) { return u32(__hadd2(h2(a), h2(b))); }
__device__ __forceinline__ uint32_t hx2p(uint32_t a, uint32_t b) { return u32(__hfma2(h2(a), __float2half2_rn(2.0f), h2(b))); }   // 2a + b
__device__ __forceinline__ uint32_t habsadd(uint32_t a, uint32_t b) { return u32(__hadd2(__habs2(h2(a)), __habs2(h2(b)))); }     // |a| + |b|
__device__ __forceinline__ uint32_t hmaxu(uint32_t a, uint32_t b) { return u32(__hmax2(h2(a), h2(b))); }
__device__ __forceinline__ uint32_t hgt_mask(uint32_t a, uint32_t b) { return __hgt2_mask(h2(a), h2(b)); }
__device__ __forceinline__ uint32_t hge_mask(uint32_t a, uint32_t b) { return __hge2_mask(h2(a), h2(b)); }
__device__ __forceinline__ uint32_t hle_mask(uint32_t a, uint32_t b) { return __hle2_mask(h2(a), h2(b)); }
__device__ __forceinline__ uint32_t heq_mask(uint32_t a, uint32_t b) { return __heq2_mask(h2(a), h2(b)); }
__device__ __forceinline__ uint32_t bsel(uint32_t m, uint32_t a, uint32_t b) { return (a & m) | (b & ~m); }                      // one LOP3

// direction class from |dx|, |dy| and the sign-difference flag (pixel_math.cuh: canny_dir)
__device__ __forceinline__ uint32_t dir_code(uint32_t ax, uint32_t ay, uint32_t sdiff)
{
    const int t22 = (int)(ax * 13573u);
    const int ay15 = (int)(ay << 15);
    const int t67 = t22 + (int)(ax << 16);
    uint32_t code = 2u + sdiff;
    code = (ay15 > t67) ? 1u : code;
    code = (ay15 < t22) ? 0u : code;
    return code;
}

// same classes as dir_code, computed from sign bits: h <=> ay*2^15 - ax*13573 < 0, v <=> ax*79109 - ay*2^15 < 0
__device__ __forceinline__ uint32_t dir_bits(uint32_t ax, uint32_t ay, uint32_t sdiff)
{
    const int t22 = (int)(ax * 13573u);
    const int nu = (int)(ay * 32768u) - t22;                 // < 0: horizontal
    const int wv = t22 + (int)(ax * 65536u) - (int)(ay * 32768u);   // < 0: vertical
    const uint32_t H = (uint32_t)(nu >> 31), V = (uint32_t)(wv >> 31);
    return ~H & ((V & 1u) | (~V & (2u | sdiff)));
}

// four 0xffff/0 half masks (pixels 0,2 in `a`; pixels 1,3 in `b`) -> nibble, bit q = pixel q
__device__ __forceinline__ uint32_t nibble_of(uint32_t a, uint32_t b)
{
    const uint32_t x = (a & 0x00040001u) | (b & 0x00080002u);
    return (x | (x >> 16)) & 0xfu;
}

enum { FAST_MAX_THREADS = 320 };

template <int NR, bool EDGE>
__global__ void __launch_bounds__(FAST_MAX_THREADS, 2) k_preprocess_fast(const __grid_constant__ FastParams P)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const PreKParams& p = P.k;
    const FastGeom& G = P.g;
    uint8_t* s_pix = smem + G.off_pix;
    uint16_t* s_mag = reinterpret_cast<uint16_t*>(smem + G.off_mag);
    const int h = p.h, w = p.w, ww = G.nsg;
    // bit planes: row y at word row y + 1; rows 0 and h + 1 stay zero
    uint32_t* s_cand = reinterpret_cast<uint32_t*>(smem + G.off_cand) + ww;
    uint32_t* s_edge = reinterpret_cast<uint32_t*>(smem + G.off_edge) + ww;
    uint32_t* s_mask = reinterpret_cast<uint32_t*>(smem + G.off_mask) + ww;
    const int plane_stride = (((h + 2) * ww * 4 + 15) & ~15) >> 2;        // words between consecutive mask planes
    int32_t* s_sdiv = reinterpret_cast<int32_t*>(smem + G.off_tab);
    int2* s_hue = reinterpret_cast<int2*>(smem + G.off_tab + 1024);
    uint8_t* s_lut = smem + G.off_lut;
    float4* s_f32lut = reinterpret_cast<float4*>(smem + G.off_f32lut);
    unsigned long long* s_red = reinterpret_cast<unsigned long long*>(smem + G.off_red);
    const uint32_t bar = smem_u32(smem + G.off_bar);

    const int tid = threadIdx.x, nthr = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int row_bytes = w * 3;
    const int prb = w >> 3;                                  // plane bytes per row
    const uint32_t frame_bytes = (uint32_t)h * row_bytes;
    const int plane_words = h * ww;
    const int MS = G.mag_stride;

    // thread -> (strip, segment): an 8-lane group is 8 adjacent strips (32 pixels) of one segment
    const int grp = lane >> 3;
    const int strip = 8 * (warp % ww) + (lane & 7);
    const int seg = 4 * (warp / ww) + grp;
    const int r0 = seg * G.seg_rows;
    const int r1 = min(h, r0 + G.seg_rows);
    const bool seg_ok = r0 < h;
    const int nstrips = w >> 2;
    const bool store_lane = seg_ok && !(lane & 1);           // even lanes store the byte shared with the odd neighbour

    // ---- one-time tables and zero borders --------------------------------------------------------------------
    for (int i = tid; i < 256; i += nthr) {
        s_sdiv[i] = i ? __double2int_rn((double)(255 << 12) / (double)i) : 0;
        const int hd = i ? __double2int_rn((double)(180 << 12) / (6.0 * (double)i)) : 0;
        s_hue[i] = make_int2(hd, 2048 - 2048 * hd);           // ((h0 + 2048) * hd + (2048 - 2048 hd)) >> 12 == (h0 * hd + 2048) >> 12
        s_lut[i] = p.lut[i];
    }
    if (tid < 8) s_f32lut[tid] = make_float4((tid & 1) ? 1.0f : 0.0f, (tid & 2) ? 1.0f : 0.0f, (tid & 4) ? 1.0f : 0.0f, 0.0f);
    if (EDGE) {   // zero borders of the magnitude plane: rows 0 and h+1, columns x = -1 and x = w
        for (int i = tid; i < MS; i += nthr) { s_mag[i] = 0; s_mag[(h + 1) * MS + i] = 0; }
        for (int i = tid; i < h + 2; i += nthr) { s_mag[i * MS + 3] = 0; s_mag[i * MS + 4 + w] = 0; }
        for (int i = tid; i < ww; i += nthr) {
            s_cand[-ww + i] = 0; s_cand[h * ww + i] = 0;
            s_edge[-ww + i] = 0; s_edge[h * ww + i] = 0;
        }
    }
    if (tid == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();

    auto issue_load = [&](int f) {
        // one elected thread: arm the barrier with the byte count, then one bulk copy per 16 KB piece
        fence_proxy_async();
        mbar_expect_tx(bar, frame_bytes);
        const uint8_t* src = p.in + (size_t)f * frame_bytes;
        for (uint32_t o = 0; o < frame_bytes; o += 16384u)
            tma_load_1d(smem_u32(s_pix) + o, src + o, min(16384u, frame_bytes - o), bar);
    };
    if (tid == 0 && (int)blockIdx.x < p.n) issue_load(blockIdx.x);

    unsigned long long st_mask[NR > 0 ? NR : 1];
#pragma unroll
    for (int k = 0; k < (NR > 0 ? NR : 1); ++k) st_mask[k] = 0;
    unsigned long long st_edge = 0, st_strong = 0, st_cand = 0, st_sweeps = 0, st_roi = 0, st_frames = 0;
    uint32_t phase = 0;

    // per-thread constants of the strip walk
    const bool left_edge = strip == 0, right_edge = strip == nstrips - 1;
    const int offL = left_edge ? 0 : -4;                    // word holding the pixel left of the strip (replicated at x = 0)
    const int offR = right_edge ? 8 : 12;                   // word holding the pixel right of the strip (replicated at x = w-1)
    const uint32_t selL0 = 0x5450u | (left_edge ? 0u : 1u), selL1 = 0x5450u | (left_edge ? 1u : 2u), selL2 = 0x5450u | (left_edge ? 2u : 3u);
    const uint32_t selR0 = 0x1012u | ((right_edge ? 5u : 4u) << 8), selR1 = 0x1012u | ((right_edge ? 6u : 5u) << 8),
                   selR2 = 0x1012u | ((right_edge ? 7u : 6u) << 8);
    const int nsteps = G.seg_rows + 2;

    for (int f = blockIdx.x; f < p.n; f += gridDim.x) {
        mbar_wait(bar, phase);
        phase ^= 1u;
        const bool use_lut = p.dynamic || !p.lut_identity;

        // ---- brightness / contrast on the resident frame (only when the table is not the identity) ----------
        if (p.dynamic) {
            const int y0 = min(40, h), y1 = min(119, h);
            unsigned long long s0 = 0, s1 = 0, s2 = 0;
            const int npix = (y1 - y0) * w;
            const uint8_t* roi = s_pix + y0 * row_bytes;
            for (int i = tid; i < npix; i += nthr) { s0 += roi[3 * i]; s1 += roi[3 * i + 1]; s2 += roi[3 * i + 2]; }
            for (int o = 16; o; o >>= 1) {
                s0 += __shfl_xor_sync(0xffffffffu, s0, o);
                s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                s2 += __shfl_xor_sync(0xffffffffu, s2, o);
            }
            if (tid < 3) s_red[tid] = 0;
            __syncthreads();
            if (lane == 0) { atomicAdd(&s_red[0], s0); atomicAdd(&s_red[1], s1); atomicAdd(&s_red[2], s2); }
            __syncthreads();
            const float fdelta = (float)brightness_delta(s_red[0], s_red[1], s_red[2], (double)npix, p.baseline);
            if (tid == 0) st_roi += s_red[0] + s_red[1] + s_red[2];
            for (int i = tid; i < 256; i += nthr) s_lut[i] = adjust_entry(i, true, fdelta, p.foff, p.fratio);
            __syncthreads();
        }
        if (use_lut) {
            uint32_t* px = reinterpret_cast<uint32_t*>(s_pix);
            for (int i = tid; i < (int)(frame_bytes >> 2); i += nthr) px[i] = lut4(s_lut, px[i]);
            __syncthreads();
        }

        // ---- P1: strip walk — Sobel / magnitude / direction -> s_mag, colour masks -> bit planes --------------
        if (EDGE || NR > 0) {
            uint32_t D[3][6], Hs[3][6];      // rolling rows: horizontal difference and horizontal smoothing, packed pairs
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 6; ++j) { D[i][j] = 0; Hs[i][j] = 0; }
            const uint8_t* strip_base = s_pix + 12 * strip;
            uint8_t* mask_base = reinterpret_cast<uint8_t*>(s_mask) + (strip >> 1);

            auto row_step = [&](int k, uint32_t (&Dn)[6], uint32_t (&Hn)[6], const uint32_t (&D0)[6], const uint32_t (&D1)[6],
                                const uint32_t (&H0)[6]) {
                // load image row y_load = r0 - 1 + k (clamped: replicated border) into slot n; emit output row y = r0 + k - 2
                const int y_load = min(max(r0 - 1 + k, 0), h - 1);
                const uint8_t* rp = strip_base + y_load * row_bytes;
                const uint32_t w0 = *reinterpret_cast<const uint32_t*>(rp);
                const uint32_t w1 = *reinterpret_cast<const uint32_t*>(rp + 4);
                const uint32_t w2 = *reinterpret_cast<const uint32_t*>(rp + 8);
                const uint32_t wl = *reinterpret_cast<const uint32_t*>(rp + offL);
                const uint32_t wr = *reinterpret_cast<const uint32_t*>(rp + offR);
                // interleaved bytes -> zero-extended pairs: E = (b0,b2), O = (b1,b3)
                const uint32_t E0 = w0 & 0x00ff00ffu, O0 = prmt(w0, 0, 0x4341);
                const uint32_t E1 = w1 & 0x00ff00ffu, O1 = prmt(w1, 0, 0x4341);
                const uint32_t E2 = w2 & 0x00ff00ffu, O2 = prmt(w2, 0, 0x4341);
                // planar pairs: A = pixels (0,2), B = pixels (1,3) of the strip, per channel
                uint32_t A[3], B[3];
                A[0] = prmt(E0, E1, 0x7610); A[1] = prmt(O0, O1, 0x7610); A[2] = prmt(E0, E2, 0x5432);
                B[0] = prmt(O0, O2, 0x5432); B[1] = prmt(E1, E2, 0x7610); B[2] = prmt(O1, O2, 0x7610);
                if (EDGE) {
                    // neighbours: Lh = pixels (-1,1), Rh = pixels (2,4)
                    uint32_t Lh[3], Rh[3];
                    Lh[0] = prmt(wl, B[0], selL0); Lh[1] = prmt(wl, B[1], selL1); Lh[2] = prmt(wl, B[2], selL2);
                    Rh[0] = prmt(A[0], wr, selR0); Rh[1] = prmt(A[1], wr, selR1); Rh[2] = prmt(A[2], wr, selR2);
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        Dn[c] = hsub(B[c], Lh[c]);                      // pixels (0,2): p[x+1] - p[x-1]
                        Dn[3 + c] = hsub(Rh[c], A[c]);                  // pixels (1,3)
                        Hn[c] = hadd(hx2p(A[c], Lh[c]), B[c]);          // p[x-1] + 2 p[x] + p[x+1]
                        Hn[3 + c] = hadd(hx2p(B[c], A[c]), Rh[c]);
                    }
                }
                const int y_row = r0 - 1 + k;
                const bool row_in = k >= 1 && y_row < r1;                // the loaded row belongs to this segment
                // ---- colour masks for the loaded row ---------------------------------------------------------
                if (NR > 0) {
                    uint32_t okm[NR > 0 ? NR : 1][2];                    // per range: half masks for pixels (0,2) and (1,3)
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const uint32_t* X = half ? B : A;
                        const uint32_t v2 = __vimax3_u16x2(X[0], X[1], X[2]);
                        const uint32_t mn2 = __vimin3_u16x2(X[0], X[1], X[2]);
                        const uint32_t d2 = v2 - mn2;
                        uint32_t s2, hh2 = 0;
                        {
                            const uint32_t vlo = v2 & 0xffffu, vhi = v2 >> 16, dlo = d2 & 0xffffu, dhi = d2 >> 16;
                            const uint32_t slo = (uint32_t)(((int)dlo * s_sdiv[vlo] + 2048) >> 12);
                            const uint32_t shi = (uint32_t)(((int)dhi * s_sdiv[vhi] + 2048) >> 12);
                            s2 = slo | (shi << 16);
                            if (P.need_hue) {
                                // hue numerator + 2048 (always positive): g-b | b-r+2d | r-g+4d, chosen by v==r, then v==g
                                const uint32_t gb = X[1] + 0x08000800u - X[2];
                                const uint32_t br = X[2] + 0x08000800u - X[0] + d2 + d2;
                                const uint32_t rg = X[0] + 0x08000800u - X[1] + (d2 << 2);
                                const uint32_t eqr = heq_mask(v2, X[0]), eqg = heq_mask(v2, X[1]);
                                const uint32_t h02 = bsel(eqr, gb, bsel(eqg, br, rg));
                                const int2 tl = s_hue[dlo], th = s_hue[dhi];
                                int hlo = ((int)(h02 & 0xffffu) * tl.x + tl.y) >> 12;
                                int hhi = ((int)(h02 >> 16) * th.x + th.y) >> 12;
                                hlo += (hlo >> 31) & 180;
                                hhi += (hhi >> 31) & 180;
                                hh2 = (uint32_t)hlo | ((uint32_t)hhi << 16);
                            }
                        }
#pragma unroll
                        for (int r = 0; r < NR; ++r) {
                            const FastRange& R = P.fr[r];
                            uint32_t ok = 0xffffffffu;
                            if (R.flags & 1u) ok &= hge_mask(hh2, R.lo[0]);
                            if (R.flags & 2u) ok &= hle_mask(hh2, R.hi[0]);
                            if (R.flags & 4u) ok &= hge_mask(s2, R.lo[1]);
                            if (R.flags & 8u) ok &= hle_mask(s2, R.hi[1]);
                            if (R.flags & 16u) ok &= hge_mask(v2, R.lo[2]);
                            if (R.flags & 32u) ok &= hle_mask(v2, R.hi[2]);
                            okm[r][half] = ok;
                        }
                    }
                    uint32_t v = 0;
#pragma unroll
                    for (int r = 0; r < NR; ++r) v |= nibble_of(okm[r][0], okm[r][1]) << (8 * r);
                    const uint32_t other = __shfl_down_sync(0xffffffffu, v, 1);
                    v |= other << 4;
                    if (store_lane && row_in) {
#pragma unroll
                        for (int r = 0; r < NR; ++r) mask_base[(r * plane_stride) * 4 + y_row * prb] = (uint8_t)(v >> (8 * r));
                    }
                }
                // ---- Sobel combine for output row y = r0 + k - 2 ----------------------------------------------
                if (EDGE && k >= 2) {
                    const int y = r0 + k - 2;
                    uint32_t mg[2], dxs[2], dys[2];
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        uint32_t m[3], dx[3], dy[3];
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const int j = 3 * half + c;
                            dx[c] = hadd(hx2p(D1[j], D0[j]), Dn[j]);            // D[y-1] + 2 D[y] + D[y+1]
                            dy[c] = hsub(Hn[j], H0[j]);                          // H[y+1] - H[y-1]
                            m[c] = habsadd(dx[c], dy[c]);
                        }
                        const uint32_t g1 = hgt_mask(m[1], m[0]);                // strictly greater: the lowest channel wins ties
                        uint32_t mm = hmaxu(m[0], m[1]);
                        uint32_t bx = bsel(g1, dx[1], dx[0]), by = bsel(g1, dy[1], dy[0]);
                        const uint32_t g2 = hgt_mask(m[2], mm);
                        mm = hmaxu(mm, m[2]);
                        bx = bsel(g2, dx[2], bx); by = bsel(g2, dy[2], by);
                        mg[half] = mm; dxs[half] = bx; dys[half] = by;
                    }
                    // per pixel direction class (pixel_math.cuh canny_dir), branch-free; pixel order 0..3 = (A.lo, B.lo, A.hi, B.hi)
                    const uint32_t ax0 = dxs[0] & 0x7fff7fffu, ay0 = dys[0] & 0x7fff7fffu;
                    const uint32_t ax1 = dxs[1] & 0x7fff7fffu, ay1 = dys[1] & 0x7fff7fffu;
                    const uint32_t sd0 = ((dxs[0] ^ dys[0]) >> 15) & 0x00010001u, sd1 = ((dxs[1] ^ dys[1]) >> 15) & 0x00010001u;
                    const uint32_t c0 = dir_bits(ax0 & 0xffffu, ay0 & 0xffffu, sd0 & 1u);      // pixel 0
                    const uint32_t c1 = dir_bits(ax1 & 0xffffu, ay1 & 0xffffu, sd1 & 1u);      // pixel 1
                    const uint32_t c2 = dir_bits(ax0 >> 16, ay0 >> 16, sd0 >> 16);             // pixel 2
                    const uint32_t c3 = dir_bits(ax1 >> 16, ay1 >> 16, sd1 >> 16);             // pixel 3
                    if (seg_ok && y < r1) {
                        uint2 v;
                        v.x = prmt(mg[0], mg[1], 0x5410) | (c0 << 11) | (c1 << 27);        // (m0, m1) + codes
                        v.y = prmt(mg[0], mg[1], 0x7632) | (c2 << 11) | (c3 << 27);        // (m2, m3) + codes
                        *reinterpret_cast<uint2*>(s_mag + (y + 1) * MS + 4 + 4 * strip) = v;
                    }
                }
            };
            // rolling window by register renaming: slots (k % 3)
#pragma unroll 1
            for (int k = 0; k < nsteps; k += 3) {
                row_step(k, D[0], Hs[0], D[1], D[2], Hs[1]);                 // new = slot0, y-1 = slot1, y = slot2
                if (k + 1 < nsteps) row_step(k + 1, D[1], Hs[1], D[2], D[0], Hs[2]);
                if (k + 2 < nsteps) row_step(k + 2, D[2], Hs[2], D[0], D[1], Hs[0]);
            }
        }
        __syncthreads();
        if (!p.need_pixels && tid == 0 && f + (int)gridDim.x < p.n) issue_load(f + gridDim.x);   // pixels are dead: prefetch

        // ---- P2: non-maximum suppression, same strip walk over the magnitude plane, two pixels per compare ------
        if (EDGE) {
            const uint16_t* mbase = s_mag + 4 + 4 * strip;
            uint8_t* cbase = reinterpret_cast<uint8_t*>(s_cand) + (strip >> 1);
            uint8_t* ebase = reinterpret_cast<uint8_t*>(s_edge) + (strip >> 1);
            // one row as packed pairs of magnitudes: P01=(m0,m1) P23=(m2,m3) L01=(m-1,m0) M12=(m1,m2) R23=(m3,m4); raw keeps the codes
            struct Row { uint32_t p01, p23, l01, m12, r23, raw01, raw23; };
            auto load_row = [&](int y) {
                const uint16_t* rp = mbase + (y + 1) * MS;
                const uint2 c = *reinterpret_cast<const uint2*>(rp);
                const uint32_t ml = rp[-1] & 0x7ffu, mr = rp[4] & 0x7ffu;
                Row r;
                r.raw01 = c.x; r.raw23 = c.y;
                r.p01 = c.x & 0x07ff07ffu; r.p23 = c.y & 0x07ff07ffu;
                r.l01 = prmt(ml, r.p01, 0x5410);
                r.m12 = prmt(r.p01, r.p23, 0x5432);
                r.r23 = prmt(r.p23, mr, 0x5432);
                return r;
            };
            const int ya = seg_ok ? r0 : 0;
            Row up = load_row(ya - 1), ce = load_row(ya);
#pragma unroll 1
            for (int k = 0; k < G.seg_rows; ++k) {
                const int y = ya + k;
                const bool row_in = y < r1;
                const Row dn = load_row(min(y + 1, h));
                uint32_t cm[2], sm[2];
#pragma unroll
                for (int pr = 0; pr < 2; ++pr) {
                    const uint32_t C = pr ? ce.p23 : ce.p01;
                    const uint32_t raw = pr ? ce.raw23 : ce.raw01;
                    const uint32_t L = pr ? ce.m12 : ce.l01, Rr = pr ? ce.r23 : ce.m12;
                    const uint32_t U = pr ? up.p23 : up.p01, Dw = pr ? dn.p23 : dn.p01;
                    const uint32_t UL = pr ? up.m12 : up.l01, DR = pr ? dn.r23 : dn.m12;
                    const uint32_t UR = pr ? up.r23 : up.m12, DL = pr ? dn.m12 : dn.l01;
                    const uint32_t t0 = hgt_mask(C, L) & hge_mask(C, Rr);          // horizontal:  m > left, m >= right
                    const uint32_t t1 = hgt_mask(C, U) & hge_mask(C, Dw);          // vertical:    m > up,   m >= down
                    const uint32_t t2 = hgt_mask(C, UL) & hgt_mask(C, DR);         // diagonal s=+1, strict on both sides
                    const uint32_t t3 = hgt_mask(C, UR) & hgt_mask(C, DL);         // diagonal s=-1
                    const uint32_t b0 = ((raw >> 11) & 0x00010001u) * 0xffffu;
                    const uint32_t b1 = ((raw >> 12) & 0x00010001u) * 0xffffu;
                    const uint32_t pick = bsel(b1, bsel(b0, t3, t2), bsel(b0, t1, t0));
                    cm[pr] = pick & hgt_mask(C, P.low2);
                    sm[pr] = cm[pr] & hgt_mask(C, P.high2);
                }
                up = ce; ce = dn;
                // pixels (0,1) sit in cm[0] halves, (2,3) in cm[1]: nibble bit q = pixel q
                uint32_t x = (cm[0] & 0x00020001u) | (cm[1] & 0x00080004u);
                uint32_t z = (sm[0] & 0x00020001u) | (sm[1] & 0x00080004u);
                uint32_t v = ((x | (x >> 16)) & 0xfu) | (((z | (z >> 16)) & 0xfu) << 8);
                const uint32_t other = __shfl_down_sync(0xffffffffu, v, 1);
                v |= other << 4;
                if (store_lane && row_in) {
                    cbase[y * prb] = (uint8_t)v;
                    ebase[y * prb] = (uint8_t)(v >> 8);
                    if (p.stats) st_strong += __popc((v >> 8) & 0xffu);
                }
            }
        }
        __syncthreads();

        // ---- P3: hysteresis: grow the strong set through candidates, one plane word per thread per sweep -------
        if (EDGE) {
            volatile uint32_t* E = s_edge;
            int any;
            do {
                int changed = 0;
                for (int t = tid; t < plane_words; t += nthr) {
                    const uint32_t c = s_cand[t];
                    const uint32_t e = E[t];
                    if (c != e) {
                        const int wi = t % ww;
                        uint32_t mid = e | E[t - ww] | E[t + ww];
                        uint32_t lft = 0, rgt = 0;
                        if (wi > 0) lft = E[t - 1] | E[t - ww - 1] | E[t + ww - 1];
                        if (wi + 1 < ww) rgt = E[t + 1] | E[t - ww + 1] | E[t + ww + 1];
                        const uint32_t spread = mid | (mid << 1) | (mid >> 1) | (lft >> 31) | (rgt << 31);
                        const uint32_t ne = flood_word((spread & c) | e, c);
                        if (ne != e) { E[t] = ne; changed = 1; }
                    }
                }
                any = __syncthreads_or(changed);
                if (tid == 0) ++st_sweeps;
            } while (any);
        }

        // ---- P4: merge + normalise, written once ---------------------------------------------------------------
        {
            uint8_t* __restrict__ gout = p.out_u8 ? p.out_u8 + (size_t)f * frame_bytes : nullptr;
            float* __restrict__ gf32 = p.out_f32 ? p.out_f32 + (size_t)f * frame_bytes : nullptr;
            const uint8_t* planes[3];
#pragma unroll
            for (int c = 0; c < 3; ++c)
                planes[c] = reinterpret_cast<const uint8_t*>(p.src[c] == SRC_EDGE ? s_edge : (p.src[c] >= SRC_MASK0 ? s_mask + (p.src[c] - SRC_MASK0) * plane_stride : nullptr));
            const int npb = h * prb;                        // groups of 8 pixels = plane bytes
            if (!p.need_pixels) {
                // all three channels are bit planes: bytes by multiply-spread, floats as bit * 0x3f800000 (masks are 0.0 / 1.0)
                int poff[3];
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    poff[c] = (p.src[c] == SRC_EDGE ? G.off_edge : G.off_mask + (p.src[c] - SRC_MASK0) * plane_stride * 4) + ww * 4;
                for (int g = tid; g < npb; g += nthr) {
                    const uint32_t b0 = smem[poff[0] + g], b1 = smem[poff[1] + g], b2 = smem[poff[2] + g];     // 8 pixels of each channel
                    uint32_t r4[2], g4[2], l4[2];
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        r4[hf] = (((b0 >> (4 * hf)) & 0xfu) * 0x00204081u) & 0x01010101u;      // bit q -> byte q
                        g4[hf] = (((b1 >> (4 * hf)) & 0xfu) * 0x00204081u) & 0x01010101u;
                        l4[hf] = (((b2 >> (4 * hf)) & 0xfu) * 0x00204081u) & 0x01010101u;
                    }
                    if (gout) {
                        uint32_t wv[6];
#pragma unroll
                        for (int hf = 0; hf < 2; ++hf) {
                            const uint32_t R = r4[hf] * 255u, Gc = g4[hf] * 255u, Bc = l4[hf] * 255u;     // planar bytes 0 / 255
                            wv[3 * hf + 0] = prmt(prmt(R, Gc, 0x1040), Bc, 0x3410);     // R0 G0 B0 R1
                            wv[3 * hf + 1] = prmt(prmt(Gc, Bc, 0x2051), R, 0x3610);     // G1 B1 R2 G2
                            wv[3 * hf + 2] = prmt(prmt(Bc, R, 0x3072), Gc, 0x3710);     // B2 R3 G3 B3
                        }
                        uint2* dst = reinterpret_cast<uint2*>(gout + (size_t)g * 24);
                        dst[0] = make_uint2(wv[0], wv[1]); dst[1] = make_uint2(wv[2], wv[3]); dst[2] = make_uint2(wv[4], wv[5]);
                    }
                    if (gf32) {
                        uint4* dst = reinterpret_cast<uint4*>(gf32 + (size_t)g * 24);
                        const uint32_t one = 0x3f800000u;
#pragma unroll
                        for (int hf = 0; hf < 2; ++hf) {
                            const uint32_t R = r4[hf], Gc = g4[hf], Bc = l4[hf];
                            dst[3 * hf + 0] = make_uint4((R & 0xffu) * one, (Gc & 0xffu) * one, (Bc & 0xffu) * one, prmt(R, 0, 0x4441) * one);
                            dst[3 * hf + 1] = make_uint4(prmt(Gc, 0, 0x4441) * one, prmt(Bc, 0, 0x4441) * one, prmt(R, 0, 0x4442) * one, prmt(Gc, 0, 0x4442) * one);
                            dst[3 * hf + 2] = make_uint4(prmt(Bc, 0, 0x4442) * one, (R >> 24) * one, (Gc >> 24) * one, (Bc >> 24) * one);
                        }
                    }
                }
            } else {
                // some channel keeps the adjusted pixel: bytes from the resident frame, floats by correctly rounded x/255
                const float rcp = 1.0f / 255.0f;
                for (int g = tid; g < 2 * npb; g += nthr) {          // groups of 4 pixels
                    const uint32_t* src = reinterpret_cast<const uint32_t*>(s_pix + (size_t)g * 12);
                    uint32_t wv[3] = {src[0], src[1], src[2]};
                    uint8_t b[12];
#pragma unroll
                    for (int k = 0; k < 12; ++k) b[k] = (uint8_t)(wv[k >> 2] >> ((k & 3) * 8));
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        if (planes[c]) {
                            const uint32_t bits = (uint32_t)planes[c][g >> 1] >> ((g & 1) * 4);
#pragma unroll
                            for (int q = 0; q < 4; ++q) b[q * 3 + c] = ((bits >> q) & 1u) ? 255 : 0;
                        }
                    }
                    if (gout) {
                        uint32_t* dst = reinterpret_cast<uint32_t*>(gout + (size_t)g * 12);
#pragma unroll
                        for (int k = 0; k < 3; ++k)
                            dst[k] = (uint32_t)b[4 * k] | ((uint32_t)b[4 * k + 1] << 8) | ((uint32_t)b[4 * k + 2] << 16) | ((uint32_t)b[4 * k + 3] << 24);
                    }
                    if (gf32) {
                        float fv[12];
#pragma unroll
                        for (int k = 0; k < 12; ++k) {
                            // x/255 correctly rounded without a division: q0 = x*rcp, one fused residual correction (exhaustively checked for 0..255)
                            const float x = (float)b[k];
                            const float q0 = __fmul_rn(x, rcp);
                            fv[k] = __fmaf_rn(__fmaf_rn(-q0, 255.0f, x), rcp, q0);
                        }
                        float4* dst = reinterpret_cast<float4*>(gf32 + (size_t)g * 12);
                        dst[0] = make_float4(fv[0], fv[1], fv[2], fv[3]);
                        dst[1] = make_float4(fv[4], fv[5], fv[6], fv[7]);
                        dst[2] = make_float4(fv[8], fv[9], fv[10], fv[11]);
                    }
                }
            }
        }
        if (p.stats) {
            for (int i = tid; i < plane_words; i += nthr) {
                if (EDGE) { st_edge += __popc(s_edge[i]); st_cand += __popc(s_cand[i]); }
#pragma unroll
                for (int k = 0; k < NR; ++k) st_mask[k] += __popc(s_mask[k * plane_stride + i]);
            }
            if (tid == 0) ++st_frames;
        }
        __syncthreads();
        if (p.need_pixels && tid == 0 && f + (int)gridDim.x < p.n) issue_load(f + gridDim.x);
    }

    if (p.stats) {
        unsigned long long v[10] = {st_frames, 0, 0, 0, 0, st_edge, st_strong, st_cand, st_sweeps, st_roi};
#pragma unroll
        for (int k = 0; k < NR; ++k) v[1 + p.range_stat[k]] = st_mask[k];
#pragma unroll
        for (int k = 0; k < 10; ++k) {
            unsigned long long x = v[k];
            for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            if (lane == 0 && x) atomicAdd(&p.stats[k], x);
        }
    }
}

}  // namespace trs
