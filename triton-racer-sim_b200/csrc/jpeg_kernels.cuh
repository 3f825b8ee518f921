// jpeg_kernels.cuh — batched baseline-JPEG decode of tub records on the GPU (sm_100a); arithmetic in jpeg_core.cuh.
//
//   k_jpeg_entropy        one THREAD per record: Huffman decoding is sequential within a scan (no restart markers in the
//                         recorder's files), so the parallelism is across the N records of the batch.  The bit stream comes in
//                         8-byte aligned loads one chunk ahead, the decoding tables sit in shared memory.  A block's coefficients
//                         stay in the thread (natural order); at the end of the block - where the lanes of a warp meet again -
//                         the thread dequantises and runs the integer IDCT itself and writes the 8x8 samples into planar
//                         Y / Cb / Cr (MCU-padded): no coefficient buffer (61 KB per 120x160 record) is written, zeroed or re-read.
//   k_jpeg_upsample_rgb   one thread per 4 output pixels: triangle-filter chroma upsampling + YCbCr -> RGB, interleaved u8 out
//                         (the (N,H,W,3) layout the rest of the path reads).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "jpeg_core.cuh"

namespace trs {

struct JpegRecord {
    uint64_t data_off;      // entropy-coded segment inside the uploaded blob
    uint32_t data_len;
    uint32_t table_set;     // index into the uploaded JpegTables array
};

struct JpegPlanes {
    uint8_t* y;             // per record: (mh*8*vs) x (mw*8*hs)
    uint8_t* cb;            // per record: (mh*8) x (mw*8)
    uint8_t* cr;
    int mw, mh;             // MCUs per row / column
    int hs, vs;             // luma sampling factors of the batch: (2,2) 4:2:0, (2,1) 4:2:2, (1,1) 4:4:4
};

enum { JPG_THREADS = 64 };

// ---- device-side bit reader: 8-byte aligned global loads one chunk ahead, bytes handed out from registers ----------------
struct JpegStream {
    const uint2* p;        // next chunk to load
    uint64_t cur, nxt;     // current chunk (bytes are taken from the low end), the chunk after it (already loaded)
    int pos;               // next byte of `cur` (0..7)
    int left;              // bytes of the entropy-coded segment not yet consumed
    uint64_t buf;          // bit buffer, consumed from the top
    int n;                 // valid bits in buf
};

__device__ __forceinline__ uint64_t jpg_ld8(const uint2* p)
{
    const uint2 v = __ldg(p);
    return ((uint64_t)v.y << 32) | v.x;
}

__device__ __forceinline__ void jpg_stream_open(JpegStream& s, const uint8_t* data, uint32_t len)
{
    const uintptr_t a = reinterpret_cast<uintptr_t>(data);
    s.p = reinterpret_cast<const uint2*>(a & ~(uintptr_t)7);
    s.pos = (int)(a & 7);
    s.cur = jpg_ld8(s.p);
    s.nxt = jpg_ld8(s.p + 1);                 // (the blob has 16 bytes of slack after the last file)
    s.p += 2;
    s.left = (int)len;
    s.buf = 0;
    s.n = 0;
}

__device__ __forceinline__ uint32_t jpg_stream_byte(JpegStream& s)          // next raw byte (caller checks s.left)
{
    const uint32_t c = (uint32_t)(s.cur >> (8 * s.pos)) & 0xffu;
    if (++s.pos == 8) { s.pos = 0; s.cur = s.nxt; s.nxt = jpg_ld8(s.p); ++s.p; }
    --s.left;
    return c;
}

__device__ __forceinline__ void jpg_stream_fill(JpegStream& s)              // top up to more than 32 valid bits
{
    // fast path: the next four bytes at once when none of them is 0xff (no stuffed zero to drop, no marker to stop at)
    if (s.n <= 32 && s.left >= 4) {
        const int sh = 8 * s.pos;
        uint32_t w = (uint32_t)(s.cur >> sh);
        if (s.pos > 4) w |= (uint32_t)(s.nxt << (64 - sh));
        const uint32_t v = ~w;
        if (((v - 0x01010101u) & ~v & 0x80808080u) == 0) {                     // ~w has no zero byte <=> w has no 0xff byte
            s.buf |= (uint64_t)__byte_perm(w, 0, 0x0123) << (32 - s.n);        // stream order is big-endian
            s.n += 32;
            s.pos += 4;
            s.left -= 4;
            if (s.pos >= 8) { s.pos -= 8; s.cur = s.nxt; s.nxt = jpg_ld8(s.p); ++s.p; }
        }
    }
#pragma unroll 1
    while (s.n <= 32) {
        uint32_t c = 0;
        if (s.left > 0) {
            c = jpg_stream_byte(s);
            if (c == 0xffu) {
                uint32_t d = 0xffu;
                if (s.left > 0) d = (uint32_t)(s.cur >> (8 * s.pos)) & 0xffu;      // peek
                if (d == 0) jpg_stream_byte(s);                                     // stuffed zero: a data byte 0xff
                else { c = 0; s.left = 0; }                                         // marker: stop consuming, pad with zeros
            }
        }
        s.buf |= (uint64_t)c << (56 - s.n);
        s.n += 8;
    }
}

// decoding tables of one set in shared memory: look-ahead packed as (nbits << 8 | symbol)
struct JpegSmemTables {
    uint16_t look[4][256];         // dc luma, ac luma, dc chroma, ac chroma
    int32_t maxcode[4][18];
    int32_t valoffset[4][17];
    uint8_t huffval[4][256];
    uint8_t natural[64];
};

__device__ __forceinline__ int jpg_dev_symbol(JpegStream& s, const JpegSmemTables& T, int t, int& err)
{
    if (s.n < 16) jpg_stream_fill(s);
    const uint32_t e = T.look[t][(uint32_t)(s.buf >> 56)];
    if (e) { const int l = (int)(e >> 8); s.buf <<= l; s.n -= l; return (int)(e & 0xffu); }
    int l = 9;
    int32_t code = (int32_t)(s.buf >> 55);
    while (l <= 16 && code > T.maxcode[t][l]) { ++l; code = (int32_t)(s.buf >> (64 - l)); }
    if (l > 16) { err = JPG_E_BADCODE; return 0; }
    s.buf <<= l; s.n -= l;
    return T.huffval[t][(code + T.valoffset[t][l]) & 0xff];
}

__device__ __forceinline__ int jpg_dev_extend(JpegStream& s, int nb)
{
    if (s.n < nb) jpg_stream_fill(s);
    const int r = (int)(s.buf >> (64 - nb));
    s.buf <<= nb; s.n -= nb;
    return r < (1 << (nb - 1)) ? r - (1 << nb) + 1 : r;
}

// Entropy decoding: one thread per record; the non-zero quantised coefficients go to a pre-zeroed (record, block, 64) int16 buffer
// in natural order.  Block order inside a record: MCU-major, then the blocks of the MCU (luma in raster order, Cb, Cr).
__global__ void __launch_bounds__(JPG_THREADS) k_jpeg_entropy(const uint8_t* __restrict__ blob, const JpegRecord* __restrict__ recs,
                                                             const JpegTables* __restrict__ tables, int n, JpegPlanes P, int* __restrict__ status)
{
    __shared__ JpegSmemTables T;
    // A CTA whose records disagree on the table set loads the sets one after the other (common case: one set for the whole batch).
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = r < n;
    const JpegRecord rec = live ? recs[r] : JpegRecord{0, 0, 0xffffffffu};
    // table sets present in this CTA, processed one after the other (one iteration when the batch has a single set)
    __shared__ uint32_t s_next;
    bool mine_done = !live;
    int err = 0;
    while (true) {
        if (threadIdx.x == 0) s_next = 0xffffffffu;
        __syncthreads();
        if (!mine_done) atomicMin(&s_next, rec.table_set);
        __syncthreads();
        const uint32_t set = s_next;
        if (set == 0xffffffffu) break;
        const JpegTables& G = tables[set];
        for (int i = threadIdx.x; i < 4 * 256; i += blockDim.x) {
            const int t = i >> 8, k = i & 255;
            const JpegHuff& h = t == 0 ? G.dc[0] : (t == 1 ? G.ac[0] : (t == 2 ? G.dc[1] : G.ac[1]));
            T.look[t][k] = (uint16_t)((h.look_nbits[k] << 8) | h.look_sym[k]);
            T.huffval[t][k] = h.huffval[k];
            if (k < 18) T.maxcode[t][k] = h.maxcode[k];
            if (k < 17) T.valoffset[t][k] = h.valoffset[k];
        }
        if (threadIdx.x < 64) T.natural[threadIdx.x] = jpg_natural_order[threadIdx.x];
        __syncthreads();
        if (!mine_done && rec.table_set == set) {
            JpegStream s;
            jpg_stream_open(s, blob + rec.data_off, rec.data_len);
            int dc[3] = {0, 0, 0};
            const int luma_blocks = P.hs * P.vs, n_mcu = P.mw * P.mh;
            const int bpm = luma_blocks + 2;                       // blocks per MCU: the luma blocks in raster order, then Cb, Cr
            const int ys = P.mw * 8 * P.hs, cs = P.mw * 8;
            uint8_t* const Yp = P.y + (size_t)r * ys * P.mh * 8 * P.vs;
            uint8_t* const Cbp = P.cb + (size_t)r * cs * P.mh * 8;
            uint8_t* const Crp = P.cr + (size_t)r * cs * P.mh * 8;
            int sub = 0, mx = 0, my = 0;
#pragma unroll 1
            for (int blk = 0; blk < n_mcu * bpm; ++blk) {
                const int comp = sub < luma_blocks ? 0 : sub - luma_blocks + 1;
                const int td = comp ? 2 : 0, ta = td + 1;
                int16_t coef[64];
#pragma unroll
                for (int k = 0; k < 64; ++k) coef[k] = 0;
                int sym = jpg_dev_symbol(s, T, td, err);
                if (sym > 11) { err = JPG_E_BADCODE; sym = 0; }     // a DC category above 11 cannot occur in 8-bit baseline data (malformed DHT)
                if (sym) dc[comp] += jpg_dev_extend(s, sym);
                coef[0] = (int16_t)dc[comp];
#pragma unroll 1
                for (int k = 1; k < 64; ++k) {
                    sym = jpg_dev_symbol(s, T, ta, err);
                    const int run = sym >> 4;
                    sym &= 15;
                    if (sym) {
                        k += run;
                        if (k > 63) { err = JPG_E_BADCODE; break; }
                        coef[T.natural[k]] = (int16_t)jpg_dev_extend(s, sym);
                    } else {
                        if (run != 15) break;
                        k += 15;
                    }
                }
                // the lanes of the warp are together again here: dequantisation + IDCT of this block with all of them active
                if (comp == 0) {
                    const int by = sub / P.hs, bx = sub - by * P.hs;
                    jpg_idct_islow(coef, G.quant[0], Yp + ((my * P.vs + by) * 8) * ys + (mx * P.hs + bx) * 8, ys);
                } else {
                    jpg_idct_islow(coef, G.quant[1], (comp == 1 ? Cbp : Crp) + my * 8 * cs + mx * 8, cs);
                }
                if (++sub == bpm) { sub = 0; if (++mx == P.mw) { mx = 0; ++my; } }
            }
            mine_done = true;
        }
        __syncthreads();
    }
    if (err) atomicMax(status, err);
}

__global__ void __launch_bounds__(256) k_jpeg_upsample_rgb(JpegPlanes P, int n, int h, int w, uint8_t* __restrict__ out)
{
    const int ys = P.mw * 8 * P.hs, cs = P.mw * 8;
    const int cw = (w + P.hs - 1) / P.hs, ch = (h + P.vs - 1) / P.vs;
    const bool h2v2 = P.hs == 2 && P.vs == 2;
    const int gpr = (w + 3) / 4;                                   // groups of 4 pixels per row
    const size_t total = (size_t)n * h * gpr;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const bool words = (w & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0;
    // (record, row, group of four pixels) of this thread's items: one division at the start, then the grid stride with carries
    const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    int g = (int)(i0 % gpr), y = (int)((i0 / gpr) % h);
    size_t r = i0 / gpr / h;
    const int sg = (int)(stride % gpr), sy = (int)((stride / gpr) % h);
    const size_t sr = stride / gpr / h;
    for (size_t i = i0; i < total; i += stride, g += sg, y += sy, r += sr) {
        if (g >= gpr) { g -= gpr; ++y; }
        if (y >= h) { y -= h; ++r; }
        const uint8_t* Y = P.y + r * (size_t)ys * P.mh * 8 * P.vs + (size_t)y * ys;
        const uint8_t* Cb = P.cb + r * (size_t)cs * P.mh * 8;
        const uint8_t* Cr = P.cr + r * (size_t)cs * P.mh * 8;
        uint8_t px[12];
        const int x0 = 4 * g;
        if (words && cw > 2 && h2v2) {
            // four pixels share four chroma columns: column sums 3 * near row + far row once, then the horizontal triangle filter
            // (jdsample.c h2v2_fancy_upsample; the edge formulas equal the general one with the edge column replicated)
            const int cy = y >> 1, cx0 = x0 >> 1;
            const int ny = (y & 1) ? min(cy + 1, ch - 1) : max(cy - 1, 0);
            const int xl = max(cx0 - 1, 0), xr1 = min(cx0 + 1, cw - 1), xr2 = min(cx0 + 2, cw - 1);
            int up[2][4];
#pragma unroll
            for (int pl = 0; pl < 2; ++pl) {
                const uint8_t* r0 = (pl ? Cr : Cb) + cy * cs;
                const uint8_t* r1 = (pl ? Cr : Cb) + ny * cs;
                const int tm = 3 * r0[xl] + r1[xl], t0 = 3 * r0[cx0] + r1[cx0], t1 = 3 * r0[xr1] + r1[xr1], t2 = 3 * r0[xr2] + r1[xr2];
                up[pl][0] = (3 * t0 + tm + 8) >> 4;
                up[pl][1] = (3 * t0 + t1 + 7) >> 4;
                up[pl][2] = (3 * t1 + t0 + 8) >> 4;
                up[pl][3] = (3 * t1 + t2 + 7) >> 4;
            }
            const uint32_t yw = *reinterpret_cast<const uint32_t*>(Y + x0);
#pragma unroll
            for (int q = 0; q < 4; ++q)
                jpg_ycc_to_rgb((int)((yw >> (8 * q)) & 0xffu), up[0][q], up[1][q], px[3 * q], px[3 * q + 1], px[3 * q + 2]);
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int x = min(x0 + q, w - 1);
                jpg_ycc_to_rgb(Y[x], jpg_chroma_at(Cb, cs, cw, ch, P.hs, P.vs, x, y), jpg_chroma_at(Cr, cs, cw, ch, P.hs, P.vs, x, y), px[3 * q], px[3 * q + 1],
                               px[3 * q + 2]);
            }
        }
        uint8_t* o = out + ((r * h + y) * (size_t)w + x0) * 3;
        if (words) {
            uint32_t* ow = reinterpret_cast<uint32_t*>(o);
#pragma unroll
            for (int k = 0; k < 3; ++k) ow[k] = (uint32_t)px[4 * k] | ((uint32_t)px[4 * k + 1] << 8) | ((uint32_t)px[4 * k + 2] << 16) | ((uint32_t)px[4 * k + 3] << 24);
        } else {
            const int nb = 3 * min(4, w - x0);
            for (int k = 0; k < nb; ++k) o[k] = px[k];
        }
    }
}

}  // namespace trs
