// jpeg_core.cuh — baseline JPEG decoding arithmetic for tub records, written once for host and device.
//
// Replaces `np.asarray(Image.open(img_path))` in the reference's tub loaders (TritonRacerSim/components/keras_train.py:41,309),
// i.e. Pillow -> libjpeg(-turbo) with its default decompression settings, for the files the reference's recorder writes
// (`Image.fromarray(img).save(path)`, components/datastorage.py:78: baseline sequential, 8 bit, YCbCr 4:2:0, one interleaved
// scan, no restart intervals) and for the 4:2:2 / 4:4:4 variants of the same.  Every stage restates the published libjpeg algorithm the default path runs, bit for bit:
//   entropy decoding          jdhuff.c decode_mcu (canonical Huffman codes, HUFF_EXTEND, zigzag order)
//   dequantisation + IDCT     jidctint.c jpeg_idct_islow (JDCT_ISLOW, CONST_BITS 13, PASS1_BITS 2)
//   chroma upsampling         jdsample.c h2v2_fancy_upsample / h2v1_fancy_upsample (triangle filters, edge replication)
//   colour conversion         jdcolor.c ycc_rgb_convert (16-bit fixed-point tables)
// The functions are pure (memory in, memory out, no threads), so the same source is compiled by nvcc into the kernels and by
// g++ into a host-side check against Pillow (tests/test_jpeg_host.py).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define TRS_JHD __host__ __device__ __forceinline__
#else
#define TRS_JHD inline
#endif

namespace trs {

enum { JPG_OK = 0, JPG_E_TRUNCATED = 1, JPG_E_BADCODE = 2 };

// one Huffman table in decoding form: 8-bit look-ahead plus the canonical-code arrays for longer codes (jdhuff.c jpeg_make_d_derived_tbl)
struct JpegHuff {
    uint8_t look_nbits[256];     // code length if the next 8 bits start with a code of <= 8 bits, else 0
    uint8_t look_sym[256];
    int32_t maxcode[18];         // largest code of length l (-1 if none); maxcode[17] is a sentinel
    int32_t valoffset[17];       // huffval index of the first code of length l, minus that code
    uint8_t huffval[256];
};

struct JpegTables {
    uint16_t quant[2][64];       // luma / chroma quantisation steps in NATURAL (row-major) order
    JpegHuff dc[2], ac[2];       // [0] luma, [1] chroma
};

// zigzag position k -> natural index (jutils.c jpeg_natural_order)
#if defined(__CUDA_ARCH__)
__device__ __constant__
#else
static const
#endif
uint8_t jpg_natural_order[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                 41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

// the same table for host-only code (parsing): a function so that it does not depend on the device copy above
inline int jpg_natural_order_host(int k)
{
    static const uint8_t t[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                  41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                  30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
    return t[k];
}

// ---- bit reader over entropy-coded data (0xFF00 byte stuffing; a marker ends the data: zero bits are supplied) -----------
struct JpegBits {
    const uint8_t* p;
    const uint8_t* end;
    uint64_t buf;       // bits are consumed from the top
    int n;              // valid bits in buf
};

TRS_JHD void jpg_fill(JpegBits& b)
{
    while (b.n <= 56) {
        uint32_t c = 0;
        if (b.p < b.end) {
            c = *b.p;
            if (c == 0xff) {
                if (b.p + 1 < b.end && b.p[1] == 0) b.p += 2;            // stuffed zero: a data byte 0xff
                else { c = 0; b.end = b.p; }                              // marker (EOI): stop consuming, pad with zeros
            } else {
                ++b.p;
            }
        }
        b.buf |= (uint64_t)c << (56 - b.n);
        b.n += 8;
    }
}
TRS_JHD uint32_t jpg_peek(JpegBits& b, int nbits) { return (uint32_t)(b.buf >> (64 - nbits)); }
TRS_JHD void jpg_skip(JpegBits& b, int nbits) { b.buf <<= nbits; b.n -= nbits; }

// next Huffman symbol (jdhuff.c HUFF_DECODE: look-ahead, then one bit at a time against maxcode[])
TRS_JHD int jpg_symbol(JpegBits& b, const JpegHuff& h, int& err)
{
    if (b.n < 16) jpg_fill(b);
    const uint32_t look = jpg_peek(b, 8);
    int l = h.look_nbits[look];
    if (l) { jpg_skip(b, l); return h.look_sym[look]; }
    l = 9;
    int32_t code = (int32_t)jpg_peek(b, 9);
    while (l <= 16 && code > h.maxcode[l]) { ++l; code = (int32_t)jpg_peek(b, l); }
    if (l > 16) { err = JPG_E_BADCODE; return 0; }
    jpg_skip(b, l);
    return h.huffval[(code + h.valoffset[l]) & 0xff];
}

// `s` more bits as a signed value (jdhuff.c HUFF_EXTEND)
TRS_JHD int jpg_receive_extend(JpegBits& b, int s)
{
    if (b.n < s) jpg_fill(b);
    const int r = (int)jpg_peek(b, s);
    jpg_skip(b, s);
    return r < (1 << (s - 1)) ? r - (1 << s) + 1 : r;
}

// one 8x8 block of quantised coefficients in natural order; last_dc is the running DC predictor of the component
TRS_JHD void jpg_decode_block(JpegBits& b, const JpegHuff& dc, const JpegHuff& ac, int& last_dc, int16_t (&coef)[64], int& err)
{
    for (int i = 0; i < 64; ++i) coef[i] = 0;
    int s = jpg_symbol(b, dc, err);
    if (s > 11) { err = JPG_E_BADCODE; return; }          // DC categories of 8-bit baseline data are 0..11
    if (s) last_dc += jpg_receive_extend(b, s);
    coef[0] = (int16_t)last_dc;
    for (int k = 1; k < 64; ++k) {
        s = jpg_symbol(b, ac, err);
        const int r = s >> 4;
        s &= 15;
        if (s) {
            k += r;
            if (k > 63) { err = JPG_E_BADCODE; return; }
            coef[jpg_natural_order[k]] = (int16_t)jpg_receive_extend(b, s);
        } else {
            if (r != 15) break;
            k += 15;
        }
    }
}

// ---- jidctint.c jpeg_idct_islow: dequantise, 2-D inverse DCT, level shift, clamp ---------------------------------------------
TRS_JHD int jpg_descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

TRS_JHD void jpg_idct_islow(const int16_t (&coef)[64], const uint16_t* quant, uint8_t* out, int out_stride)
{
    const int CB = 13, P1 = 2;
    const int F_0_298631336 = 2446, F_0_390180644 = 3196, F_0_541196100 = 4433, F_0_765366865 = 6270, F_0_899976223 = 7373,
              F_1_175875602 = 9633, F_1_501321110 = 12299, F_1_847759065 = 15137, F_1_961570560 = 16069, F_2_053119869 = 16819,
              F_2_562915447 = 20995, F_3_072711026 = 25172;
    int ws[64];
    for (int c = 0; c < 8; ++c) {                                                   // pass 1: columns
        const int i0 = coef[c] * quant[c], i1 = coef[8 + c] * quant[8 + c], i2 = coef[16 + c] * quant[16 + c], i3 = coef[24 + c] * quant[24 + c],
                  i4 = coef[32 + c] * quant[32 + c], i5 = coef[40 + c] * quant[40 + c], i6 = coef[48 + c] * quant[48 + c],
                  i7 = coef[56 + c] * quant[56 + c];
        int z2 = i2, z3 = i6;
        int z1 = (z2 + z3) * F_0_541196100;
        int tmp2 = z1 + z3 * (-F_1_847759065);
        int tmp3 = z1 + z2 * F_0_765366865;
        z2 = i0; z3 = i4;
        int tmp0 = (z2 + z3) * (1 << CB);
        int tmp1 = (z2 - z3) * (1 << CB);
        const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
        tmp0 = i7; tmp1 = i5; tmp2 = i3; tmp3 = i1;
        z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
        int z4 = tmp1 + tmp3;
        const int z5 = (z3 + z4) * F_1_175875602;
        tmp0 *= F_0_298631336; tmp1 *= F_2_053119869; tmp2 *= F_3_072711026; tmp3 *= F_1_501321110;
        z1 *= -F_0_899976223; z2 *= -F_2_562915447; z3 *= -F_1_961570560; z4 *= -F_0_390180644;
        z3 += z5; z4 += z5;
        tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
        ws[c] = jpg_descale(tmp10 + tmp3, CB - P1);      ws[56 + c] = jpg_descale(tmp10 - tmp3, CB - P1);
        ws[8 + c] = jpg_descale(tmp11 + tmp2, CB - P1);  ws[48 + c] = jpg_descale(tmp11 - tmp2, CB - P1);
        ws[16 + c] = jpg_descale(tmp12 + tmp1, CB - P1); ws[40 + c] = jpg_descale(tmp12 - tmp1, CB - P1);
        ws[24 + c] = jpg_descale(tmp13 + tmp0, CB - P1); ws[32 + c] = jpg_descale(tmp13 - tmp0, CB - P1);
    }
    for (int r = 0; r < 8; ++r) {                                                   // pass 2: rows
        const int* w = ws + 8 * r;
        int z2 = w[2], z3 = w[6];
        int z1 = (z2 + z3) * F_0_541196100;
        int tmp2 = z1 + z3 * (-F_1_847759065);
        int tmp3 = z1 + z2 * F_0_765366865;
        int tmp0 = (w[0] + w[4]) * (1 << CB);
        int tmp1 = (w[0] - w[4]) * (1 << CB);
        const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
        tmp0 = w[7]; tmp1 = w[5]; tmp2 = w[3]; tmp3 = w[1];
        z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
        int z4 = tmp1 + tmp3;
        const int z5 = (z3 + z4) * F_1_175875602;
        tmp0 *= F_0_298631336; tmp1 *= F_2_053119869; tmp2 *= F_3_072711026; tmp3 *= F_1_501321110;
        z1 *= -F_0_899976223; z2 *= -F_2_562915447; z3 *= -F_1_961570560; z4 *= -F_0_390180644;
        z3 += z5; z4 += z5;
        tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
        const int sh = CB + P1 + 3;
        const int v[8] = {tmp10 + tmp3, tmp11 + tmp2, tmp12 + tmp1, tmp13 + tmp0, tmp13 - tmp0, tmp12 - tmp1, tmp11 - tmp2, tmp10 - tmp3};
        uint8_t* o = out + r * out_stride;
        uint32_t px[8];
        for (int k = 0; k < 8; ++k) {
            int x = jpg_descale(v[k], sh) + 128;                                    // range_limit: centred clamp to 0..255
            px[k] = (uint32_t)(x < 0 ? 0 : (x > 255 ? 255 : x));
        }
#if defined(__CUDA_ARCH__)
        if ((reinterpret_cast<uintptr_t>(o) & 7) == 0) {                             // one 8-byte store per row
            *reinterpret_cast<uint2*>(o) = make_uint2(px[0] | (px[1] << 8) | (px[2] << 16) | (px[3] << 24), px[4] | (px[5] << 8) | (px[6] << 16) | (px[7] << 24));
            continue;
        }
#endif
        for (int k = 0; k < 8; ++k) o[k] = (uint8_t)px[k];
    }
}

// ---- jdsample.c h2v2_fancy_upsample + jdcolor.c ycc_rgb_convert for ONE output pixel ------------------------------------------
// chroma plane `c` of cw x ch real samples (stride cs): the value at output position (x, y) of the 2x upsampled plane
TRS_JHD int jpg_upsample_h2v2(const uint8_t* c, int cs, int cw, int ch, int x, int y)
{
    const int cy = y >> 1, cx = x >> 1;
    if (cw <= 2) return c[cy * cs + cx];                          // jdsample.c: the fancy filter needs more than two chroma columns, else replication
    const int ny = (y & 1) ? (cy + 1 < ch ? cy + 1 : ch - 1) : (cy > 0 ? cy - 1 : 0);           // the nearer neighbouring row, replicated at the edges
    const uint8_t* r0 = c + cy * cs;
    const uint8_t* r1 = c + ny * cs;
    const int thiscol = 3 * r0[cx] + r1[cx];
    if (x & 1) {
        if (cx == cw - 1) return (thiscol * 4 + 7) >> 4;
        const int nextcol = 3 * r0[cx + 1] + r1[cx + 1];
        return (thiscol * 3 + nextcol + 7) >> 4;
    }
    if (cx == 0) return (thiscol * 4 + 8) >> 4;
    const int lastcol = 3 * r0[cx - 1] + r1[cx - 1];
    return (thiscol * 3 + lastcol + 8) >> 4;
}

// jdsample.c h2v1_fancy_upsample (4:2:2): horizontal triangle filter only, (3 * this + neighbour + 1 or 2) >> 2, edges replicated
TRS_JHD int jpg_upsample_h2v1(const uint8_t* c, int cs, int cw, int x, int y)
{
    const uint8_t* r0 = c + y * cs;
    const int cx = x >> 1;
    if (cw <= 2) return r0[cx];
    if (x & 1) return (3 * r0[cx] + r0[cx + 1 < cw ? cx + 1 : cx] + 2) >> 2;
    return (3 * r0[cx] + r0[cx > 0 ? cx - 1 : 0] + 1) >> 2;
}

// chroma sample for output pixel (x, y) under luma sampling (hs, vs) in {(2,2), (2,1), (1,1)}
TRS_JHD int jpg_chroma_at(const uint8_t* c, int cs, int cw, int ch, int hs, int vs, int x, int y)
{
    if (hs == 2 && vs == 2) return jpg_upsample_h2v2(c, cs, cw, ch, x, y);
    if (hs == 2) return jpg_upsample_h2v1(c, cs, cw, x, y);
    return c[y * cs + x];
}

TRS_JHD void jpg_ycc_to_rgb(int y, int cb, int cr, uint8_t& r, uint8_t& g, uint8_t& b)
{
    const int SB = 16, HALF = 1 << (SB - 1);
    const int xb = cb - 128, xr = cr - 128;
    const int cr_r = (91881 * xr + HALF) >> SB;               // FIX(1.40200)
    const int cb_b = (116130 * xb + HALF) >> SB;              // FIX(1.77200)
    const int cr_g = -46802 * xr;                             // -FIX(0.71414)
    const int cb_g = -22554 * xb + HALF;                      // -FIX(0.34414), rounding folded in
    int R = y + cr_r, G = y + ((cb_g + cr_g) >> SB), B = y + cb_b;
    r = (uint8_t)(R < 0 ? 0 : (R > 255 ? 255 : R));
    g = (uint8_t)(G < 0 ? 0 : (G > 255 ? 255 : G));
    b = (uint8_t)(B < 0 ? 0 : (B > 255 ? 255 : B));
}

}  // namespace trs
