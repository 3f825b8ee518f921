// jpeg_host.h — host-side parsing of a tub record's JPEG file into the tables and the entropy-coded segment the kernels need.
// Accepts what the reference's recorder writes (components/datastorage.py:78, Pillow defaults) and its close relatives: baseline
// sequential DCT (SOF0), 8-bit samples, three components (ids 1, 2, 3 = Y, Cb, Cr) with luma sampled 2x2 (4:2:0), 2x1 (4:2:2) or
// 1x1 (4:4:4) and chroma 1x1, one interleaved scan, no restart intervals.  Anything else is reported as unsupported (the caller
// raises; there is no CPU decoder to fall back to).
#pragma once
#include <stdint.h>
#include <string.h>

#include "jpeg_core.cuh"

namespace trs {

enum { JPG_E_FORMAT = 10, JPG_E_UNSUPPORTED = 11 };

struct JpegScan {
    int h, w;
    int hs, vs;                       // luma sampling factors: (2,2), (2,1) or (1,1)
    uint32_t data_off, data_len;      // entropy-coded segment: from the byte after the SOS header to the end of the file
};

// jdhuff.c jpeg_make_d_derived_tbl: code lengths -> canonical codes -> look-ahead + maxcode / valoffset
inline int jpg_build_huff(const uint8_t* bits /*[16]*/, const uint8_t* vals, int nvals, JpegHuff* h)
{
    memset(h, 0, sizeof *h);
    uint8_t huffsize[257];
    uint32_t huffcode[257];
    int p = 0;
    for (int l = 1; l <= 16; ++l)
        for (int i = 0; i < bits[l - 1]; ++i) {
            if (p >= 256) return JPG_E_FORMAT;
            huffsize[p++] = (uint8_t)l;
        }
    if (p != nvals) return JPG_E_FORMAT;
    huffsize[p] = 0;
    uint32_t code = 0;
    int si = huffsize[0];
    for (int k = 0; huffsize[k];) {
        while (huffsize[k] == si) huffcode[k++] = code++;
        if (code > (1u << si)) return JPG_E_FORMAT;
        code <<= 1;
        ++si;
    }
    for (int i = 0; i < nvals; ++i) h->huffval[i] = vals[i];
    p = 0;
    for (int l = 1; l <= 16; ++l) {
        if (bits[l - 1]) {
            h->valoffset[l] = p - (int32_t)huffcode[p];
            p += bits[l - 1];
            h->maxcode[l] = (int32_t)huffcode[p - 1];
        } else {
            h->maxcode[l] = -1;
        }
    }
    h->maxcode[17] = 0xFFFFF;
    p = 0;
    for (int l = 1; l <= 8; ++l)
        for (int i = 0; i < bits[l - 1]; ++i, ++p) {
            const uint32_t look = huffcode[p] << (8 - l);
            for (int c = 0; c < (1 << (8 - l)); ++c) {
                h->look_nbits[look + c] = (uint8_t)l;
                h->look_sym[look + c] = vals[p];
            }
        }
    return 0;
}

inline int jpg_parse(const uint8_t* f, size_t len, JpegTables* T, JpegScan* S)
{
    if (len < 4 || f[0] != 0xff || f[1] != 0xd8) return JPG_E_FORMAT;
    uint16_t qt[4][64];
    bool have_qt[4] = {false, false, false, false};
    JpegHuff hd[4], ha[4];
    bool have_hd[4] = {false, false, false, false}, have_ha[4] = {false, false, false, false};
    int comp_tq[3] = {-1, -1, -1};
    bool have_sof = false;
    size_t i = 2;
    while (i + 4 <= len) {
        if (f[i] != 0xff) return JPG_E_FORMAT;
        const int m = f[i + 1];
        if (m == 0xff) { ++i; continue; }                       // fill byte
        const size_t L = ((size_t)f[i + 2] << 8) | f[i + 3];
        if (L < 2 || i + 2 + L > len) return JPG_E_FORMAT;
        const uint8_t* seg = f + i + 4;
        const size_t sl = L - 2;
        if (m == 0xdb) {                                         // DQT
            size_t o = 0;
            while (o < sl) {
                const int pq = seg[o] >> 4, tq = seg[o] & 15;
                if (pq != 0 || tq > 3 || o + 65 > sl) return JPG_E_UNSUPPORTED;
                for (int k = 0; k < 64; ++k) qt[tq][jpg_natural_order_host(k)] = seg[o + 1 + k];
                have_qt[tq] = true;
                o += 65;
            }
        } else if (m == 0xc4) {                                  // DHT
            size_t o = 0;
            while (o + 17 <= sl) {
                const int tc = seg[o] >> 4, th = seg[o] & 15;
                if (tc > 1 || th > 3) return JPG_E_FORMAT;
                int n = 0;
                for (int k = 0; k < 16; ++k) n += seg[o + 1 + k];
                if (n > 256 || o + 17 + n > sl) return JPG_E_FORMAT;
                const int rc = jpg_build_huff(seg + o + 1, seg + o + 17, n, tc ? &ha[th] : &hd[th]);
                if (rc) return rc;
                (tc ? have_ha : have_hd)[th] = true;
                o += 17 + n;
            }
        } else if (m == 0xc0) {                                  // SOF0: baseline
            if (sl < 15 || seg[0] != 8 || seg[5] != 3) return JPG_E_UNSUPPORTED;
            S->h = (seg[1] << 8) | seg[2];
            S->w = (seg[3] << 8) | seg[4];
            const int hv0 = seg[7];
            if (hv0 != 0x22 && hv0 != 0x21 && hv0 != 0x11) return JPG_E_UNSUPPORTED;
            S->hs = hv0 >> 4;
            S->vs = hv0 & 15;
            for (int c = 0; c < 3; ++c) {
                if (seg[6 + 3 * c] != c + 1 || (c > 0 && seg[7 + 3 * c] != 0x11) || seg[8 + 3 * c] > 3) return JPG_E_UNSUPPORTED;
                comp_tq[c] = seg[8 + 3 * c];
            }
            if (comp_tq[1] != comp_tq[2]) return JPG_E_UNSUPPORTED;
            have_sof = true;
        } else if (m >= 0xc1 && m <= 0xcf && m != 0xc4 && m != 0xc8 && m != 0xcc) {
            return JPG_E_UNSUPPORTED;                            // progressive, extended, lossless, arithmetic ...
        } else if (m == 0xdd) {                                  // DRI
            if (sl >= 2 && ((seg[0] << 8) | seg[1]) != 0) return JPG_E_UNSUPPORTED;
        } else if (m == 0xda) {                                  // SOS
            if (!have_sof || sl < 10 || seg[0] != 3) return JPG_E_UNSUPPORTED;
            int td[3], ta[3];
            for (int c = 0; c < 3; ++c) {
                if (seg[1 + 2 * c] != c + 1) return JPG_E_UNSUPPORTED;
                td[c] = seg[2 + 2 * c] >> 4;
                ta[c] = seg[2 + 2 * c] & 15;
                if (td[c] > 3 || ta[c] > 3 || !have_hd[td[c]] || !have_ha[ta[c]]) return JPG_E_FORMAT;
            }
            if (td[1] != td[2] || ta[1] != ta[2] || seg[7] != 0 || seg[8] != 63 || seg[9] != 0) return JPG_E_UNSUPPORTED;
            if (!have_qt[comp_tq[0]] || !have_qt[comp_tq[1]]) return JPG_E_FORMAT;
            memcpy(T->quant[0], qt[comp_tq[0]], sizeof qt[0]);
            memcpy(T->quant[1], qt[comp_tq[1]], sizeof qt[0]);
            T->dc[0] = hd[td[0]]; T->ac[0] = ha[ta[0]];
            T->dc[1] = hd[td[1]]; T->ac[1] = ha[ta[1]];
            S->data_off = (uint32_t)(i + 2 + L);
            S->data_len = (uint32_t)(len - S->data_off);
            return 0;
        }
        i += 2 + L;
    }
    return JPG_E_FORMAT;
}

}  // namespace trs
