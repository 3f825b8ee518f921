// Host build of the product's JPEG arithmetic (csrc/jpeg_core.cuh + csrc/jpeg_host.h are pure host/device functions) for the CPU-side
// parity test against Pillow (tests/test_jpeg_host.py).  Test infrastructure: mirrors the record loop of k_jpeg_entropy_idct /
// k_jpeg_upsample_rgb one record at a time.
#include <stdlib.h>

#include "../triton-racer-sim_b200/csrc/jpeg_core.cuh"
#include "../triton-racer-sim_b200/csrc/jpeg_host.h"
using namespace trs;

extern "C" int jpg_decode_rgb_host(const uint8_t* file, size_t len, int h, int w, uint8_t* rgb)
{
    JpegTables T;
    JpegScan S;
    int rc = jpg_parse(file, len, &T, &S);
    if (rc) return 100 + rc;
    if (S.h != h || S.w != w) return 99;
    const int hs = S.hs, vs = S.vs, lb = hs * vs;
    const int mw = (w + 8 * hs - 1) / (8 * hs), mh = (h + 8 * vs - 1) / (8 * vs), ys = mw * 8 * hs, cs = mw * 8;
    uint8_t* Y = (uint8_t*)malloc((size_t)ys * mh * 8 * vs);
    uint8_t* Cb = (uint8_t*)malloc((size_t)cs * mh * 8);
    uint8_t* Cr = (uint8_t*)malloc((size_t)cs * mh * 8);
    JpegBits b{file + S.data_off, file + S.data_off + S.data_len, 0, 0};
    int dc[3] = {0, 0, 0}, err = 0;
    int16_t coef[64];
    for (int my = 0; my < mh; ++my)
        for (int mx = 0; mx < mw; ++mx) {
            for (int k = 0; k < lb; ++k) {
                jpg_decode_block(b, T.dc[0], T.ac[0], dc[0], coef, err);
                jpg_idct_islow(coef, T.quant[0], Y + ((my * vs + k / hs) * 8) * ys + (mx * hs + k % hs) * 8, ys);
            }
            jpg_decode_block(b, T.dc[1], T.ac[1], dc[1], coef, err);
            jpg_idct_islow(coef, T.quant[1], Cb + my * 8 * cs + mx * 8, cs);
            jpg_decode_block(b, T.dc[1], T.ac[1], dc[2], coef, err);
            jpg_idct_islow(coef, T.quant[1], Cr + my * 8 * cs + mx * 8, cs);
        }
    const int cw = (w + hs - 1) / hs, ch = (h + vs - 1) / vs;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            uint8_t* o = rgb + ((size_t)y * w + x) * 3;
            jpg_ycc_to_rgb(Y[y * ys + x], jpg_chroma_at(Cb, cs, cw, ch, hs, vs, x, y), jpg_chroma_at(Cr, cs, cw, ch, hs, vs, x, y), o[0], o[1], o[2]);
        }
    free(Y); free(Cb); free(Cr);
    return err;
}
