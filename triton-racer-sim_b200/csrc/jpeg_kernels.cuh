// jpeg_kernels.cuh — batched baseline-JPEG decode of tub records on the GPU (sm_100a); arithmetic in jpeg_core.cuh.
//
//   k_jpeg_entropy_idct   one THREAD per record: Huffman decoding is sequential within a scan (no restart markers in the
//                         recorder's files), so the parallelism is across the N records of the batch.  Each thread walks its
//                         record's MCUs, decodes a block into a private coefficient array, runs the integer IDCT and writes
//                         the 8x8 samples into planar Y / Cb / Cr buffers (MCU-padded).  Tables live in shared memory.
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
    uint8_t* y;             // per record: (mh*16) x (mw*16)
    uint8_t* cb;            // per record: (mh*8) x (mw*8)
    uint8_t* cr;
    int mw, mh;
};

enum { JPG_THREADS = 64 };

__global__ void __launch_bounds__(JPG_THREADS) k_jpeg_entropy_idct(const uint8_t* __restrict__ blob, const JpegRecord* __restrict__ recs,
                                                                  const JpegTables* __restrict__ tables, int n, JpegPlanes P,
                                                                  int* __restrict__ status)
{
    __shared__ JpegTables s_tab;                                  // table set 0 (a tub written by one recorder has one set)
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(tables);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&s_tab);
        for (int i = threadIdx.x; i < (int)(sizeof(JpegTables) / 4); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const JpegRecord rec = recs[r];
    const JpegTables& T = rec.table_set == 0 ? s_tab : tables[rec.table_set];
    const int ys = P.mw * 16, cs = P.mw * 8;
    uint8_t* Y = P.y + (size_t)r * ys * P.mh * 16;
    uint8_t* Cb = P.cb + (size_t)r * cs * P.mh * 8;
    uint8_t* Cr = P.cr + (size_t)r * cs * P.mh * 8;
    JpegBits b{blob + rec.data_off, blob + rec.data_off + rec.data_len, 0, 0};
    int dc[3] = {0, 0, 0}, err = 0;
    int16_t coef[64];
    for (int my = 0; my < P.mh; ++my)
        for (int mx = 0; mx < P.mw; ++mx) {
#pragma unroll 1
            for (int k = 0; k < 4; ++k) {                          // Y blocks of the MCU in raster order
                jpg_decode_block(b, T.dc[0], T.ac[0], dc[0], coef, err);
                jpg_idct_islow(coef, T.quant[0], Y + (my * 16 + (k >> 1) * 8) * ys + mx * 16 + (k & 1) * 8, ys);
            }
            jpg_decode_block(b, T.dc[1], T.ac[1], dc[1], coef, err);
            jpg_idct_islow(coef, T.quant[1], Cb + my * 8 * cs + mx * 8, cs);
            jpg_decode_block(b, T.dc[1], T.ac[1], dc[2], coef, err);
            jpg_idct_islow(coef, T.quant[1], Cr + my * 8 * cs + mx * 8, cs);
        }
    if (err) atomicMax(status, err);
}

__global__ void __launch_bounds__(256) k_jpeg_upsample_rgb(JpegPlanes P, int n, int h, int w, uint8_t* __restrict__ out)
{
    const int ys = P.mw * 16, cs = P.mw * 8;
    const int cw = (w + 1) / 2, ch = (h + 1) / 2;
    const int gpr = (w + 3) / 4;                                   // groups of 4 pixels per row
    const size_t total = (size_t)n * h * gpr;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const bool words = (w & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int g = (int)(i % gpr);
        const size_t t = i / gpr;
        const int y = (int)(t % h);
        const size_t r = t / h;
        const uint8_t* Y = P.y + r * (size_t)ys * P.mh * 16 + (size_t)y * ys;
        const uint8_t* Cb = P.cb + r * (size_t)cs * P.mh * 8;
        const uint8_t* Cr = P.cr + r * (size_t)cs * P.mh * 8;
        uint8_t px[12];
        const int x0 = 4 * g;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int x = min(x0 + q, w - 1);
            jpg_ycc_to_rgb(Y[x], jpg_upsample_h2v2(Cb, cs, cw, ch, x, y), jpg_upsample_h2v2(Cr, cs, cw, ch, x, y), px[3 * q], px[3 * q + 1], px[3 * q + 2]);
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
