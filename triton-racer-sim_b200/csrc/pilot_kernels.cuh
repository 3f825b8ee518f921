// pilot_kernels.cuh — forward pass of the reference's pilots for N frames on sm_100a.
//
// Reference: Keras_2D_CNN.get_model / Keras_2D_FULL_HOUSE.get_model (TritonRacerSim/components/keras_train.py:127-174,184-245):
// seven VALID convolutions with ReLU (5x5/2 x3, 3x3/1 x4; 24-32-64-64-64-128-128 filters), Flatten, and a few small Dense layers;
// called per frame from KerasPilot.step (components/keras_pilot.py:59,71,81,104).  Dropout is the identity at inference.
//
// This is the one dense contraction of the path, so it runs on the 5th-generation tensor cores:
//   * every convolution is an implicit GEMM  C[pixel, filter] = sum_k A[pixel, k] * B[filter, k]  with k = (kh, kw, channel);
//     in NHWC the kw*C values of one kernel row are CONTIGUOUS in the input, so the A operand needs no im2col buffer: a 5-D
//     TMA tensor map with overlapping strides (run element, output x, kernel row, output y, frame) gathers a
//     [<=128 pixels] x [64 k-values] tile straight from the activation tensor into 128-byte-swizzled shared memory;
//   * `tcgen05.mma.cta_group::1.kind::f16` (fp16 operands, fp32 accumulation in TMEM) issued by one elected thread;
//   * warp roles: warp 0 TMA producer, warp 1 TMEM allocation + MMA issue, warps 2..5 epilogue (tcgen05.ld -> bias + ReLU -> fp16 NHWC
//     stores, which IS the next layer's input layout); full/empty mbarrier ring between producer and MMA, tcgen05.commit
//     frees the stages and signals the epilogue;
//   * fp16 storage has the 10-bit mantissa of the TF32 arithmetic TensorFlow uses for these layers on a GPU; accumulation is fp32.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace trs {
namespace pilot {

constexpr int BLOCK_M = 128;        // accumulator rows = TMEM lanes
constexpr int BLOCK_K = 64;         // fp16: 128 bytes per operand row = one 128-byte swizzle atom
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 192;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;

// One implicit-GEMM launch.  A tile of the M dimension is a box (bx output columns, by output rows, bn frames).
struct GemmGeom {
    int bx, by, bn;          // box extents; bx*by*bn <= 128 rows
    int wo, ho, nf;          // output width / height, frames in this launch
    int x_tiles, y_tiles;    // ceil(wo/bx), ceil(ho/by)
    int kchunks;             // 64-element chunks per kernel row
    int nkb;                 // K blocks = kernel rows * kchunks
    int n_valid;             // output columns that exist (<= NPAD)
    int ldc;                 // output row pitch in elements
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// A wait that cannot hang the GPU: a barrier that does not flip within ~2 s of SM clocks is a protocol bug -> trap.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    if (mbar_try(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try(bar, parity))
        if (clock64() - t0 > 4000000000LL) __trap();
}

__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4)
{
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}

// Shared-memory matrix descriptor, K-major operand rows of 128 bytes, SWIZZLE_128B: 8-row groups are 1024 bytes apart.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);   // start address, 16-byte units
    d |= (uint64_t)1 << 16;                      // leading byte offset: unused for swizzled K-major operands
    d |= (uint64_t)(1024 >> 4) << 32;            // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                      // descriptor version of sm_100
    d |= (uint64_t)2 << 61;                      // SWIZZLE_128B
    return d;
}

// kind::f16 instruction descriptor: fp16 x fp16 -> fp32, both operands K-major, M = 128.
__host__ __device__ constexpr uint32_t umma_idesc_f16(int n)
{
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}

constexpr int gemm_smem_bytes(int npad, int stages) { return stages * (A_STAGE_BYTES + npad * BLOCK_K * 2) + 1024; }

// C[tile rows, NPAD] = A-box . B^T (+ bias, ReLU, fp16) or raw fp32 partial sums (OUT_F32: the flattened features times the first
// Dense layers of the heads; bias and activation are applied by k_pilot_heads together with the feature branches).
template <int NPAD, int STAGES, bool OUT_F32>
__global__ void __launch_bounds__(GEMM_THREADS) k_pilot_gemm(const __grid_constant__ CUtensorMap map_a,
                                                             const __grid_constant__ CUtensorMap map_b, const GemmGeom g,
                                                             const float* __restrict__ bias, void* __restrict__ out)
{
    constexpr int B_STAGE_BYTES = NPAD * BLOCK_K * 2;
    constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    constexpr int TMEM_COLS = NPAD < 32 ? 32 : NPAD;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[STAGES];
    __shared__ __align__(8) uint64_t bar_empty[STAGES];
    __shared__ __align__(8) uint64_t bar_acc;
    __shared__ uint32_t tmem_slot;

    const uint32_t tiles = (smem_u32(smem_raw) + 1023u) & ~1023u;      // swizzle atoms are 1024-byte aligned
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // tile -> box origin
    int t = blockIdx.x;
    const int xt = t % g.x_tiles;  t /= g.x_tiles;
    const int yt = t % g.y_tiles;  t /= g.y_tiles;
    const int x0 = xt * g.bx, y0 = yt * g.by, n0 = t * g.bn;
    const uint32_t rows_box = (uint32_t)(g.bx * g.by * g.bn);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_b) : "memory");
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) {
                mbar_init(smem_u32(&bar_full[s]), 1);
                mbar_init(smem_u32(&bar_empty[s]), 1);
            }
            mbar_init(smem_u32(&bar_acc), 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < g.nkb; ++kb) {
                const int s = kb % STAGES, round = kb / STAGES;
                if (round > 0) mbar_wait(smem_u32(&bar_empty[s]), (uint32_t)(round - 1) & 1u);
                const uint32_t full = smem_u32(&bar_full[s]);
                mbar_expect_tx(full, rows_box * (BLOCK_K * 2) + B_STAGE_BYTES);
                const int kr = kb / g.kchunks, ch = kb - kr * g.kchunks;
                const uint32_t a_dst = tiles + (uint32_t)s * STAGE_BYTES;
                tma_load_5d(a_dst, &map_a, full, ch * BLOCK_K, x0, kr, y0, n0);
                tma_load_2d(a_dst + A_STAGE_BYTES, &map_b, full, kb * BLOCK_K, 0);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(NPAD);
            for (int kb = 0; kb < g.nkb; ++kb) {
                const int s = kb % STAGES, round = kb / STAGES;
                mbar_wait(smem_u32(&bar_full[s]), (uint32_t)round & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_src = tiles + (uint32_t)s * STAGE_BYTES;
                const uint64_t da = umma_desc_sw128(a_src), db = umma_desc_sw128(a_src + A_STAGE_BYTES);
#pragma unroll
                for (int k = 0; k < BLOCK_K / UMMA_K; ++k)      // +32 bytes inside the swizzle atom per K step
                    umma_f16(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (uint32_t)((kb | k) != 0));
                umma_commit(smem_u32(&bar_empty[s]));           // stage free once these MMAs have read it
            }
            umma_commit(smem_u32(&bar_acc));                    // accumulator complete
        }
        __syncwarp();
    } else {
        // epilogue: warp w may touch TMEM lanes 32*(w%4) .. +31; thread = one accumulator row = one output pixel
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const int x = x0 + r % g.bx;
        const int t2 = r / g.bx;
        const int y = y0 + t2 % g.by;
        const int n = n0 + t2 / g.by;
        const bool live = (uint32_t)r < rows_box && x < g.wo && y < g.ho && n < g.nf;
        const size_t row = ((size_t)n * g.ho + y) * g.wo + x;
        mbar_wait(smem_u32(&bar_acc), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
        for (int c = 0; c < NPAD; c += 16) {
            if (c >= g.n_valid) break;                          // warp-uniform
            uint32_t v[16];
            tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (!live) continue;
            if (OUT_F32) {
                float4* dst = reinterpret_cast<float4*>(static_cast<float*>(out) + row * g.ldc + c);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (c + 4 * j < g.n_valid)
                        dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                             __uint_as_float(v[4 * j + 3]));
            } else {
                uint4* dst = reinterpret_cast<uint4*>(static_cast<__half*>(out) + row * g.ldc + c);
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    if (c + 8 * j >= g.n_valid) break;
                    uint32_t p[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int col = c + 8 * j + 2 * e;
                        const float a = fmaxf(__uint_as_float(v[8 * j + 2 * e]) + __ldg(bias + col), 0.0f);
                        const float b = fmaxf(__uint_as_float(v[8 * j + 2 * e + 1]) + __ldg(bias + col + 1), 0.0f);
                        const __half2 h = __floats2half2_rn(a, b);
                        p[e] = *reinterpret_cast<const uint32_t*>(&h);
                    }
                    dst[j] = make_uint4(p[0], p[1], p[2], p[3]);
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// (N,H,W,3) u8 -> (N,H,W,4) fp16 = x / 255 (keras_pilot.py:49-50), fourth channel zero: 8 bytes per pixel, so a pixel step of the
// stride-2 first convolution is 16 bytes, the granularity a TMA stride needs.  Four pixels (12 bytes in, 32 bytes out) per thread.
__global__ void __launch_bounds__(256) k_pilot_input(const uint8_t* __restrict__ in, __half* __restrict__ out, size_t n_quads)
{
    __shared__ __half lut[256];
    lut[threadIdx.x] = __float2half_rn(__fdiv_rn((float)threadIdx.x, 255.0f));
    __syncthreads();
    const uint32_t* in32 = reinterpret_cast<const uint32_t*>(in);
    uint4* out16 = reinterpret_cast<uint4*>(out);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_quads; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t w0 = __ldg(in32 + 3 * i), w1 = __ldg(in32 + 3 * i + 1), w2 = __ldg(in32 + 3 * i + 2);
        const uint8_t b[12] = {(uint8_t)w0, (uint8_t)(w0 >> 8), (uint8_t)(w0 >> 16), (uint8_t)(w0 >> 24),
                               (uint8_t)w1, (uint8_t)(w1 >> 8), (uint8_t)(w1 >> 16), (uint8_t)(w1 >> 24),
                               (uint8_t)w2, (uint8_t)(w2 >> 8), (uint8_t)(w2 >> 16), (uint8_t)(w2 >> 24)};
        uint32_t o[8];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const uint32_t r = __half_as_ushort(lut[b[3 * p]]), gch = __half_as_ushort(lut[b[3 * p + 1]]);
            const uint32_t bl = __half_as_ushort(lut[b[3 * p + 2]]);
            o[2 * p] = r | (gch << 16);
            o[2 * p + 1] = bl;
        }
        out16[2 * i] = make_uint4(o[0], o[1], o[2], o[3]);
        out16[2 * i + 1] = make_uint4(o[4], o[5], o[6], o[7]);
    }
}

// The small Dense layers after Flatten.  One "head" = [feature branch 1 -> w1 -> w2 -> w3 (ReLU each)] ++ flattened image features
// -> 100 -> 50 -> 25 (ReLU) -> n_out (linear); Keras_2D_CNN has one head with two outputs (keras_train.py:149-166), the full-house
// model two heads sharing the image features (keras_train.py:206-236).  The products of the image features with the first Dense
// layer come from k_pilot_gemm (fp32, no bias); everything else is a few thousand multiply-adds per frame: one warp per frame,
// lanes over output units, weights (Keras [in, out] layout: consecutive lanes read consecutive words) in shared memory.
struct HeadDesc {
    int fw[3];          // feature branch widths (0: no feature input)
    int n_prev;         // features carried over from head 0's branch: the full-house steering head sees Concatenate([x ++ y, s])
    int n_out;          // 1 or 2
    int out_slot;       // first column of the (N,2) output this head writes
    int part_col;       // first column of this head's 100 partial sums in the GEMM output
    // offsets (in floats) into the weight blob
    int f_w[3], f_b[3]; // feature Dense kernels / biases
    int d1y_w, d1_b;    // rows of the first Dense kernel that multiply the feature branch ([w3, 100]); its bias
    int d2_w, d2_b, d3_w, d3_b, o_w, o_b;
};
struct HeadsArgs {
    HeadDesc head[2];
    int n_heads;
    int blob_floats;
    int ldp;            // pitch of the partial-sum matrix
};

__device__ __forceinline__ void dense_warp(const float* __restrict__ w, const float* __restrict__ b, const float* __restrict__ in, int n_in,
                                           float* __restrict__ outv, int n_out, bool relu, int lane)
{
    for (int j = lane; j < n_out; j += 32) {
        float acc = b[j];
        for (int k = 0; k < n_in; ++k) acc = fmaf(in[k], w[k * n_out + j], acc);
        outv[j] = relu ? fmaxf(acc, 0.0f) : acc;
    }
    __syncwarp();
}

constexpr int HEADS_WARPS = 16;
constexpr int HEADS_SCRATCH = 320;

__global__ void __launch_bounds__(HEADS_WARPS * 32) k_pilot_heads(const HeadsArgs a, const float* __restrict__ blob,
                                                                  const float* __restrict__ partial, const float* __restrict__ feat0,
                                                                  const float* __restrict__ feat1, float* __restrict__ out, int n)
{
    extern __shared__ float sm[];
    float* w = sm;                                                   // weight blob
    float* scratch = sm + a.blob_floats;                             // 2 x 128 + 64 floats per warp
    for (int i = threadIdx.x; i < a.blob_floats; i += blockDim.x) w[i] = __ldg(blob + i);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* p = scratch + warp * HEADS_SCRATCH;
    float* q = p + 128;
    float* keep = q + 128;                                           // head 0's feature branch output
    for (int f = blockIdx.x * HEADS_WARPS + warp; f < n; f += gridDim.x * HEADS_WARPS) {
        for (int h = 0; h < a.n_heads; ++h) {
            const HeadDesc& d = a.head[h];
            const float* feat = h == 0 ? feat0 : feat1;
            int n_y = 0;
            if (d.fw[0] > 0) {
                if (lane == 0) q[0] = __ldg(feat + f);
                __syncwarp();
                dense_warp(w + d.f_w[0], w + d.f_b[0], q, 1, p, d.fw[0], true, lane);
                dense_warp(w + d.f_w[1], w + d.f_b[1], p, d.fw[0], q, d.fw[1], true, lane);
                dense_warp(w + d.f_w[2], w + d.f_b[2], q, d.fw[1], p, d.fw[2], true, lane);
                n_y = d.fw[2];
                if (h == 0 && a.n_heads > 1) {
                    for (int k = lane; k < n_y; k += 32) keep[k] = p[k];
                    __syncwarp();
                }
            }
            // first Dense layer: image part from the tensor cores + feature part + bias
            for (int j = lane; j < 100; j += 32) {
                float acc = __ldg(partial + (size_t)f * a.ldp + d.part_col + j) + w[d.d1_b + j];
                for (int k = 0; k < d.n_prev; ++k) acc = fmaf(keep[k], w[d.d1y_w + k * 100 + j], acc);
                for (int k = 0; k < n_y; ++k) acc = fmaf(p[k], w[d.d1y_w + (d.n_prev + k) * 100 + j], acc);
                q[j] = fmaxf(acc, 0.0f);
            }
            __syncwarp();
            dense_warp(w + d.d2_w, w + d.d2_b, q, 100, p, 50, true, lane);
            dense_warp(w + d.d3_w, w + d.d3_b, p, 50, q, 25, true, lane);
            dense_warp(w + d.o_w, w + d.o_b, q, 25, p, d.n_out, false, lane);
            if (lane < d.n_out) out[(size_t)f * 2 + d.out_slot + lane] = p[lane];
            __syncwarp();
        }
    }
}

}  // namespace pilot
}  // namespace trs
