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
    int tiles;               // x_tiles * y_tiles * ceil(nf / bn)
    int last_steps;          // MMAs (K = 16 each) the last chunk of a kernel row needs: the rest of its 64 columns is padding
};

// Walks the tiles c, c + G, c + 2G, ... of a persistent CTA without dividing per tile: (x tile, y tile, frame tile) with carries.
struct TileWalk {
    int tile, xt, yt, nt;
    int gx, gy, gn, stride;
    __device__ __forceinline__ void init(const GemmGeom& g, int first, int step)
    {
        tile = first; stride = step;
        int t = first;
        xt = t % g.x_tiles;  t /= g.x_tiles;
        yt = t % g.y_tiles;  nt = t / g.y_tiles;
        t = step;
        gx = t % g.x_tiles;  t /= g.x_tiles;
        gy = t % g.y_tiles;  gn = t / g.y_tiles;
    }
    __device__ __forceinline__ void next(const GemmGeom& g)
    {
        tile += stride;
        xt += gx;
        if (xt >= g.x_tiles) { xt -= g.x_tiles; ++yt; }
        yt += gy;
        if (yt >= g.y_tiles) { yt -= g.y_tiles; ++nt; }
        nt += gn;
    }
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
// keep a loop-invariant value in a register: without this the compiler rematerialises shared-window addresses (S2UR SR_CgaCtaId + ULEA,
// ~30 cycles of latency each) and kernel parameters (LDC) inside the single-thread issue loops, whose cost is pure dependent latency
__device__ __forceinline__ uint32_t keep(uint32_t x) { asm volatile("" : "+r"(x)); return x; }
__device__ __forceinline__ int keep(int x) { asm volatile("" : "+r"(x)); return x; }
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity)      // non-blocking probe
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}

// A wait that cannot hang the GPU: a barrier that has not flipped after 2^22 polls (each poll already blocks for the hardware's
// try_wait time limit) is a protocol bug -> trap instead of spinning forever.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t spins = 0;
    while (!mbar_try(bar, parity))
        if (++spins == (1u << 22)) __trap();
}

// The same for warps that are not on the critical path (epilogue warps waiting for an accumulator, the MMA thread of conv1 waiting
// for the builders): back off between polls so that the spin does not take issue slots from the warps doing the work.
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity)
{
    uint32_t spins = 0;
    while (!mbar_try(bar, parity)) {
        __nanosleep(40);
        if (++spins == (1u << 22)) __trap();
    }
}

// One lane of a converged warp.  Code under `if (elect_one())` is known to the compiler to run in a single lane, so the uniform-datapath
// instructions in it (UTCHMMA, UTMALDG, UTCBAR) are emitted once; under `if (lane == 0)` each of them is wrapped in an
// ELECT / PLOP3 / BRA.U.ANY loop over the "active lanes", which made the MMA-issuing thread the bottleneck of every layer
// (profiles/r02_pilot.md: 134 instructions per four MMAs).
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4)
{
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
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

// A operand from tensor memory (row m = TMEM lane m, two fp16 k-values per 32-bit column): no shared-memory read for A.
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}

constexpr int gemm_smem_bytes(int npad, int stages) { return stages * (A_STAGE_BYTES + npad * BLOCK_K * 2) + 1024; }

__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// C[tile rows, NPAD] = A-box . B^T (+ bias, ReLU, fp16) or raw fp32 partial sums (OUT_F32: the flattened features times the first
// Dense layers of the heads; bias and activation are applied by k_pilot_heads together with the feature branches).
// Persistent: CTA c takes tiles c, c + gridDim.x, ...; the producer runs ahead across tile boundaries through the stage ring, and
// the accumulator is double-buffered in TMEM (2 x NPAD columns) so the epilogue of one tile overlaps the MMAs of the next.
template <int NPAD, int STAGES, bool OUT_F32>
__global__ void __launch_bounds__(GEMM_THREADS) k_pilot_gemm(const __grid_constant__ CUtensorMap map_a,
                                                             const __grid_constant__ CUtensorMap map_b, const GemmGeom g,
                                                             const float* __restrict__ bias, void* __restrict__ out)
{
    constexpr int B_STAGE_BYTES = NPAD * BLOCK_K * 2;
    constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    constexpr int TMEM_COLS = 2 * NPAD;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[STAGES];
    __shared__ __align__(8) uint64_t bar_empty[STAGES];
    __shared__ __align__(8) uint64_t bar_acc_full[2];
    __shared__ __align__(8) uint64_t bar_acc_empty[2];
    __shared__ uint32_t tmem_slot;
    __shared__ float bias_s[NPAD];

    const uint32_t tiles = (smem_u32(smem_raw) + 1023u) & ~1023u;      // swizzle atoms are 1024-byte aligned
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rows_box = (uint32_t)(g.bx * g.by * g.bn);

    if (!OUT_F32)
        for (int i = threadIdx.x; i < NPAD; i += GEMM_THREADS) bias_s[i] = __ldg(bias + i);
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
            for (int a = 0; a < 2; ++a) {
                mbar_init(smem_u32(&bar_acc_full[a]), 1);
                mbar_init(smem_u32(&bar_acc_empty[a]), 4);      // one arrival per epilogue warp
            }
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
        if (elect_one()) {
            const uint32_t full0 = smem_u32(&bar_full[0]), empty0 = smem_u32(&bar_empty[0]);
            uint32_t s = 0, ph = 0;                                  // ring position; ph = parity of the stage's NEXT completion
            TileWalk tw;
            for (tw.init(g, blockIdx.x, gridDim.x); tw.tile < g.tiles; tw.next(g)) {
                const int x0 = tw.xt * g.bx, y0 = tw.yt * g.by, n0 = tw.nt * g.bn;
                int kr = 0, ch = 0;
                for (int kb = 0; kb < g.nkb; ++kb) {
                    mbar_wait(empty0 + 8 * s, ph ^ 1u);              // (a fresh barrier passes a wait on parity 1)
                    const uint32_t full = full0 + 8 * s;
                    mbar_expect_tx(full, rows_box * (BLOCK_K * 2) + B_STAGE_BYTES);
                    const uint32_t a_dst = tiles + s * STAGE_BYTES;
                    tma_load_5d(a_dst, &map_a, full, ch * BLOCK_K, x0, kr, y0, n0);
                    tma_load_2d(a_dst + A_STAGE_BYTES, &map_b, full, kb * BLOCK_K, 0);
                    if (++ch == g.kchunks) { ch = 0; ++kr; }
                    if (++s == STAGES) { s = 0; ph ^= 1u; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_f16(NPAD);
            const uint32_t full0 = smem_u32(&bar_full[0]), empty0 = smem_u32(&bar_empty[0]);
            const uint32_t accf0 = smem_u32(&bar_acc_full[0]), acce0 = smem_u32(&bar_acc_empty[0]);
            const uint64_t da0 = umma_desc_sw128(tiles), db0 = umma_desc_sw128(tiles + A_STAGE_BYTES);
            uint32_t s = 0, ph = 0, ti = 0;
            for (int tile = blockIdx.x; tile < g.tiles; tile += gridDim.x, ++ti) {
                const uint32_t acc = ti & 1u, use = ti >> 1;
                mbar_wait(acce0 + 8 * acc, (use & 1u) ^ 1u);         // epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem_base + acc * NPAD;
                int ch = 0;
                for (int kb = 0; kb < g.nkb; ++kb) {
                    mbar_wait(full0 + 8 * s, ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint64_t da = da0 + (uint64_t)(s * (STAGE_BYTES >> 4)), db = db0 + (uint64_t)(s * (STAGE_BYTES >> 4));
                    const bool last = ++ch == g.kchunks;            // the last chunk of a kernel row may be mostly padding
                    if (last) ch = 0;
                    umma_f16(d_tmem, da, db, idesc, (uint32_t)(kb != 0));
                    if (!last || g.last_steps == BLOCK_K / UMMA_K) {   // +32 bytes inside the swizzle atom per K step
                        umma_f16(d_tmem, da + 2u, db + 2u, idesc, 1u);
                        umma_f16(d_tmem, da + 4u, db + 4u, idesc, 1u);
                        umma_f16(d_tmem, da + 6u, db + 6u, idesc, 1u);
                    } else {
                        for (int k = 1; k < g.last_steps; ++k) umma_f16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, 1u);
                    }
                    umma_commit(empty0 + 8 * s);                    // stage free once these MMAs have read it
                    if (++s == STAGES) { s = 0; ph ^= 1u; }
                }
                umma_commit(accf0 + 8 * acc);                       // accumulator complete
            }
        }
        __syncwarp();
    } else {
        // epilogue: warp w may touch TMEM lanes 32*(w%4) .. +31; thread = one accumulator row = one output pixel
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const int rx = r % g.bx, t2 = r / g.bx;
        const int ry = t2 % g.by, rn = t2 / g.by;
        uint32_t ti = 0;
        TileWalk tw;
        for (tw.init(g, blockIdx.x, gridDim.x); tw.tile < g.tiles; tw.next(g), ++ti) {
            const int x = tw.xt * g.bx + rx, y = tw.yt * g.by + ry, n = tw.nt * g.bn + rn;
            const bool live = (uint32_t)r < rows_box && x < g.wo && y < g.ho && n < g.nf;
            const size_t row = ((size_t)n * g.ho + y) * g.wo + x;
            const uint32_t acc = ti & 1u, use = ti >> 1;
            mbar_wait_relaxed(smem_u32(&bar_acc_full[acc]), use & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem_base + acc * NPAD + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
            for (int c = 0; c < NPAD; c += 16) {
                if (c >= g.n_valid) break;                          // warp-uniform
                uint32_t v[16];
                tmem_ld16(taddr + (uint32_t)c, v);
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
                            const float a = fmaxf(__uint_as_float(v[8 * j + 2 * e]) + bias_s[col], 0.0f);
                            const float b = fmaxf(__uint_as_float(v[8 * j + 2 * e + 1]) + bias_s[col + 1], 0.0f);
                            const __half2 h = __floats2half2_rn(a, b);
                            p[e] = *reinterpret_cast<const uint32_t*>(&h);
                        }
                        dst[j] = make_uint4(p[0], p[1], p[2], p[3]);
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bar_acc_empty[acc]));
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// Stride-2 convolutions with five kernel rows (conv2, conv3: keras_train.py:137-139 | 199-201) as GEMMs per INPUT row.
//
// In k_pilot_gemm a kernel-row run of the input (kw * C contiguous values of input row iy) is the A operand of an N = filters MMA once
// per (output row, kernel row) it belongs to: 2.5 times on average at stride 2, and with N = 32 / 64 those MMAs are bound by their
// operand fetch from shared memory, not by arithmetic.  Here the accumulator rows are (frame, output column) and the
// accumulator COLUMNS are (output row of the tile, filter): input row iy = 2 oy + kr contributes through kernel rows of its own parity
// only, so ONE MMA with N = 3 F (even rows: kernel rows 4, 2, 0) or 2 F (odd rows: 3, 1) multiplies the run with the weights of all those
// kernel rows at once and adds into the column blocks of the output rows oy = (iy - kr) / 2, which are adjacent.  The weights sit in shared
// memory for the CTA's lifetime as [W4 | W2 | W0] and [W3 | W1] (per 64-wide K chunk); at the top and bottom of a tile a sub-range of those
// blocks is addressed through the descriptor's start address.  Every A run is fetched once per tile (2 OYT + 3 input rows for OYT output rows).
// The first contribution to an output row (kernel row 0) must overwrite: that K step is issued as two MMAs (accumulating blocks, fresh block).
// Warp roles and barriers as in k_pilot_gemm; one CTA per SM (two accumulators of 256 TMEM columns: OYT = 256 / F output rows each).
template <int F>
constexpr int rowconv_smem_bytes(int kchunks, int stages) { return 1024 + kchunks * 5 * F * BLOCK_K * 2 + stages * A_STAGE_BYTES; }

template <int F, int KCH, int LAST_STEPS, int STAGES, int ACC_COLS = 256>
__global__ void __launch_bounds__(GEMM_THREADS) k_pilot_rowconv(const __grid_constant__ CUtensorMap map_a,
                                                                const __grid_constant__ CUtensorMap map_b, const GemmGeom g,
                                                                const float* __restrict__ bias, __half* __restrict__ out)
{
    constexpr int OYT = ACC_COLS / F;                     // output rows per tile (ACC_COLS = 128: two CTAs per SM, two independent MMA chains)
    constexpr int BLK = F * BLOCK_K * 2;                  // bytes of one kernel row's filters, one K chunk (a multiple of 1024)
    constexpr int TMEM_COLS = 2 * ACC_COLS;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[STAGES];
    __shared__ __align__(8) uint64_t bar_empty[STAGES];
    __shared__ __align__(8) uint64_t bar_acc_full[2];
    __shared__ __align__(8) uint64_t bar_acc_empty[2];
    __shared__ __align__(8) uint64_t bar_w;
    __shared__ uint32_t tmem_slot;
    __shared__ float bias_s[F];

    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t b_even = base, b_odd = b_even + (uint32_t)(KCH * 3 * BLK), a_stages = b_odd + (uint32_t)(KCH * 2 * BLK);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rows_box = (uint32_t)(g.bx * g.bn);

    for (int i = threadIdx.x; i < F; i += GEMM_THREADS) bias_s[i] = __ldg(bias + i);
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
            for (int a = 0; a < 2; ++a) {
                mbar_init(smem_u32(&bar_acc_full[a]), 1);
                mbar_init(smem_u32(&bar_acc_empty[a]), 4);
            }
            mbar_init(smem_u32(&bar_w), 1);
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
        if (elect_one()) {
            // the weights, once: kernel row kr goes to block j of its parity's operand ([4 | 2 | 0] or [3 | 1])
            const uint32_t wbar = smem_u32(&bar_w);
            mbar_expect_tx(wbar, (uint32_t)(KCH * 5 * BLK));
            for (int kr = 0; kr < 5; ++kr) {
                const int odd = kr & 1, j = odd ? (3 - kr) / 2 : (4 - kr) / 2;
                for (int c = 0; c < KCH; ++c)
                    tma_load_2d((odd ? b_odd + (uint32_t)(c * 2 * BLK) : b_even + (uint32_t)(c * 3 * BLK)) + (uint32_t)(j * BLK), &map_b, wbar,
                                (kr * KCH + c) * BLOCK_K, 0);
            }
            const uint32_t full0 = smem_u32(&bar_full[0]), empty0 = smem_u32(&bar_empty[0]);
            uint32_t s = 0, ph = 0;
            TileWalk tw;
            for (tw.init(g, blockIdx.x, gridDim.x); tw.tile < g.tiles; tw.next(g)) {
                const int x0 = tw.xt * g.bx, oy0 = tw.yt * OYT, n0 = tw.nt * g.bn;
                const int nrows = 2 * min(OYT, g.ho - oy0) + 3;
                for (int r = 0; r < nrows; ++r) {
                    const int iy = 2 * oy0 + r;
                    const int oyc = min(iy >> 1, g.ho - 1), krc = iy - 2 * oyc;      // the map addresses input rows as 2 oy + kr
                    for (int ch = 0; ch < KCH; ++ch) {
                        mbar_wait(empty0 + 8 * s, ph ^ 1u);
                        const uint32_t full = full0 + 8 * s;
                        mbar_expect_tx(full, rows_box * (BLOCK_K * 2));
                        tma_load_5d(a_stages + s * A_STAGE_BYTES, &map_a, full, ch * BLOCK_K, x0, krc, oyc, n0);
                        if (++s == STAGES) { s = 0; ph ^= 1u; }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (elect_one()) {
            const uint32_t full0 = keep(smem_u32(&bar_full[0])), empty0 = keep(smem_u32(&bar_empty[0]));
            const uint32_t accf0 = keep(smem_u32(&bar_acc_full[0])), acce0 = keep(smem_u32(&bar_acc_empty[0]));
            const uint64_t da0 = umma_desc_sw128(a_stages);
            const int ho = keep(g.ho);
            const uint32_t idesc1 = keep(umma_idesc_f16(F));
            mbar_wait(smem_u32(&bar_w), 0);
            uint32_t s = 0, ph = 0, ti = 0;
            bool ready = false;                                                        // the next stage's barrier was seen complete already
            TileWalk tw;
            for (tw.init(g, blockIdx.x, gridDim.x); tw.tile < g.tiles; tw.next(g), ++ti) {
                const int oyn = min(OYT, ho - tw.yt * OYT), nrows = 2 * oyn + 3;
                const uint32_t acc = ti & 1u, use = ti >> 1;
                mbar_wait(acce0 + 8 * acc, (use & 1u) ^ 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem_base + acc * (uint32_t)ACC_COLS;
                for (int r = 0; r < nrows; ++r) {
                    const int odd = r & 1, hh = r >> 1;
                    // blocks j_lo .. j_hi of this parity's operand are live; block j adds into output row oyl0 + (j - j_lo) of the tile
                    const int j_lo = odd ? max(0, 1 - hh) : max(0, 2 - hh);
                    const int j_hi = odd ? min(1, oyn - hh) : min(2, oyn + 1 - hh);
                    const int oyl0 = (odd ? hh - 1 : hh - 2) + j_lo;
                    const bool fresh = !odd && j_hi == 2;                              // kernel row 0: first contribution to output row hh
                    const uint32_t b_par = odd ? b_odd : b_even, b_chunk = (uint32_t)((odd ? 2 : 3) * BLK);
                    const uint32_t idesc_all = umma_idesc_f16((j_hi - j_lo + 1) * F), idesc_acc = umma_idesc_f16((j_hi - j_lo) * F);
                    const uint32_t d_row = d_tmem + (uint32_t)(oyl0 * F);
                    uint64_t db = umma_desc_sw128(b_par + (uint32_t)(j_lo * BLK));
#pragma unroll
                    for (int ch = 0; ch < KCH; ++ch, db += (uint64_t)(b_chunk >> 4)) {
                        if (!ready) mbar_wait(full0 + 8 * s, ph);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint64_t da = da0 + (uint64_t)(s * (A_STAGE_BYTES >> 4));
                        const uint32_t s_now = s;
                        if (++s == STAGES) { s = 0; ph ^= 1u; }
                        ready = mbar_test(full0 + 8 * s, ph);                         // probe the next stage while this one's MMAs are issued
                        if (ch == 0 && fresh) {
                            if (j_lo < 2) umma_f16(d_row, da, db, idesc_acc, 1u);
                            umma_f16(d_tmem + (uint32_t)(hh * F), da, db + (uint64_t)(((2 - j_lo) * BLK) >> 4), idesc1, 0u);
                        } else {
                            umma_f16(d_row, da, db, idesc_all, 1u);
                        }
#pragma unroll
                        for (int k = 1; k < (ch == KCH - 1 ? LAST_STEPS : BLOCK_K / UMMA_K); ++k) {    // +32 bytes inside the swizzle atom per K step
                            umma_f16(d_row, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc_all, 1u);
                        }
                        umma_commit(empty0 + 8 * s_now);
                    }
                }
                umma_commit(accf0 + 8 * acc);
            }
        }
        __syncwarp();
    } else {
        // epilogue: thread = accumulator row = (frame, output column); its OYT x F columns are OYT pixels of F channels
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const int rx = r % g.bx, rn = r / g.bx;
        uint32_t ti = 0;
        TileWalk tw;
        for (tw.init(g, blockIdx.x, gridDim.x); tw.tile < g.tiles; tw.next(g), ++ti) {
            const int x = tw.xt * g.bx + rx, oy0 = tw.yt * OYT, n = tw.nt * g.bn + rn;
            const int oyn = min(OYT, g.ho - oy0);
            const bool live = (uint32_t)r < rows_box && x < g.wo && n < g.nf;
            const uint32_t acc = ti & 1u, use = ti >> 1;
            mbar_wait_relaxed(smem_u32(&bar_acc_full[acc]), use & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem_base + acc * (uint32_t)ACC_COLS + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
            for (int oyl = 0; oyl < oyn; ++oyl) {
                const size_t row = ((size_t)n * g.ho + oy0 + oyl) * g.wo + x;
#pragma unroll
                for (int c = 0; c < F; c += 32) {
                    uint32_t v[16], u[16];
                    tmem_ld16(taddr + (uint32_t)(oyl * F + c), v);
                    tmem_ld16(taddr + (uint32_t)(oyl * F + c + 16), u);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (!live) continue;
                    uint4* dst = reinterpret_cast<uint4*>(out + row * g.ldc + c);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (c + 8 * j >= g.n_valid) break;
                        uint32_t p[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int col = 8 * j + 2 * e;
                            const uint32_t xa = col < 16 ? v[col & 15] : u[col & 15], xb = col < 16 ? v[(col + 1) & 15] : u[(col + 1) & 15];
                            const float a = fmaxf(__uint_as_float(xa) + bias_s[c + col], 0.0f);
                            const float b = fmaxf(__uint_as_float(xb) + bias_s[c + col + 1], 0.0f);
                            const __half2 h = __floats2half2_rn(a, b);
                            p[e] = *reinterpret_cast<const uint32_t*>(&h);
                        }
                        dst[j] = make_uint4(p[0], p[1], p[2], p[3]);
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bar_acc_empty[acc]));
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// First convolution (5x5 / 2 on the 3-channel u8 frame, keras_train.py:135 | 197) straight from the camera bytes.
//
// Through the TMA path above this layer costs more than the other six together: a kernel row is only 15 values, so 5 of every 8
// 16-byte chunks TMA moves are padding, and every input byte is fetched 6 times from L2.  Here the CTA stages the tile's input patch
// (a few KB of u8) in shared memory once, and four builder warps write the im2col rows themselves, already in the 128-byte-swizzled
// K-major layout tcgen05.mma reads: K = 5 kernel rows x 16 (15 bytes of a row = 5 pixels x RGB, + 1 whose weight is zero).  Bytes
// become fp16 by bit tricks (0x6400 | b is the half 1024 + b; one HSUB2 gives b exactly), and the 1 / 255 of
// `np.asarray(img, float32) / 255` (keras_pilot.py:49-50) is folded into the fp16 weights.  The weights (32 x 80) stay resident in
// shared memory for the CTA's lifetime.  Warps: 0..3 builders, 4 TMEM + MMA issue, 5..8 epilogue.
constexpr int C1_THREADS = 288;
constexpr int C1_STAGES = 2;
constexpr int C1_NPAD = 32;
constexpr int C1_PATCH_REGS = 6;                          // patch words a builder thread carries for the next tile
constexpr int C1_PATCH_WORDS = 128 * C1_PATCH_REGS;
constexpr int C1_K = 80;                                  // 5 kernel rows x 16
constexpr int C1_A_COLS = C1_K / 2;                       // TMEM columns of one A stage (two fp16 per column)
constexpr int C1_TMEM_A0 = 2 * C1_NPAD;                   // accumulators first, then the A stages
constexpr int C1_TMEM_COLS = 256;                         // 2 x 32 + 2 x 40 = 144 -> next power of two
constexpr int C1_B_BYTES = 2 * C1_NPAD * BLOCK_K * 2;
constexpr int c1_smem_bytes() { return C1_B_BYTES + 2 * C1_PATCH_WORDS * 4 + 1024; }

struct Conv1Geom {
    int h, w;                // input frame
    int pitch_words;         // patch row pitch in 32-bit words
    int patch_rows;          // 2 * by + 3
    unsigned long long total_bytes;   // bytes of the frames buffer of this launch
};

__device__ __forceinline__ void named_bar_sync(int id, int threads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

__global__ void __launch_bounds__(C1_THREADS) k_pilot_conv1(const __grid_constant__ CUtensorMap map_b, const GemmGeom g, const Conv1Geom c,
                                                            const uint8_t* __restrict__ frames, const float* __restrict__ bias,
                                                            __half* __restrict__ out)
{
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[C1_STAGES];
    __shared__ __align__(8) uint64_t bar_empty[C1_STAGES];
    __shared__ __align__(8) uint64_t bar_acc_full[2];
    __shared__ __align__(8) uint64_t bar_acc_empty[2];
    __shared__ __align__(8) uint64_t bar_w;
    __shared__ uint32_t tmem_slot;
    __shared__ float bias_s[C1_NPAD];

    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t b_smem = base;
    uint8_t* const gen_base = smem_raw + (base - smem_u32(smem_raw));
    uint32_t* const patch = reinterpret_cast<uint32_t*>(gen_base + C1_B_BYTES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rows_box = (uint32_t)(g.bx * g.by);

    for (int i = threadIdx.x; i < C1_NPAD; i += C1_THREADS) bias_s[i] = __ldg(bias + i);
    if (warp == 4) {
        if (lane == 0) {
            for (int s = 0; s < C1_STAGES; ++s) {
                mbar_init(smem_u32(&bar_full[s]), 4);            // one arrival per builder warp
                mbar_init(smem_u32(&bar_empty[s]), 1);
            }
            for (int a = 0; a < 2; ++a) {
                mbar_init(smem_u32(&bar_acc_full[a]), 1);
                mbar_init(smem_u32(&bar_acc_empty[a]), 4);
            }
            mbar_init(smem_u32(&bar_w), 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_b) : "memory");
            // the weights: two 64-wide atoms of 32 rows, once per CTA
            mbar_expect_tx(smem_u32(&bar_w), C1_B_BYTES);
            tma_load_2d(b_smem, &map_b, smem_u32(&bar_w), 0, 0);
            tma_load_2d(b_smem + C1_NPAD * BLOCK_K * 2, &map_b, smem_u32(&bar_w), BLOCK_K, 0);
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(C1_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_slot;

    if (warp < 4) {
        // ---- builders: thread r owns accumulator row r = output pixel (ry, rx) of the tile ----
        const int r = threadIdx.x;
        const int rx = r % g.bx, ry = r / g.bx;
        const int words = c.pitch_words * c.patch_rows;
        const uint32_t row_bytes = (uint32_t)c.w * 3u, frame_bytes = (uint32_t)c.h * row_bytes;
        const uint32_t total = (uint32_t)c.total_bytes;
        const uint32_t* const in32 = reinterpret_cast<const uint32_t*>(frames);
        // the patch words this thread carries: (patch row, word in row), fixed for the whole launch
        int pj[C1_PATCH_REGS], pk[C1_PATCH_REGS];
#pragma unroll
        for (int i = 0; i < C1_PATCH_REGS; ++i) {
            const int wi = r + i * 128;
            pj[i] = wi < words ? wi / c.pitch_words : -1;
            pk[i] = wi - pj[i] * c.pitch_words;
        }
        uint32_t pre[C1_PATCH_REGS];
        auto origin_of = [&](const TileWalk& t) { return (uint32_t)t.nt * frame_bytes + ((uint32_t)(2 * t.yt * g.by) * (uint32_t)c.w + 2u * t.xt * g.bx) * 3u; };
        // Rows below the frame (last y tile) belong to the next frame or lie past the buffer: only dead output rows read them, so
        // the one guard needed is the end of the buffer.  When a frame row is a whole number of words every patch row has the
        // origin's misalignment and a word's source is origin + a per-thread constant.
        const bool word_rows = (row_bytes & 3u) == 0;
        uint32_t poff[C1_PATCH_REGS];
#pragma unroll
        for (int i = 0; i < C1_PATCH_REGS; ++i) poff[i] = (uint32_t)pj[i] * row_bytes + 4u * (uint32_t)pk[i];
        auto fetch = [&](const TileWalk& t) {                        // patch words of tile t -> registers
            const uint32_t origin = origin_of(t);
#pragma unroll
            for (int i = 0; i < C1_PATCH_REGS; ++i) {
                pre[i] = 0;
                const uint32_t src = word_rows ? (origin & ~3u) + poff[i]
                                               : ((origin + (uint32_t)pj[i] * row_bytes) & ~3u) + 4u * (uint32_t)pk[i];
                if (pj[i] >= 0 && src < total) pre[i] = __ldg(in32 + (src >> 2));
            }
        };
        uint32_t it = 0;
        TileWalk cur, nxt;
        cur.init(g, blockIdx.x, gridDim.x);
        nxt = cur;
        if (cur.tile < g.tiles) fetch(cur);
        for (; cur.tile < g.tiles; cur = nxt, ++it) {
            uint32_t* const pbuf = patch + (it & 1u) * C1_PATCH_WORDS;
#pragma unroll
            for (int i = 0; i < C1_PATCH_REGS; ++i)
                if (pj[i] >= 0) pbuf[r + i * 128] = pre[i];
            named_bar_sync(1, 128);                                  // patch of this tile complete (the other buffer is free: see below)
            nxt.next(g);
            if (nxt.tile < g.tiles) fetch(nxt);
            const uint32_t origin = origin_of(cur);
            const uint32_t s = it % C1_STAGES, round = it / C1_STAGES;
            if (round > 0) mbar_wait(smem_u32(&bar_empty[s]), (round - 1) & 1u);
            // rows of the tile that do not exist still take part in the warp-wide TMEM stores: they build from the patch origin
            const bool live = (uint32_t)r < rows_box;
            const int bx_ = live ? rx : 0, by_ = live ? ry : 0;
            const uint32_t a_tmem = tmem_base + ((uint32_t)(warp * 32) << 16) + C1_TMEM_A0 + s * C1_A_COLS;
            const __half2 k1024 = __half2half2(__ushort_as_half((unsigned short)0x6400));
#pragma unroll
            for (int kr = 0; kr < 5; ++kr) {
                const int j = 2 * by_ + kr;
                const uint32_t o = ((origin + (uint32_t)j * row_bytes) & 3u) + 6u * (uint32_t)bx_;   // row misalignment + pixel offset
                const uint32_t* wsrc = pbuf + j * c.pitch_words + (o >> 2);
                const uint32_t sel = 0x3210u + 0x1111u * (o & 3u);
                const uint32_t w0 = wsrc[0], w1 = wsrc[1], w2 = wsrc[2], w3 = wsrc[3], w4 = wsrc[4];
                const uint32_t bw[4] = {__byte_perm(w0, w1, sel), __byte_perm(w1, w2, sel), __byte_perm(w2, w3, sel),
                                        __byte_perm(w3, w4, sel)};
                uint32_t hv[8];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    // bytes (x0 x1 x2 x3) -> halves 1024 + x (0x64xx), minus 1024 -> the integers 0..255, exact in fp16
                    const uint32_t lo = __byte_perm(bw[q], 0x64646464u, 0x4140), hi = __byte_perm(bw[q], 0x64646464u, 0x4342);
                    const __half2 hl = __hsub2(*reinterpret_cast<const __half2*>(&lo), k1024);
                    const __half2 hh = __hsub2(*reinterpret_cast<const __half2*>(&hi), k1024);
                    hv[2 * q] = *reinterpret_cast<const uint32_t*>(&hl);
                    hv[2 * q + 1] = *reinterpret_cast<const uint32_t*>(&hh);
                }
                tmem_st8(a_tmem + 8u * kr, hv);                     // k = 16 kr .. +15 of this thread's row = 8 columns of its lane
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bar_full[s]));
            // The patch buffer written next iteration is the one read two barriers ago: every builder has passed this tile's
            // named barrier after finishing the reads of the previous tile, so no second barrier is needed.
        }
    } else if (warp == 4) {
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_f16(C1_NPAD);
            mbar_wait(smem_u32(&bar_w), 0);
            const uint64_t db0 = umma_desc_sw128(b_smem), db1 = umma_desc_sw128(b_smem + C1_NPAD * BLOCK_K * 2);
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < g.tiles; tile += gridDim.x, ++it) {
                const uint32_t acc = it & 1u, use = it >> 1;
                if (use > 0) mbar_wait(smem_u32(&bar_acc_empty[acc]), (use - 1) & 1u);
                const uint32_t s = it % C1_STAGES, round = it / C1_STAGES;
                mbar_wait_relaxed(smem_u32(&bar_full[s]), round & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem_base + acc * C1_NPAD;
                const uint32_t a_tmem = tmem_base + C1_TMEM_A0 + s * C1_A_COLS;
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16_ts(d_tmem, a_tmem + 8u * k, db0 + (uint64_t)(k * 2), idesc, (uint32_t)(k != 0));
                umma_f16_ts(d_tmem, a_tmem + 32u, db1, idesc, 1u);   // k = 64..79
                umma_commit(smem_u32(&bar_empty[s]));
                umma_commit(smem_u32(&bar_acc_full[acc]));
            }
        }
        __syncwarp();
    } else {
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const int rx = r % g.bx, ry = r / g.bx;
        uint32_t ti = 0;
        TileWalk tw;
        for (tw.init(g, blockIdx.x, gridDim.x); tw.tile < g.tiles; tw.next(g), ++ti) {
            const int x = tw.xt * g.bx + rx, y = tw.yt * g.by + ry;
            const bool live = (uint32_t)r < rows_box && x < g.wo && y < g.ho;
            const size_t row = ((size_t)tw.nt * g.ho + y) * g.wo + x;
            const uint32_t acc = ti & 1u, use = ti >> 1;
            mbar_wait_relaxed(smem_u32(&bar_acc_full[acc]), use & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem_base + acc * C1_NPAD + ((uint32_t)(q * 32) << 16);
            uint32_t v[16], u[16];
            tmem_ld16(taddr, v);
            tmem_ld16(taddr + 16, u);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bar_acc_empty[acc]));      // values are in registers: the accumulator is free
            if (live) {
                uint4* dst = reinterpret_cast<uint4*>(out + row * g.ldc);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (8 * j >= g.n_valid) break;
                    uint32_t p[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int col = 8 * j + 2 * e;
                        const uint32_t xa = col < 16 ? v[col & 15] : u[col & 15], xb = col < 16 ? v[(col + 1) & 15] : u[(col + 1) & 15];
                        const float a = fmaxf(__uint_as_float(xa) + bias_s[col], 0.0f);
                        const float b = fmaxf(__uint_as_float(xb) + bias_s[col + 1], 0.0f);
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
    if (warp == 4) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C1_TMEM_COLS) : "memory");
    }
}

// First convolution per INPUT row (the formulation of k_pilot_rowconv with the builders of k_pilot_conv1).
//
// k_pilot_conv1 builds five im2col rows (one per kernel row) for every output pixel; here a builder thread turns the 15 bytes of ONE
// kernel-row run of an input row into 16 fp16 values once, and one MMA with N = 3 x 32 (even input rows: kernel rows 4, 2, 0) or
// 2 x 32 (odd rows: 3, 1) adds it into the accumulator columns of the output rows it belongs to: 17 row builds per 7 output rows
// instead of 35.  Accumulator rows = (frame, output column), columns = (output row of the tile, filter); A goes through tensor memory
// (tcgen05.st, TS-form MMA: K = 16 is one step).  The tile's input patch (<= 17 rows x 256 bytes x bn frames of u8) arrives by ONE 3-D TMA
// box per tile through a three-deep ring instead of per-thread global loads.  Needs rows of a multiple of 16 bytes (width % 16 == 0).
// Warps: C1R_SETS builder sets of four (row pairs round-robin), then four epilogue warps, the TMA producer, TMEM + MMA issue.  One CTA per SM.
constexpr int C1R_SETS = 2;                               // builder sets of four warps (a set covers the 128 TMEM lanes)
constexpr int C1R_THREADS = 128 * C1R_SETS + 192;         // + four epilogue warps, the TMA producer, the MMA issuer
constexpr int C1R_OYT = 7;                                // output rows per tile: 7 x 32 accumulator columns, twice (the rest of TMEM holds A stages)
constexpr int C1R_ROWS = 2 * C1R_OYT + 3;                 // input rows of a tile
constexpr int C1R_ROWB = 256;                             // patch row pitch: (2 bx + 3) pixels x 3 bytes <= 256
constexpr int C1R_ASTAGES = 4;                            // A stages of 16 TMEM columns (the runs of input rows 2t and 2t + 1); eight stages with
                                                          // six-row tiles, or a third builder set, measured no faster: the MMA chain paces the kernel
constexpr int C1R_PSTAGES = 3;
constexpr int C1R_ACC_COLS = C1R_OYT * C1_NPAD;           // 224
constexpr int C1R_TMEM_A0 = 2 * C1R_ACC_COLS;             // 448: four A stages of 16 columns behind the accumulators
constexpr int C1R_W_BYTES = 5 * C1_NPAD * BLOCK_K * 2;    // [W4 | W2 | W0 | W3 | W1], 32 filters x 64 K slots (16 used) each
constexpr int c1r_smem_bytes(int bn) { return 1024 + C1R_W_BYTES + C1R_PSTAGES * C1R_ROWS * C1R_ROWB * bn; }

__global__ void __launch_bounds__(C1R_THREADS) k_pilot_conv1r(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_b,
                                                              const GemmGeom g, const float* __restrict__ bias, __half* __restrict__ out)
{
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_pfull[C1R_PSTAGES];
    __shared__ __align__(8) uint64_t bar_pempty[C1R_PSTAGES];
    __shared__ __align__(8) uint64_t bar_afull[C1R_ASTAGES];
    __shared__ __align__(8) uint64_t bar_aempty[C1R_ASTAGES];
    __shared__ __align__(8) uint64_t bar_acc_full[2];
    __shared__ __align__(8) uint64_t bar_acc_empty[2];
    __shared__ __align__(8) uint64_t bar_w;
    __shared__ uint32_t tmem_slot;
    __shared__ float bias_s[C1_NPAD];

    constexpr int BLK = C1_NPAD * BLOCK_K * 2;            // 4096: one kernel row's filters
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t b_even = base, b_odd = base + 3 * BLK, patch0 = base + C1R_W_BYTES;
    const uint32_t pstage_bytes = (uint32_t)(C1R_ROWS * C1R_ROWB * g.bn);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rows_box = (uint32_t)(g.bx * g.bn);

    for (int i = threadIdx.x; i < C1_NPAD; i += C1R_THREADS) bias_s[i] = __ldg(bias + i);
    constexpr int W_EPI = 4 * C1R_SETS, W_TMA = W_EPI + 4, W_MMA = W_EPI + 5;
    if (warp == W_MMA) {
        if (lane == 0) {
            for (int s = 0; s < C1R_PSTAGES; ++s) { mbar_init(smem_u32(&bar_pfull[s]), 1); mbar_init(smem_u32(&bar_pempty[s]), 4 * C1R_SETS); }
            for (int s = 0; s < C1R_ASTAGES; ++s) { mbar_init(smem_u32(&bar_afull[s]), 4); mbar_init(smem_u32(&bar_aempty[s]), 1); }
            for (int a = 0; a < 2; ++a) { mbar_init(smem_u32(&bar_acc_full[a]), 1); mbar_init(smem_u32(&bar_acc_empty[a]), 4); }
            mbar_init(smem_u32(&bar_w), 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_slot;

    if (warp < W_EPI) {
        // ---- builders: thread r of a set owns accumulator row r = (frame rn, output column rx) of the tile ----
        const int set = warp >> 2;
        const int r = threadIdx.x & 127;
        const bool live = (uint32_t)r < rows_box;
        const int rx = live ? r % g.bx : 0, rn = live ? r / g.bx : 0;
        const uint32_t lane_off = (uint32_t)(rn * C1R_ROWS * C1R_ROWB + 6 * rx);       // this row's first byte in patch row 0 of its frame
        const uint32_t a_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + C1R_TMEM_A0;
        const uint32_t pfull0 = keep(smem_u32(&bar_pfull[0])), pempty0 = keep(smem_u32(&bar_pempty[0]));
        const uint32_t afull0 = keep(smem_u32(&bar_afull[0])), aempty0 = keep(smem_u32(&bar_aempty[0]));
        const __half2 k1024 = __half2half2(__ushort_as_half((unsigned short)0x6400));
        const uint32_t sel = 0x3210u + 0x1111u * (lane_off & 3u);                     // 6 rx is even: misalignment 0 or 2
        // one kernel-row run (15 bytes from `wsrc` on, word-aligned below it) -> 16 fp16 values (8 registers)
        auto build_run = [&](uint32_t wsrc, uint32_t (&hv)[8]) {
            uint32_t w0, w1, w2, w3, w4;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(wsrc));
            asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(w1) : "r"(wsrc));
            asm volatile("ld.shared.u32 %0, [%1+8];" : "=r"(w2) : "r"(wsrc));
            asm volatile("ld.shared.u32 %0, [%1+12];" : "=r"(w3) : "r"(wsrc));
            asm volatile("ld.shared.u32 %0, [%1+16];" : "=r"(w4) : "r"(wsrc));
            const uint32_t bw[4] = {__byte_perm(w0, w1, sel), __byte_perm(w1, w2, sel), __byte_perm(w2, w3, sel), __byte_perm(w3, w4, sel)};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                // bytes (x0 x1 x2 x3) -> halves 1024 + x (0x64xx), minus 1024 -> the integers 0..255, exact in fp16
                const uint32_t lo = __byte_perm(bw[q], 0x64646464u, 0x4140), hi = __byte_perm(bw[q], 0x64646464u, 0x4342);
                const __half2 hl = __hsub2(*reinterpret_cast<const __half2*>(&lo), k1024);
                const __half2 hh = __hsub2(*reinterpret_cast<const __half2*>(&hi), k1024);
                hv[2 * q] = *reinterpret_cast<const uint32_t*>(&hl);
                hv[2 * q + 1] = *reinterpret_cast<const uint32_t*>(&hh);
            }
        };
        // The stores of a pair are left in flight while the next pair is converted: their completion (tcgen05.wait::st) and the arrival
        // on the stage's barrier come right before the next pair's stores.
        uint32_t pending = 0;                                                         // barrier of the pair whose stores are in flight
        auto flush = [&]() {
            if (pending) {
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(pending);
                pending = 0;
            }
        };
        uint32_t it = 0, ps = 0, pph = 0;                                             // it: row pairs since the kernel started
        TileWalk tw;
        for (tw.init(g, blockIdx.x, gridDim.x); tw.tile < g.tiles; tw.next(g)) {
            const int oyn = min(C1R_OYT, g.ho - tw.yt * C1R_OYT), npairs = oyn + 2;   // rows 2t, 2t + 1; the last pair has no odd row
            mbar_wait(pfull0 + 8 * ps, pph);
            const uint32_t src0 = patch0 + ps * pstage_bytes + (lane_off & ~3u);
            for (int t = 0; t < npairs; ++t, ++it) {
                if ((int)(it % C1R_SETS) != set) continue;
                const uint32_t s = it % C1R_ASTAGES, round = it / C1R_ASTAGES;
                uint32_t he[8], ho_[8];
                build_run(src0 + (uint32_t)(2 * t * C1R_ROWB), he);
                build_run(src0 + (uint32_t)(min(2 * t + 1, C1R_ROWS - 1) * C1R_ROWB), ho_);   // (the last pair's odd row is never multiplied)
                flush();
                mbar_wait(aempty0 + 8 * s, (round & 1u) ^ 1u);                       // the MMAs that read this A stage last have completed
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                tmem_st8(a_lane + 16u * s, he);
                tmem_st8(a_lane + 16u * s + 8u, ho_);
                pending = afull0 + 8 * s;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(pempty0 + 8 * ps);                             // this warp has read the last byte of the patch
            if (++ps == C1R_PSTAGES) { ps = 0; pph ^= 1u; }
        }
        flush();
    } else if (warp == W_TMA) {
        if (elect_one()) {
            const uint32_t wbar = smem_u32(&bar_w);
            mbar_expect_tx(wbar, (uint32_t)C1R_W_BYTES);
            for (int j = 0; j < 5; ++j) tma_load_2d(base + (uint32_t)(j * BLK), &map_b, wbar, 0, j * C1_NPAD);
            const uint32_t pfull0 = keep(smem_u32(&bar_pfull[0])), pempty0 = keep(smem_u32(&bar_pempty[0]));
            uint32_t ps = 0, pph = 0;
            TileWalk tw;
            for (tw.init(g, blockIdx.x, gridDim.x); tw.tile < g.tiles; tw.next(g)) {
                mbar_wait(pempty0 + 8 * ps, pph ^ 1u);
                mbar_expect_tx(pfull0 + 8 * ps, pstage_bytes);
                tma_load_3d(patch0 + ps * pstage_bytes, &map_in, pfull0 + 8 * ps, 6 * tw.xt * g.bx, 2 * tw.yt * C1R_OYT, tw.nt * g.bn);
                if (++ps == C1R_PSTAGES) { ps = 0; pph ^= 1u; }
            }
        }
        __syncwarp();
    } else if (warp == W_MMA) {
        if (elect_one()) {
            const uint32_t afull0 = keep(smem_u32(&bar_afull[0])), aempty0 = keep(smem_u32(&bar_aempty[0]));
            const uint32_t accf0 = keep(smem_u32(&bar_acc_full[0])), acce0 = keep(smem_u32(&bar_acc_empty[0]));
            const int ho = keep(g.ho);
            const uint32_t idesc1 = keep(umma_idesc_f16(C1_NPAD));
            mbar_wait(smem_u32(&bar_w), 0);
            const uint64_t db_e = umma_desc_sw128(b_even), db_o = umma_desc_sw128(b_odd);      // [W4 | W2 | W0], [W3 | W1]
            const uint32_t idesc2 = keep(umma_idesc_f16(2 * C1_NPAD)), idesc3 = keep(umma_idesc_f16(3 * C1_NPAD));
            const uint32_t a_tmem0 = tmem_base + C1R_TMEM_A0;
            uint32_t it = 0, ti = 0;
            TileWalk tw;
            for (tw.init(g, blockIdx.x, gridDim.x); tw.tile < g.tiles; tw.next(g), ++ti) {
                const int oyn = min(C1R_OYT, ho - tw.yt * C1R_OYT), npairs = oyn + 2;
                const uint32_t acc = ti & 1u, use = ti >> 1;
                mbar_wait(acce0 + 8 * acc, (use & 1u) ^ 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem_base + acc * (uint32_t)C1R_ACC_COLS;
                for (int t = 0; t < npairs; ++t, ++it) {
                    const uint32_t s = it % C1R_ASTAGES, round = it / C1R_ASTAGES;
                    const uint32_t a_e = a_tmem0 + 16u * s, a_o = a_e + 8u;
                    mbar_wait(afull0 + 8 * s, round & 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (t >= 2 && t < oyn) {
                        // steady state: even row 2t adds kernel rows 4, 2 into output rows t - 2, t - 1 and starts output row t with kernel
                        // row 0; odd row 2t + 1 adds kernel rows 3, 1 into output rows t - 1, t
                        const uint32_t d = d_tmem + (uint32_t)((t - 2) * C1_NPAD);
                        umma_f16_ts(d, a_e, db_e, idesc2, 1u);
                        umma_f16_ts(d + 2u * C1_NPAD, a_e, db_e + (uint64_t)((2 * BLK) >> 4), idesc1, 0u);
                        umma_f16_ts(d + (uint32_t)C1_NPAD, a_o, db_o, idesc2, 1u);
                    } else {
                        // the top and bottom of the tile: only the kernel rows whose output row lies inside it
                        const int je_lo = max(0, 2 - t), je_hi = min(2, oyn + 1 - t);
                        const uint32_t d_e = d_tmem + (uint32_t)((t - 2 + je_lo) * C1_NPAD);
                        const uint64_t db = db_e + (uint64_t)((je_lo * BLK) >> 4);
                        if (je_hi == 2) {                                            // kernel row 0 starts output row t
                            if (je_lo < 2) umma_f16_ts(d_e, a_e, db, umma_idesc_f16((2 - je_lo) * C1_NPAD), 1u);
                            umma_f16_ts(d_tmem + (uint32_t)(t * C1_NPAD), a_e, db_e + (uint64_t)((2 * BLK) >> 4), idesc1, 0u);
                        } else {
                            umma_f16_ts(d_e, a_e, db, umma_idesc_f16((je_hi - je_lo + 1) * C1_NPAD), 1u);
                        }
                        if (t <= oyn) {                                              // odd row 2t + 1 exists
                            const int jo_lo = max(0, 1 - t), jo_hi = min(1, oyn - t);
                            umma_f16_ts(d_tmem + (uint32_t)((t - 1 + jo_lo) * C1_NPAD), a_o, db_o + (uint64_t)((jo_lo * BLK) >> 4),
                                        umma_idesc_f16((jo_hi - jo_lo + 1) * C1_NPAD), 1u);
                        }
                    }
                    umma_commit(aempty0 + 8 * s);
                }
                umma_commit(accf0 + 8 * acc);
            }
        }
        __syncwarp();
    } else {
        // ---- epilogue warps: thread = accumulator row = (frame, output column); OYT x 32 columns = OYT pixels ----
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const int rx = r % g.bx, rn = r / g.bx;
        uint32_t ti = 0;
        TileWalk tw;
        for (tw.init(g, blockIdx.x, gridDim.x); tw.tile < g.tiles; tw.next(g), ++ti) {
            const int x = tw.xt * g.bx + rx, oy0 = tw.yt * C1R_OYT, n = tw.nt * g.bn + rn;
            const int oyn = min(C1R_OYT, g.ho - oy0);
            const bool live = (uint32_t)r < rows_box && x < g.wo && n < g.nf;
            const uint32_t acc = ti & 1u, use = ti >> 1;
            mbar_wait_relaxed(smem_u32(&bar_acc_full[acc]), use & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem_base + acc * (uint32_t)C1R_ACC_COLS + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
            for (int oyl = 0; oyl < oyn; ++oyl) {
                const size_t row = ((size_t)n * g.ho + oy0 + oyl) * g.wo + x;
                uint32_t v[16], u[16];
                tmem_ld16(taddr + (uint32_t)(oyl * C1_NPAD), v);
                tmem_ld16(taddr + (uint32_t)(oyl * C1_NPAD + 16), u);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (!live) continue;
                uint4* dst = reinterpret_cast<uint4*>(out + row * g.ldc);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (8 * j >= g.n_valid) break;
                    uint32_t p[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int col = 8 * j + 2 * e;
                        const uint32_t xa = col < 16 ? v[col & 15] : u[col & 15], xb = col < 16 ? v[(col + 1) & 15] : u[(col + 1) & 15];
                        const float a = fmaxf(__uint_as_float(xa) + bias_s[col], 0.0f);
                        const float b = fmaxf(__uint_as_float(xb) + bias_s[col + 1], 0.0f);
                        const __half2 h = __floats2half2_rn(a, b);
                        p[e] = *reinterpret_cast<const uint32_t*>(&h);
                    }
                    dst[j] = make_uint4(p[0], p[1], p[2], p[3]);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bar_acc_empty[acc]));
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == W_MMA) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
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

// The glue behind the model for ModelType.CNN_2D / CNN_2D_SPD_FTR (keras_pilot.py:59-63 | 71-76): __cap on both outputs
// (keras_pilot.py:142-145), __smooth_steering on the first (147-153), breaking 0.0; results as float64 like the reference's floats.
__global__ void __launch_bounds__(256) k_pilot_cap(const float* __restrict__ model_out, int n, int smooth, double threshold,
                                                   double* __restrict__ steering, double* __restrict__ throttle,
                                                   double* __restrict__ breaking)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = (double)model_out[2 * i], t = (double)model_out[2 * i + 1];
    s = s < -1.0 ? -1.0 : (s > 1.0 ? 1.0 : s);
    t = t < -1.0 ? -1.0 : (t > 1.0 ? 1.0 : t);
    if (smooth) {
        if (s > threshold) s = 1.0;
        else if (s < -threshold) s = -1.0;
    }
    steering[i] = s;
    throttle[i] = t;
    breaking[i] = 0.0;
}

}  // namespace pilot
}  // namespace trs
