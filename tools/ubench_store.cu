// ubench_store.cu — what the SM -> L2 store path sustains for the output pattern of the fused kernel (P4):
// a CTA of `threads` threads writes frames of 230,400 B (f32) [+ 57,600 B (u8)] with contiguous per-warp stores.
// Variants: store width (128 / 256 bit), lanes per warp (30 / 32), CTAs per SM, threads per CTA, output region
// (streaming = every frame its own slot; resident = slots reused, stays in L2).  Reports B/clk/SM and GB/s.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

template <int WIDTH, bool WITH_U8>
__global__ void k_store(uint8_t* f32out, uint8_t* u8out, int frames, int slots, unsigned long long* cyc, unsigned seed)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int LN = 30;                                    // active lanes per warp
    const size_t FB = 230400, UB = 57600;
    const int chunk_b = WIDTH / 8;                        // bytes per lane per store
    const int per_it = LN * chunk_b;                      // bytes per warp-iteration
    const int nit = (int)(FB / per_it);                   // warp-iterations per frame
    const int per = (nit + nw - 1) / nw;
    const int w0 = warp * per, w1 = min(nit, w0 + per);
    unsigned long long t0 = clock64();
    for (int f = blockIdx.x; f < frames; f += gridDim.x) {
        const size_t slot = (size_t)(f % slots);
        uint8_t* fp = f32out + slot * FB + (size_t)w0 * per_it + lane * chunk_b;
        uint8_t* up = u8out + slot * UB + ((size_t)w0 * per_it + lane * chunk_b) / 4;
        if (lane < LN) {
            unsigned v = seed + f;
#pragma unroll 4
            for (int wi = w0; wi < w1; ++wi) {
                if (WIDTH == 128) {
                    *reinterpret_cast<uint4*>(fp) = make_uint4(v, v + 1, v + 2, v + 3);
                    if (WITH_U8) *reinterpret_cast<uint32_t*>(up) = v;
                } else {
                    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(fp), "r"(v), "r"(v + 1), "r"(v + 2), "r"(v + 3), "r"(v + 4),
                                 "r"(v + 5), "r"(v + 6), "r"(v + 7) : "memory");
                    if (WITH_U8) *reinterpret_cast<uint2*>(up) = make_uint2(v, v + 1);
                }
                fp += per_it; up += per_it / 4; v += 7;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) cyc[blockIdx.x] = clock64() - t0;
}

// one warp per CTA-frame streams the whole frame: lane L owns 640 pixels = 7,680 B of f32 and 1,920 B of u8, written as full 32-byte sectors
__global__ void k_store_lane(uint8_t* f32out, uint8_t* u8out, int frames, int slots, unsigned long long* cyc, unsigned seed, int store_warps)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t FB = 230400, UB = 57600;
    unsigned long long t0 = clock64();
    for (int f = blockIdx.x; f < frames; f += gridDim.x) {
        const size_t slot = (size_t)(f % slots);
        if (warp < store_warps && lane < 30) {
            // with k store warps, warp w takes every k-th 32-byte sector of each lane's run
            uint8_t* fp = f32out + slot * FB + (size_t)lane * 7680;
            uint8_t* up = u8out + slot * UB + (size_t)lane * 1920;
            unsigned v = seed + f;
#pragma unroll 4
            for (int i = warp; i < 240; i += store_warps) {
                asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(fp + 32 * i), "r"(v), "r"(v + 1), "r"(v + 2), "r"(v + 3), "r"(v + 4),
                             "r"(v + 5), "r"(v + 6), "r"(v + 7) : "memory");
                if ((i & 3) == 0)
                    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(up + 8 * i), "r"(v), "r"(v + 1), "r"(v + 2), "r"(v + 3), "r"(v + 4),
                                 "r"(v + 5), "r"(v + 6), "r"(v + 7) : "memory");
                v += 7;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) cyc[blockIdx.x] = clock64() - t0;
}

void run_lane(const char* name, int grid_sms, int ctas_per_sm, int store_warps, int frames_per_cta, int slots_per_cta, uint8_t* f32out, uint8_t* u8out, unsigned long long* cyc)
{
    const int grid = grid_sms * ctas_per_sm;
    const int frames = grid * frames_per_cta, slots = grid * slots_per_cta;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_store_lane<<<grid, 32 * store_warps>>>(f32out, u8out, frames, slots, cyc, 1, store_warps);
    cudaEventRecord(e0);
    k_store_lane<<<grid, 32 * store_warps>>>(f32out, u8out, frames, slots, cyc, 2, store_warps);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long* h = (unsigned long long*)malloc(sizeof(unsigned long long) * grid);
    cudaMemcpy(h, cyc, sizeof(unsigned long long) * grid, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < grid; ++i) avg += (double)h[i]; avg /= grid;
    const double bytes_cta = (double)frames_per_cta * 288000.0;
    printf("%-44s %2d CTA/SM x %d store warp(s) on %3d SMs, %s: %6.1f B/clk/SM, %7.1f GB/s, %8.0f clk per frame per CTA  [%s]\n", name, ctas_per_sm, store_warps, grid_sms,
           slots_per_cta >= frames_per_cta ? "streaming" : "L2-resident", bytes_cta * ctas_per_sm / avg, bytes_cta * grid / (ms * 1e6), avg / frames_per_cta,
           cudaGetErrorString(cudaGetLastError()));
    free(h);
}

// TMA bulk stores: the CTA keeps `nbuf` staging buffers of `piece` bytes in shared memory and streams frames out of them
// (no refill: this measures the copy engine's shared -> global rate only)
__global__ void k_store_tma(uint8_t* f32out, int frames, int slots, unsigned long long* cyc, int piece, int nbuf)
{
    extern __shared__ __align__(128) uint8_t sm[];
    for (int i = threadIdx.x; i < piece * nbuf / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const size_t FB = 288000;
    unsigned long long t0 = clock64();
    if (threadIdx.x == 0) {
        const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sm);
        int b = 0;
        for (int f = blockIdx.x; f < frames; f += gridDim.x) {
            uint8_t* dst = f32out + (size_t)(f % slots) * FB;
            for (size_t o = 0; o + piece <= FB; o += piece) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + o), "r"(sbase + b * piece), "r"(piece) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                b = (b + 1) % nbuf;
                asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(1) : "memory");
            }
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) cyc[blockIdx.x] = clock64() - t0;
}

void run_tma(int sms, int ctas_per_sm, int piece, int nbuf, int frames_per_cta, int slots_per_cta, uint8_t* f32out, unsigned long long* cyc)
{
    const int grid = sms * ctas_per_sm;
    const int frames = grid * frames_per_cta, slots = grid * slots_per_cta;
    cudaFuncSetAttribute(k_store_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, piece * nbuf);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_store_tma<<<grid, 128, piece * nbuf>>>(f32out, frames, slots, cyc, piece, nbuf);
    cudaEventRecord(e0);
    k_store_tma<<<grid, 128, piece * nbuf>>>(f32out, frames, slots, cyc, piece, nbuf);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long* h = (unsigned long long*)malloc(sizeof(unsigned long long) * grid);
    cudaMemcpy(h, cyc, sizeof(unsigned long long) * grid, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < grid; ++i) avg += (double)h[i]; avg /= grid;
    const double bytes_cta = (double)frames_per_cta * (288000 / piece) * piece;
    printf("TMA bulk store, %5d B pieces x %d buffers        %2d CTA/SM, %s: %6.1f B/clk/SM, %7.1f GB/s  [%s]\n", piece, nbuf, ctas_per_sm,
           slots_per_cta >= frames_per_cta ? "streaming" : "L2-resident", bytes_cta * ctas_per_sm / avg, bytes_cta * grid / (ms * 1e6), cudaGetErrorString(cudaGetLastError()));
    free(h);
}

template <int WIDTH, bool WITH_U8>
void run(const char* name, int sms, int ctas_per_sm, int threads, int frames_per_cta, int slots_per_cta, uint8_t* f32out, uint8_t* u8out, unsigned long long* cyc)
{
    const int grid = sms * ctas_per_sm;
    const int frames = grid * frames_per_cta, slots = grid * slots_per_cta;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_store<WIDTH, WITH_U8><<<grid, threads>>>(f32out, u8out, frames, slots, cyc, 1);
    cudaEventRecord(e0);
    k_store<WIDTH, WITH_U8><<<grid, threads>>>(f32out, u8out, frames, slots, cyc, 2);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long* h = (unsigned long long*)malloc(sizeof(unsigned long long) * grid);
    cudaMemcpy(h, cyc, sizeof(unsigned long long) * grid, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < grid; ++i) avg += (double)h[i]; avg /= grid;
    const double bytes_cta = (double)frames_per_cta * (230400.0 + (WITH_U8 ? 57600.0 : 0.0));
    printf("%-44s %2d CTA/SM x %3d thr, %s: %6.1f B/clk/SM, %7.1f GB/s, %8.0f clk per frame per CTA  [%s]\n", name, ctas_per_sm, threads,
           slots_per_cta >= frames_per_cta ? "streaming" : "L2-resident", bytes_cta * ctas_per_sm / avg, bytes_cta * grid / (ms * 1e6), avg / frames_per_cta,
           cudaGetErrorString(cudaGetLastError()));
    free(h);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    const int FR = 64;                                   // frames per CTA
    uint8_t *f32out, *u8out; unsigned long long* cyc;
    cudaMalloc(&f32out, (size_t)sms * 2 * FR * 288000 /* the bulk-store probe writes 288,000 B per slot */); cudaMalloc(&u8out, (size_t)sms * 2 * FR * 57600); cudaMalloc(&cyc, 8 * sms * 2);
    printf("device %s, %d SMs\n", p.name, sms);
    run<128, true>("STG.128 x30 + STG.32 x30 (P4 pattern)", sms, 2, 320, FR, FR, f32out, u8out, cyc);
    run<128, true>("STG.128 x30 + STG.32 x30 (P4 pattern)", sms, 2, 320, FR, 1, f32out, u8out, cyc);
    run<128, true>("STG.128 x30 + STG.32 x30 (P4 pattern)", sms, 1, 320, FR, FR, f32out, u8out, cyc);
    run<128, true>("STG.128 x30 + STG.32 x30 (P4 pattern)", sms, 1, 320, FR, 1, f32out, u8out, cyc);
    run<128, false>("STG.128 x30 only", sms, 2, 320, FR, FR, f32out, u8out, cyc);
    run<128, false>("STG.128 x30 only", sms, 1, 320, FR, 1, f32out, u8out, cyc);
    run<256, true>("STG.256 x30 + STG.64 x30", sms, 2, 320, FR, FR, f32out, u8out, cyc);
    run<256, true>("STG.256 x30 + STG.64 x30", sms, 1, 320, FR, 1, f32out, u8out, cyc);
    run<256, false>("STG.256 x30 only", sms, 1, 320, FR, 1, f32out, u8out, cyc);
    run<128, true>("STG.128 x30 + STG.32 x30", sms, 1, 64, FR, 1, f32out, u8out, cyc);
    run<128, true>("STG.128 x30 + STG.32 x30", sms, 1, 128, FR, 1, f32out, u8out, cyc);
    run<128, true>("STG.128 x30 + STG.32 x30", sms, 1, 640, FR, 1, f32out, u8out, cyc);
    run<128, true>("STG.128 x30 + STG.32 x30", sms, 1, 1024, FR, 1, f32out, u8out, cyc);
    // is the ceiling per SM or chip-wide?  the same pattern on a quarter / a sixteenth of the SMs
    run<128, true>("STG.128 x30 + STG.32 x30, 37 SMs", 37, 1, 320, FR, 1, f32out, u8out, cyc);
    run<128, true>("STG.128 x30 + STG.32 x30, 37 SMs", 37, 1, 320, FR, FR, f32out, u8out, cyc);
    run<128, true>("STG.128 x30 + STG.32 x30, 9 SMs", 9, 1, 320, FR, 1, f32out, u8out, cyc);
    // lane-contiguous runs of full sectors (a store warp that keeps the bit planes in registers)
    run_lane("STG.256 lane-contiguous (7680 B per lane)", sms, 2, 1, FR, FR, f32out, u8out, cyc);
    run_lane("STG.256 lane-contiguous (7680 B per lane)", sms, 2, 1, FR, 1, f32out, u8out, cyc);
    run_lane("STG.256 lane-contiguous (7680 B per lane)", sms, 2, 2, FR, FR, f32out, u8out, cyc);
    run_lane("STG.256 lane-contiguous (7680 B per lane)", sms, 1, 1, FR, FR, f32out, u8out, cyc);
    run_tma(sms, 1, 4800, 2, FR, 1, f32out, cyc);
    run_tma(sms, 1, 16000, 2, FR, 1, f32out, cyc);
    run_tma(sms, 1, 16000, 4, FR, 1, f32out, cyc);
    run_tma(sms, 2, 16000, 2, FR, 1, f32out, cyc);
    run_tma(sms, 2, 16000, 2, FR, FR, f32out, cyc);
    run_tma(37, 1, 16000, 4, FR, 1, f32out, cyc);
    return 0;
}
