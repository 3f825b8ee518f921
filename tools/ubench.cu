// ubench.cu — instruction-throughput probes on sm_100a used to size the per-pixel instruction budget.
// Each probe runs ILP independent dependency chains per thread, 1024 threads per CTA, one CTA per SM, and reports
// warp-instructions per clock per SM (4.0 = every scheduler issues every cycle).
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define ITERS 4096
#define ILP 8

template <class F>
__global__ void __launch_bounds__(1024, 1) probe(F f, unsigned* out, unsigned long long* cycles, unsigned seed)
{
    unsigned r[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) r[i] = seed * (threadIdx.x + 1) + i * 0x9e3779b9u;
    unsigned a = seed | 1, b = seed * 3 + 5;
    __syncthreads();
    unsigned long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) r[i] = f(r[i], a, b);
    }
    unsigned long long t1 = clock64();
    unsigned acc = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc ^= r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

struct OpLop { static constexpr int n = 1; __device__ unsigned operator()(unsigned x, unsigned a, unsigned b) const { return (x ^ a) & (x | b); } };  // LOP3
struct OpAdd { static constexpr int n = 1; __device__ unsigned operator()(unsigned x, unsigned a, unsigned b) const { return x + a + b; } };             // IADD3
struct OpPrmt { static constexpr int n = 1; __device__ unsigned operator()(unsigned x, unsigned a, unsigned b) const { return __byte_perm(x, a, 0x2103); } };
struct OpImad { static constexpr int n = 1; __device__ unsigned operator()(unsigned x, unsigned a, unsigned b) const { return x * a + b; } };
struct OpShf { static constexpr int n = 1; __device__ unsigned operator()(unsigned x, unsigned a, unsigned b) const { return __funnelshift_l(x, a, 7); } };
struct OpHfma2 { static constexpr int n = 1; __device__ unsigned operator()(unsigned x, unsigned a, unsigned b) const {
    __half2 r = __hfma2(*reinterpret_cast<__half2*>(&x), *reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b)); return *reinterpret_cast<unsigned*>(&r); } };
struct OpHadd2 { static constexpr int n = 1; __device__ unsigned operator()(unsigned x, unsigned a, unsigned b) const {
    __half2 r = __hadd2(*reinterpret_cast<__half2*>(&x), *reinterpret_cast<__half2*>(&a)); return *reinterpret_cast<unsigned*>(&r); } };
struct OpHmax2 { static constexpr int n = 1; __device__ unsigned operator()(unsigned x, unsigned a, unsigned b) const {
    __half2 r = __hmax2(*reinterpret_cast<__half2*>(&x), *reinterpret_cast<__half2*>(&a)); return *reinterpret_cast<unsigned*>(&r); } };
struct OpHset2 { static constexpr int n = 1; __device__ unsigned operator()(unsigned x, unsigned a, unsigned b) const {
    return __hgt2_mask(*reinterpret_cast<__half2*>(&x), *reinterpret_cast<__half2*>(&a)) ^ b; } };   // HSET2 + LOP
struct OpVmax2 { static constexpr int n = 1; __device__ unsigned operator()(unsigned x, unsigned a, unsigned b) const { return __vmaxu2(x, a); } };
struct OpVadd2 { static constexpr int n = 1; __device__ unsigned operator()(unsigned x, unsigned a, unsigned b) const { return __vadd2(x, a); } };
struct OpVabsd4 { static constexpr int n = 1; __device__ unsigned operator()(unsigned x, unsigned a, unsigned b) const { return __vabsdiffu4(x, a); } };
struct OpVmax3 { static constexpr int n = 1; __device__ unsigned operator()(unsigned x, unsigned a, unsigned b) const { return __vimax3_s16x2(x, a, b); } };
struct OpFfma { static constexpr int n = 1; __device__ unsigned operator()(unsigned x, unsigned a, unsigned b) const {
    return __float_as_uint(fmaf(__uint_as_float(x), 1.0001f, 0.5f)); } };
struct OpFadd { static constexpr int n = 1; __device__ unsigned operator()(unsigned x, unsigned a, unsigned b) const {
    return __float_as_uint(__fadd_rn(__uint_as_float(x), __uint_as_float(a))); } };
struct OpPopc { static constexpr int n = 1; __device__ unsigned operator()(unsigned x, unsigned a, unsigned b) const { return __popc(x) + a; } };   // POPC + IADD
struct OpBrev { static constexpr int n = 1; __device__ unsigned operator()(unsigned x, unsigned a, unsigned b) const { return __brev(x) ^ a; } };
struct OpSel { static constexpr int n = 2; __device__ unsigned operator()(unsigned x, unsigned a, unsigned b) const { return x > a ? x - b : x + b; } };  // ISETP + SEL-ish
struct OpI2f { static constexpr int n = 1; __device__ unsigned operator()(unsigned x, unsigned a, unsigned b) const { return __float_as_uint((float)(x & 0xff)) ^ a; } };
struct OpFdiv { static constexpr int n = 1; __device__ unsigned operator()(unsigned x, unsigned a, unsigned b) const { return __float_as_uint(__fdiv_rn((float)(x & 0xff), 255.0f)) ^ a; } };
struct OpShfl { static constexpr int n = 1; __device__ unsigned operator()(unsigned x, unsigned a, unsigned b) const { return __shfl_xor_sync(0xffffffffu, x, 1) + a; } };
struct OpMixLopHfma { static constexpr int n = 2; __device__ unsigned operator()(unsigned x, unsigned a, unsigned b) const {
    unsigned y = (x ^ a) & (x | b); __half2 r = __hadd2(*reinterpret_cast<__half2*>(&y), *reinterpret_cast<__half2*>(&a)); return *reinterpret_cast<unsigned*>(&r); } };
struct OpMixAddImad { static constexpr int n = 2; __device__ unsigned operator()(unsigned x, unsigned a, unsigned b) const { return (x + a + b) * a + b; } };
struct OpMixLopImad { static constexpr int n = 2; __device__ unsigned operator()(unsigned x, unsigned a, unsigned b) const { return ((x ^ a) & (x | b)) * a + b; } };
struct OpMix3 { static constexpr int n = 3; __device__ unsigned operator()(unsigned x, unsigned a, unsigned b) const {   // 2 ALU : 1 FMA
    unsigned y = (x ^ a) & (x | b); y = __byte_perm(y, a, 0x2103); return y * a + b; } };
struct OpHaddDenorm { static constexpr int n = 1; __device__ unsigned operator()(unsigned x, unsigned a, unsigned b) const {
    unsigned xx = x & 0x03ff03ffu; __half2 r = __hadd2(*reinterpret_cast<__half2*>(&xx), __halves2half2(__ushort_as_half(5), __ushort_as_half(9))); return *reinterpret_cast<unsigned*>(&r); } };  // LOP + HADD2 on subnormals

__global__ void __launch_bounds__(1024, 1) probe_lds(unsigned* out, unsigned long long* cycles, int stride_words, int vec)
{
    __shared__ uint4 s[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) s[i] = make_uint4(i, i * 3, i * 5, i * 7);
    __syncthreads();
    unsigned acc = 0;
    const unsigned* sw = reinterpret_cast<const unsigned*>(s);
    unsigned idx = (threadIdx.x * stride_words) & 8191;
    unsigned long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (vec == 4) { uint4 v = s[((idx >> 2) + i * 32 + it) & 2047]; acc += v.x ^ v.y ^ v.z ^ v.w; }
            else acc += sw[(idx + i * 1024 + it) & 8191];
        }
    }
    unsigned long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <class F>
void run(const char* name, F f, int sms, unsigned* out, unsigned long long* cyc)
{
    probe<<<sms, 1024>>>(f, out, cyc, 12345u);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<<<sms, 1024>>>(f, out, cyc, 12345u);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long h[256]; cudaMemcpy(h, cyc, sizeof(unsigned long long) * sms, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < sms; ++i) avg += (double)h[i]; avg /= sms;
    double winstr = (double)ITERS * ILP * F::n * 32.0;      // warp instructions per CTA (32 warps)
    printf("%-16s %6.3f warp-instr/clk/SM  (%.0f cycles, %.3f ms, eff clock %.0f MHz)\n", name, winstr / avg, avg, ms, avg / (ms * 1e3));
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    printf("device %s, %d SMs, clock %d kHz\n", p.name, sms, p.clockRate);
    unsigned* out; unsigned long long* cyc;
    cudaMalloc(&out, sizeof(unsigned) * sms * 1024); cudaMalloc(&cyc, sizeof(unsigned long long) * sms);
    run("LOP3x2", OpLop(), sms, out, cyc);      // note: (x^a)&(x|b) may be one LOP3
    run("IADD3", OpAdd(), sms, out, cyc);
    run("PRMT", OpPrmt(), sms, out, cyc);
    run("IMAD", OpImad(), sms, out, cyc);
    run("SHF", OpShf(), sms, out, cyc);
    run("HFMA2", OpHfma2(), sms, out, cyc);
    run("HADD2", OpHadd2(), sms, out, cyc);
    run("HADD2.denorm+LOP", OpHaddDenorm(), sms, out, cyc);
    run("HMNMX2", OpHmax2(), sms, out, cyc);
    run("HSET2+LOP", OpHset2(), sms, out, cyc);
    run("VIMNMX.U16x2", OpVmax2(), sms, out, cyc);
    run("VIADD.16x2", OpVadd2(), sms, out, cyc);
    run("VABSDIFF4", OpVabsd4(), sms, out, cyc);
    run("VIMNMX3.S16x2", OpVmax3(), sms, out, cyc);
    run("FFMA", OpFfma(), sms, out, cyc);
    run("FADD", OpFadd(), sms, out, cyc);
    run("POPC+IADD", OpPopc(), sms, out, cyc);
    run("BREV+LOP", OpBrev(), sms, out, cyc);
    run("ISETP+SEL(2)", OpSel(), sms, out, cyc);
    run("I2F+LOPs", OpI2f(), sms, out, cyc);
    run("FDIV_RN(1)", OpFdiv(), sms, out, cyc);
    run("SHFL+IADD", OpShfl(), sms, out, cyc);
    run("mix LOP+HADD2(2)", OpMixLopHfma(), sms, out, cyc);
    run("mix IADD+IMAD(2)", OpMixAddImad(), sms, out, cyc);
    run("mix LOP+IMAD(2)", OpMixLopImad(), sms, out, cyc);
    run("mix 2ALU+IMAD(3)", OpMix3(), sms, out, cyc);
    for (int vec = 1; vec <= 4; vec += 3)
        for (int stride = 1; stride <= 4; stride *= 2) {
            probe_lds<<<sms, 1024>>>(out, cyc, stride * vec, vec);
            cudaDeviceSynchronize();
            unsigned long long h[256]; cudaMemcpy(h, cyc, sizeof(unsigned long long) * sms, cudaMemcpyDeviceToHost);
            double avg = 0; for (int i = 0; i < sms; ++i) avg += (double)h[i]; avg /= sms;
            printf("LDS.%d stride %d words: %6.3f warp-instr/clk/SM, %.1f B/clk/SM\n", vec * 32, stride * vec, ITERS * ILP * 32.0 / avg, ITERS * ILP * 32.0 * 32 * 4 * vec / avg);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
