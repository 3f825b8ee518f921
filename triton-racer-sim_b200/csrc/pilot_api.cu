// pilot_api.cu — host side of the pilots' forward pass (include/trs_b200.h: trs_pilot_*): weight repacking from Keras layout, the
// tiling of every layer, TMA tensor maps, launch sequence.  Kernels: pilot_kernels.cuh.  No CPU evaluation of the network lives here.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <string>
#include <vector>

#include "pilot_kernels.cuh"
#include "trs_internal.h"

using namespace trs::pilot;

#define CU(call)                                                       \
    do {                                                               \
        cudaError_t _e = (call);                                       \
        if (_e != cudaSuccess) return trs_i_cuda_fail(_e, #call);      \
    } while (0)

namespace {

constexpr int N_CONV = 7;
// keras_train.py:135-152 / 197-211: (kernel, stride, filters)
const int CONV_K[N_CONV] = {5, 5, 5, 3, 3, 3, 3};
const int CONV_S[N_CONV] = {2, 2, 2, 1, 1, 1, 1};
const int CONV_F[N_CONV] = {24, 32, 64, 64, 64, 128, 128};

struct Layer {
    int kh = 0, kw = 0, stride = 1, cin = 0, cin_mem = 0, cout = 0, npad = 0, stages = 0;
    int hi = 0, wi = 0, ho = 0, wo = 0;
    GemmGeom g{};
    bool row = false;             // stride-2, five-row layer run by k_pilot_rowconv (GEMMs per input row)
    GemmGeom gr{};                // its tiling: lanes = bx output columns x bn frames, by = output rows per tile
    alignas(64) CUtensorMap map_a_row;
    alignas(64) CUtensorMap map_a;
    alignas(64) CUtensorMap map_b;
    __half* w_dev = nullptr;      // [npad][nkb * 64]
    float* b_dev = nullptr;       // [npad]
    void* in = nullptr;           // activation read
    void* out = nullptr;          // activation written
    size_t out_bytes_per_frame = 0;
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = (EncodeTiledFn)p;
    return fn;
}

// Tile of the M dimension: a box of bx output columns x by output rows x bn frames with at most 128 rows that wastes the fewest
// accumulator rows (frames only stack when a box holds whole rows).
int conv1_pitch_words(int bx) { return (6 * bx + 20 + 3) / 4 + 1; }      // a patch row: 2 bx + 3 pixels, misalignment, one spare word

void choose_box(int wo, int ho, int* bx, int* by, int* bn, bool conv1 = false)
{
    double best = -1;
    *bx = *by = *bn = 1;
    for (int x = 1; x <= std::min(wo, BLOCK_M); ++x)
        for (int y = 1; y <= ho && x * y <= BLOCK_M; ++y) {
            if (conv1 && conv1_pitch_words(x) * (2 * y + 3) > C1_PATCH_WORDS) continue;   // the input patch is staged through registers
            const int nmax = (x == wo && !conv1) ? BLOCK_M / (x * y) : 1;
            for (int n = 1; n <= nmax; ++n) {
                const double tiles = (double)((wo + x - 1) / x) * ((ho + y - 1) / y) / n;
                const double eff = (double)wo * ho / (tiles * BLOCK_M);
                // on ties the widest box wins: a warp's 32 rows then span the fewest image rows (fewer bank conflicts in conv1's
                // patch reads, longer contiguous runs in every epilogue)
                if (eff > best + 1e-9 || (eff > best - 1e-9 && x > *bx)) { best = eff; *bx = x; *by = y; *bn = n; }
            }
        }
}

// k_pilot_rowconv: accumulator rows = bx output columns x bn frames (<= 128): the split of the output width that fills the most rows
void choose_row_box(int wo, int* bx, int* bn)
{
    double best = -1;
    *bx = 1; *bn = 1;
    for (int x = 1; x <= std::min(wo, BLOCK_M); ++x) {
        const int xt = (wo + x - 1) / x, n = BLOCK_M / x;
        const double eff = (double)wo * n / ((double)xt * BLOCK_M);
        if (eff > best + 1e-9 || (eff > best - 1e-9 && x > *bx)) { best = eff; *bx = x; *bn = n; }
    }
}

constexpr int ROW_STAGES = 6;
// conv2 (32 filters): one CTA per SM with 2 x 256 accumulator columns (eight output rows per tile).  Two CTAs per SM with 2 x 128 columns
// each (four-row tiles, 16 % more input rows) measured the same.
constexpr int ROW2_ACC = 256, ROW2_STAGES = 6;

const trs_tensor* find(const trs_tensor* w, int n, const std::string& name)
{
    for (int i = 0; i < n; ++i)
        if (w[i].name && name == w[i].name) return &w[i];
    return nullptr;
}

int need(const trs_tensor* w, int n, const std::string& name, std::initializer_list<int> shape, const trs_tensor** out)
{
    const trs_tensor* t = find(w, n, name);
    if (!t || !t->data) return trs_i_fail(TRS_E_ARG, "weight '%s' is missing", name.c_str());
    bool ok = t->ndim == (int)shape.size();
    int i = 0;
    for (int s : shape) { if (ok && t->shape[i] != s) ok = false; ++i; }
    if (!ok) {
        char have[64] = "", want[64] = "";
        for (int k = 0; k < t->ndim && k < 4; ++k) snprintf(have + strlen(have), sizeof have - strlen(have), "%d,", t->shape[k]);
        for (int s : shape) snprintf(want + strlen(want), sizeof want - strlen(want), "%d,", s);
        return trs_i_fail(TRS_E_ARG, "weight '%s' has shape (%s), the model needs (%s)", name.c_str(), have, want);
    }
    *out = t;
    return 0;
}

template <int NPAD, int STAGES, bool F32>
int launch_gemm(const Layer& L, int nf, int sm_count, cudaStream_t st)
{
    constexpr int smem = gemm_smem_bytes(NPAD, STAGES);
    GemmGeom g = L.g;
    g.nf = nf;
    const long long tiles = (long long)g.x_tiles * g.y_tiles * ((nf + g.bn - 1) / g.bn);
    if (tiles > 0x7fffffffLL) return trs_i_fail(TRS_E_RANGE, "too many tiles in one launch: lower max_batch");
    g.tiles = (int)tiles;
    const int per_sm = NPAD <= 128 ? 2 : 1;                                   // shared memory and 2 x NPAD TMEM columns per CTA
    const unsigned grid = (unsigned)std::min<long long>(tiles, (long long)sm_count * per_sm);
    k_pilot_gemm<NPAD, STAGES, F32><<<grid, GEMM_THREADS, smem, st>>>(L.map_a, L.map_b, g, L.b_dev, L.out);
    CU(cudaGetLastError());
    trs_i_count_launches(1);
    return 0;
}

template <int F, int KCH, int LAST, int STAGES, int ACC_COLS>
int launch_rowconv(const Layer& L, int nf, int sm_count, cudaStream_t st)
{
    GemmGeom g = L.gr;
    g.nf = nf;
    const long long tiles = (long long)g.x_tiles * g.y_tiles * ((nf + g.bn - 1) / g.bn);
    if (tiles > 0x7fffffffLL) return trs_i_fail(TRS_E_RANGE, "too many tiles in one launch: lower max_batch");
    g.tiles = (int)tiles;
    const unsigned grid = (unsigned)std::min<long long>(tiles, (long long)sm_count * (512 / (2 * ACC_COLS)));      // TMEM columns per CTA: 2 x ACC_COLS
    k_pilot_rowconv<F, KCH, LAST, STAGES, ACC_COLS><<<grid, GEMM_THREADS, rowconv_smem_bytes<F>(KCH, STAGES), st>>>(
        L.map_a_row, L.map_b, g, L.b_dev, static_cast<__half*>(L.out));
    CU(cudaGetLastError());
    trs_i_count_launches(1);
    return 0;
}

template <int NPAD, int STAGES, bool F32>
cudaError_t allow_smem()
{
    return cudaFuncSetAttribute(k_pilot_gemm<NPAD, STAGES, F32>, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_smem_bytes(NPAD, STAGES));
}

}  // namespace

struct trs_pilot {
    trs_ctx* ctx = nullptr;
    int device = 0, sm_count = 0;
    int kind = 0, h = 0, w = 0, cap = 0;
    bool no_rowconv = false;          // TRS_PILOT_ROWCONV=0 (tests run both formulations of conv1 / conv2 / conv3)
    bool c1_rows = false;             // conv1 per input row (k_pilot_conv1r): frame rows are a multiple of 16 bytes
    GemmGeom c1r{};
    __half* w1r_dev = nullptr;        // [W4 | W2 | W0 | W3 | W1] x 32 filters x 64 K slots
    alignas(64) CUtensorMap map_w1r;
    Layer L[N_CONV + 1];              // seven convolutions + the first Dense layers of the heads as one GEMM
    Conv1Geom c1{};
    float* partial = nullptr;         // (cap, ldp) fp32
    float* blob_dev = nullptr;
    HeadsArgs heads{};
    int heads_smem = 0;
    int last_n = 0;
};

namespace {

int encode_maps(trs_pilot* p, Layer& L, bool weights_only = false)
{
    EncodeTiledFn enc = encode_fn();
    if (!enc) return trs_i_fail(TRS_E_STATE, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t es = 2;
    if (!weights_only) {
        const cuuint64_t dims[5] = {(cuuint64_t)L.kw * L.cin_mem, (cuuint64_t)L.wo, (cuuint64_t)L.kh, (cuuint64_t)L.ho, (cuuint64_t)p->cap};
        const cuuint64_t strides[4] = {(cuuint64_t)L.stride * L.cin_mem * es, (cuuint64_t)L.wi * L.cin_mem * es,
                                       (cuuint64_t)L.stride * L.wi * L.cin_mem * es, (cuuint64_t)L.hi * L.wi * L.cin_mem * es};
        const cuuint32_t box[5] = {(cuuint32_t)BLOCK_K, (cuuint32_t)L.g.bx, 1u, (cuuint32_t)L.g.by, (cuuint32_t)L.g.bn};
        const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
        for (int i = 0; i < 4; ++i)
            if (strides[i] % 16) return trs_i_fail(TRS_E_ARG, "activation stride %llu of a %dx%dx%d layer is not a multiple of 16 bytes (frame width must be even)",
                                                   (unsigned long long)strides[i], L.hi, L.wi, L.cin_mem);
        CUresult r = enc(&L.map_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, L.in, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return trs_i_fail(TRS_E_STATE, "cuTensorMapEncodeTiled (activations %dx%dx%d) failed: CUresult %d", L.hi, L.wi, L.cin_mem, (int)r);
        if (L.row) {
            // the same view, one input row of bx output columns x bn frames per box
            const cuuint32_t rbox[5] = {(cuuint32_t)BLOCK_K, (cuuint32_t)L.gr.bx, 1u, 1u, (cuuint32_t)L.gr.bn};
            r = enc(&L.map_a_row, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, L.in, dims, strides, rbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return trs_i_fail(TRS_E_STATE, "cuTensorMapEncodeTiled (input rows %dx%dx%d) failed: CUresult %d", L.hi, L.wi, L.cin_mem, (int)r);
        }
    }
    {
        const cuuint64_t dims[2] = {(cuuint64_t)L.g.nkb * BLOCK_K, (cuuint64_t)L.npad};
        const cuuint64_t strides[1] = {(cuuint64_t)L.g.nkb * BLOCK_K * es};
        const cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)L.npad};
        const cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&L.map_b, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, L.w_dev, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return trs_i_fail(TRS_E_STATE, "cuTensorMapEncodeTiled (weights) failed: CUresult %d", (int)r);
    }
    return 0;
}

// B operand of a layer: [npad][kernel rows][kchunks * 64] fp16, zero where a (kw, channel) run is shorter than its chunks or a
// channel / filter is padding.  `rows` picks the filters: rows[j] = (source tensor, column) or nullptr.
struct FilterSrc { const float* data; int col; int n_cols; };

int upload_weights(Layer& L, const std::vector<FilterSrc>& filt, const std::vector<float>& bias, bool conv1 = false)
{
    const size_t ktot = (size_t)L.g.nkb * BLOCK_K;
    std::vector<__half> hb((size_t)L.npad * ktot, __float2half(0.0f));
    const int run = L.kw * L.cin_mem;
    for (int j = 0; j < (int)filt.size(); ++j) {
        if (!filt[j].data) continue;
        if (conv1) {
            // k = kernel row * 16 + (kw * 3 + c); the kernel feeds the raw bytes 0..255, so the 1 / 255 of keras_pilot.py:50 is here
            for (int r = 0; r < L.kh; ++r)
                for (int e = 0; e < run; ++e)
                    hb[(size_t)j * ktot + (size_t)r * 16 + e] =
                        __float2half_rn(filt[j].data[((size_t)r * run + e) * filt[j].n_cols + filt[j].col] / 255.0f);
            continue;
        }
        for (int r = 0; r < L.kh; ++r)
            for (int e = 0; e < run; ++e) {
                const int kw = e / L.cin_mem, c = e % L.cin_mem;
                if (c >= L.cin) continue;
                // Keras (kh, kw, in, out): ((r * KW + kw) * Cin + c) * n_cols + col;  Dense (in, out) is the kh = kw = 1 case
                const float v = filt[j].data[(((size_t)r * L.kw + kw) * L.cin + c) * filt[j].n_cols + filt[j].col];
                hb[(size_t)j * ktot + (size_t)r * L.g.kchunks * BLOCK_K + e] = __float2half_rn(v);
            }
    }
    CU(cudaMalloc(&L.w_dev, hb.size() * sizeof(__half)));
    CU(cudaMemcpy(L.w_dev, hb.data(), hb.size() * sizeof(__half), cudaMemcpyHostToDevice));
    std::vector<float> b(L.npad, 0.0f);
    for (size_t j = 0; j < bias.size() && j < b.size(); ++j) b[j] = bias[j];
    CU(cudaMalloc(&L.b_dev, b.size() * sizeof(float)));
    CU(cudaMemcpy(L.b_dev, b.data(), b.size() * sizeof(float), cudaMemcpyHostToDevice));
    return 0;
}

int append(std::vector<float>& blob, const trs_tensor* t)
{
    size_t n = 1;
    for (int i = 0; i < t->ndim; ++i) n *= (size_t)t->shape[i];
    const int off = (int)blob.size();
    blob.insert(blob.end(), t->data, t->data + n);
    return off;
}

void destroy(trs_pilot* p)
{
    if (!p) return;
    TrsDeviceGuard guard(p->device);
    for (Layer& L : p->L) {
        cudaFree(L.w_dev);
        cudaFree(L.b_dev);
        if (&L != &p->L[N_CONV]) cudaFree(L.out);
    }
    cudaFree(p->partial);
    cudaFree(p->blob_dev);
    cudaFree(p->w1r_dev);
    delete p;
}

int build(trs_pilot* p, const trs_tensor* w, int nw)
{
    // geometry of the convolutions (VALID padding: keras_train.py:135-152)
    int hi = p->h, wi = p->w, cin = 3, cin_mem = 3;
    for (int i = 0; i < N_CONV; ++i) {
        Layer& L = p->L[i];
        L.kh = L.kw = CONV_K[i];
        L.stride = CONV_S[i];
        L.cin = cin; L.cin_mem = cin_mem; L.cout = CONV_F[i];
        L.npad = L.cout <= 32 ? 32 : (L.cout <= 64 ? 64 : 128);
        L.stages = L.npad == 128 ? 3 : 4;
        L.hi = hi; L.wi = wi;
        L.ho = (hi - L.kh) / L.stride + 1;
        L.wo = (wi - L.kw) / L.stride + 1;
        if (hi < L.kh || wi < L.kw || L.ho < 1 || L.wo < 1)
            return trs_i_fail(TRS_E_RANGE, "a %dx%d frame is too small: conv%d would have no output", p->h, p->w, i + 1);
        GemmGeom& g = L.g;
        choose_box(L.wo, L.ho, &g.bx, &g.by, &g.bn, i == 0);
        g.wo = L.wo; g.ho = L.ho; g.nf = 0;
        g.x_tiles = (L.wo + g.bx - 1) / g.bx;
        g.y_tiles = (L.ho + g.by - 1) / g.by;
        g.kchunks = (L.kw * L.cin_mem + BLOCK_K - 1) / BLOCK_K;
        g.nkb = i == 0 ? 2 : L.kh * g.kchunks;             // conv1: K = 5 x 16 in two atoms (k_pilot_conv1)
        g.n_valid = L.cout;
        g.ldc = L.cout;
        g.last_steps = ((L.kw * L.cin_mem - (g.kchunks - 1) * BLOCK_K) + UMMA_K - 1) / UMMA_K;
        L.out_bytes_per_frame = (size_t)L.ho * L.wo * L.cout * sizeof(__half);
        // the two stride-2 layers behind conv1 (24 -> 32 and 32 -> 64 channels: kernel-row runs of 120 and 160 values) have row-GEMM instantiations
        const bool row2 = L.npad == 32 && g.kchunks == 2 && g.last_steps == 4, row3 = L.npad == 64 && g.kchunks == 3 && g.last_steps == 2;
        if (i > 0 && L.kh == 5 && L.stride == 2 && L.cout == L.npad && (row2 || row3) && !p->no_rowconv) {
            const int oyt = (row2 ? ROW2_ACC : 256) / L.npad;
            GemmGeom& r = L.gr;
            r = g;
            choose_row_box(L.wo, &r.bx, &r.bn);
            r.by = oyt;
            r.x_tiles = (L.wo + r.bx - 1) / r.bx;
            r.y_tiles = (L.ho + oyt - 1) / oyt;
            L.row = true;
        }
        hi = L.ho; wi = L.wo; cin = cin_mem = L.cout;
    }
    const int flat = hi * wi * cin;                       // Flatten of the NHWC tensor (keras_train.py:153)
    const bool full = p->kind == TRS_PILOT_CNN_2D_FULL_HOUSE;
    const int nfeat = p->kind == TRS_PILOT_CNN_2D_SPD_FTR ? 1 : 0;

    // workspace
    for (int i = 0; i < N_CONV; ++i) {
        CU(cudaMalloc(&p->L[i].out, (size_t)p->cap * p->L[i].out_bytes_per_frame));
        p->L[i].in = i == 0 ? nullptr : p->L[i - 1].out;     // conv1 reads the caller's u8 frames
    }
    p->c1.h = p->h; p->c1.w = p->w;
    p->c1.pitch_words = conv1_pitch_words(p->L[0].g.bx);
    p->c1.patch_rows = 2 * p->L[0].g.by + 3;

    // convolution weights
    for (int i = 0; i < N_CONV; ++i) {
        Layer& L = p->L[i];
        const std::string name = "conv" + std::to_string(i + 1);
        const trs_tensor *k, *b;
        int rc;
        if ((rc = need(w, nw, name + "/kernel", {L.kh, L.kw, L.cin, L.cout}, &k))) return rc;
        if ((rc = need(w, nw, name + "/bias", {L.cout}, &b))) return rc;
        std::vector<FilterSrc> filt(L.cout);
        for (int j = 0; j < L.cout; ++j) filt[j] = {k->data, j, L.cout};
        if ((rc = upload_weights(L, filt, std::vector<float>(b->data, b->data + L.cout), i == 0))) return rc;
        if ((rc = encode_maps(p, L, i == 0))) return rc;
    }

    // conv1 per input row: geometry, weights regrouped by kernel-row parity, their tensor map
    if (!p->no_rowconv && p->w % 16 == 0) {
        Layer& L = p->L[0];
        GemmGeom& r = p->c1r;
        r = L.g;
        double best = -1;
        for (int x = 1; x <= std::min(L.wo, 40); ++x) {                 // (2 x + 3) pixels x 3 bytes (+ misalignment) inside a 256-byte patch row
            const int xt = (L.wo + x - 1) / x, n = std::min(BLOCK_M / x, 4);
            const double eff = (double)L.wo * n / ((double)xt * BLOCK_M);
            if (eff > best + 1e-9 || (eff > best - 1e-9 && x > r.bx)) { best = eff; r.bx = x; r.bn = n; }
        }
        r.by = C1R_OYT;
        r.x_tiles = (L.wo + r.bx - 1) / r.bx;
        r.y_tiles = (L.ho + C1R_OYT - 1) / C1R_OYT;
        const trs_tensor* k = find(w, nw, "conv1/kernel");
        std::vector<__half> hb((size_t)5 * C1_NPAD * BLOCK_K, __float2half(0.0f));
        const int order[5] = {4, 2, 0, 3, 1};
        for (int b = 0; b < 5; ++b)
            for (int f = 0; f < L.cout; ++f)
                for (int e = 0; e < 15; ++e)
                    hb[((size_t)b * C1_NPAD + f) * BLOCK_K + e] = __float2half_rn(k->data[((size_t)order[b] * 15 + e) * L.cout + f] / 255.0f);
        CU(cudaMalloc(&p->w1r_dev, hb.size() * sizeof(__half)));
        CU(cudaMemcpy(p->w1r_dev, hb.data(), hb.size() * sizeof(__half), cudaMemcpyHostToDevice));
        EncodeTiledFn enc = encode_fn();
        if (!enc) return trs_i_fail(TRS_E_STATE, "cuTensorMapEncodeTiled is not available from this driver");
        const cuuint64_t dims[2] = {(cuuint64_t)BLOCK_K, (cuuint64_t)5 * C1_NPAD};
        const cuuint64_t strides[1] = {(cuuint64_t)BLOCK_K * 2};
        const cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)C1_NPAD};
        const cuuint32_t estr[2] = {1, 1};
        CUresult cr = enc(&p->map_w1r, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, p->w1r_dev, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) return trs_i_fail(TRS_E_STATE, "cuTensorMapEncodeTiled (conv1 row weights) failed: CUresult %d", (int)cr);
        CU(cudaFuncSetAttribute(k_pilot_conv1r, cudaFuncAttributeMaxDynamicSharedMemorySize, c1r_smem_bytes(r.bn)));
        p->c1_rows = true;
    }

    // heads (keras_train.py:155-166 | 213-241)
    std::vector<float> blob;
    HeadsArgs& H = p->heads;
    H.n_heads = full ? 2 : 1;
    const char* d_names[2][4] = {{"dense1", "dense2", "dense3", full ? "output_speed" : "output_layer"},
                                 {"dense4", "dense5", "dense6", "out_steering"}};
    const char* f_names[2][3] = {{"feature1", "feature2", "feature3"}, {"current_spd_1", "current_spd_2", "current_spd_3"}};
    Layer& D = p->L[N_CONV];
    D.kh = D.kw = 1; D.stride = 1; D.cin = D.cin_mem = flat; D.hi = D.wi = D.ho = D.wo = 1;
    D.npad = full ? 256 : 128;
    D.stages = 3;
    D.cout = full ? 228 : 100;
    {
        GemmGeom& g = D.g;
        g.bx = 1; g.by = 1; g.bn = BLOCK_M;
        g.wo = 1; g.ho = 1; g.x_tiles = 1; g.y_tiles = 1;
        g.kchunks = (flat + BLOCK_K - 1) / BLOCK_K;
        g.nkb = g.kchunks;
        g.n_valid = D.cout;
        g.ldc = D.npad;
        g.last_steps = ((flat - (g.kchunks - 1) * BLOCK_K) + UMMA_K - 1) / UMMA_K;
    }
    H.ldp = D.npad;
    std::vector<FilterSrc> filt(D.cout, FilterSrc{nullptr, 0, 0});
    for (int h = 0; h < H.n_heads; ++h) {
        HeadDesc& d = H.head[h];
        int rc;
        const int widths[3] = {full ? 16 : 4 * nfeat, full ? 32 : 8 * nfeat, full ? 64 : 16 * nfeat};
        int prev = 1;
        for (int i = 0; i < 3; ++i) {
            d.fw[i] = widths[i];
            if (!widths[i]) continue;
            const trs_tensor *k, *b;
            if ((rc = need(w, nw, std::string(f_names[h][i]) + "/kernel", {prev, widths[i]}, &k))) return rc;
            if ((rc = need(w, nw, std::string(f_names[h][i]) + "/bias", {widths[i]}, &b))) return rc;
            d.f_w[i] = append(blob, k);
            d.f_b[i] = append(blob, b);
            prev = widths[i];
        }
        d.n_prev = (full && h == 1) ? widths[2] : 0;        // keras_train.py:215,231: x = [image, y]; s = Concatenate([x, s])
        const int ny = widths[2] + d.n_prev;
        const trs_tensor *k1, *b1, *k2, *b2, *k3, *b3, *ko, *bo;
        d.n_out = full ? 1 : 2;
        d.out_slot = full ? (h == 0 ? 1 : 0) : 0;          // keras_train.py:239: Concatenate([out_steering, out_speed])
        d.part_col = h * 128;
        if ((rc = need(w, nw, std::string(d_names[h][0]) + "/kernel", {flat + ny, 100}, &k1))) return rc;
        if ((rc = need(w, nw, std::string(d_names[h][0]) + "/bias", {100}, &b1))) return rc;
        if ((rc = need(w, nw, std::string(d_names[h][1]) + "/kernel", {100, 50}, &k2))) return rc;
        if ((rc = need(w, nw, std::string(d_names[h][1]) + "/bias", {50}, &b2))) return rc;
        if ((rc = need(w, nw, std::string(d_names[h][2]) + "/kernel", {50, 25}, &k3))) return rc;
        if ((rc = need(w, nw, std::string(d_names[h][2]) + "/bias", {25}, &b3))) return rc;
        if ((rc = need(w, nw, std::string(d_names[h][3]) + "/kernel", {25, d.n_out}, &ko))) return rc;
        if ((rc = need(w, nw, std::string(d_names[h][3]) + "/bias", {d.n_out}, &bo))) return rc;
        for (int j = 0; j < 100; ++j) filt[d.part_col + j] = {k1->data, j, 100};      // image rows of the kernel go to the GEMM
        d.d1y_w = (int)blob.size();
        blob.insert(blob.end(), k1->data + (size_t)flat * 100, k1->data + (size_t)(flat + ny) * 100);
        d.d1_b = append(blob, b1);
        d.d2_w = append(blob, k2); d.d2_b = append(blob, b2);
        d.d3_w = append(blob, k3); d.d3_b = append(blob, b3);
        d.o_w = append(blob, ko);  d.o_b = append(blob, bo);
    }
    {
        int rc;
        if ((rc = upload_weights(D, filt, std::vector<float>()))) return rc;
        D.in = p->L[N_CONV - 1].out;
        // the flattened features of one frame are one "kernel row": dims (flat, 1, 1, 1, cap)
        D.kw = 1; D.cin_mem = flat; D.wi = 1; D.hi = 1;
        CU(cudaMalloc(&p->partial, (size_t)p->cap * H.ldp * sizeof(float)));
        D.out = p->partial;
        if ((rc = encode_maps(p, D))) return rc;
    }
    H.blob_floats = (int)blob.size();
    CU(cudaMalloc(&p->blob_dev, blob.size() * sizeof(float)));
    CU(cudaMemcpy(p->blob_dev, blob.data(), blob.size() * sizeof(float), cudaMemcpyHostToDevice));
    p->heads_smem = (H.blob_floats + HEADS_WARPS * HEADS_SCRATCH) * (int)sizeof(float);
    if (p->heads_smem > 200 * 1024) return trs_i_fail(TRS_E_RANGE, "head weights (%d bytes) do not fit shared memory", p->heads_smem);
    CU(cudaFuncSetAttribute(k_pilot_heads, cudaFuncAttributeMaxDynamicSharedMemorySize, p->heads_smem));
    CU(cudaFuncSetAttribute(k_pilot_conv1, cudaFuncAttributeMaxDynamicSharedMemorySize, c1_smem_bytes()));
    CU((allow_smem<32, 5, false>()));
    CU((allow_smem<64, 4, false>()));
    CU((allow_smem<128, 3, false>()));
    CU((allow_smem<128, 3, true>()));
    CU((allow_smem<256, 4, true>()));
    for (int i = 1; i < N_CONV; ++i) {
        const Layer& L = p->L[i];
        if (!L.row) continue;
        if (L.npad == 32) CU(cudaFuncSetAttribute(k_pilot_rowconv<32, 2, 4, ROW2_STAGES, ROW2_ACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, rowconv_smem_bytes<32>(2, ROW2_STAGES)));
        else CU(cudaFuncSetAttribute(k_pilot_rowconv<64, 3, 2, ROW_STAGES, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, rowconv_smem_bytes<64>(3, ROW_STAGES)));
    }
    return 0;
}

int run_layer(const Layer& L, int nf, bool f32, int sms, cudaStream_t st)
{
    if (L.row) return L.npad == 32 ? launch_rowconv<32, 2, 4, ROW2_STAGES, ROW2_ACC>(L, nf, sms, st) : launch_rowconv<64, 3, 2, ROW_STAGES, 256>(L, nf, sms, st);
    if (f32) return L.npad == 256 ? launch_gemm<256, 4, true>(L, nf, sms, st) : launch_gemm<128, 3, true>(L, nf, sms, st);
    switch (L.npad) {
        case 32: return launch_gemm<32, 5, false>(L, nf, sms, st);
        case 64: return launch_gemm<64, 4, false>(L, nf, sms, st);
        default: return launch_gemm<128, 3, false>(L, nf, sms, st);
    }
}

}  // namespace

extern "C" {

int trs_pilot_create(trs_ctx* ctx, int model_type, int h, int w, const trs_tensor* weights, int n_weights, int max_batch, trs_pilot** out)
{
    if (!ctx || !out || !weights || n_weights <= 0) return trs_i_fail(TRS_E_ARG, "null argument");
    *out = nullptr;
    if (model_type < TRS_PILOT_CNN_2D || model_type > TRS_PILOT_CNN_2D_FULL_HOUSE) return trs_i_fail(TRS_E_ARG, "unknown model type %d", model_type);
    if (h < 1 || w < 1 || (w & 1) || ((long long)h * w) % 4) return trs_i_fail(TRS_E_ARG, "frame %dx%d: the width must be even and h*w a multiple of 4", h, w);
    if (max_batch < 1) return trs_i_fail(TRS_E_ARG, "max_batch=%d", max_batch);
    if ((unsigned long long)max_batch * h * w * 3 >= (1ull << 32))
        return trs_i_fail(TRS_E_RANGE, "max_batch=%d frames of %dx%d exceed 4 GiB per chunk (the first layer addresses a chunk with 32 bits)", max_batch, h, w);
    trs_pilot* p = new (std::nothrow) trs_pilot();
    if (!p) return trs_i_fail(TRS_E_ARG, "out of host memory");
    p->ctx = ctx;
    p->device = trs_i_ctx_device(ctx);
    p->sm_count = trs_i_ctx_sm_count(ctx);
    p->kind = model_type; p->h = h; p->w = w; p->cap = max_batch;
    if (const char* e = getenv("TRS_PILOT_ROWCONV")) p->no_rowconv = e[0] == '0';
    TrsDeviceGuard guard(p->device);
    if (guard.err != cudaSuccess) { delete p; return trs_i_cuda_fail(guard.err, "cudaSetDevice"); }
    const int rc = build(p, weights, n_weights);
    if (rc) { destroy(p); return rc; }
    *out = p;
    return 0;
}

int trs_pilot_destroy(trs_pilot* p)
{
    destroy(p);
    return 0;
}

int trs_pilot_forward(trs_pilot* p, const uint8_t* frames_dev, int n, const float* spd_feature_dev, const float* loc_feature_dev,
                      float* out_dev, void* stream)
{
    if (!p || !frames_dev || !out_dev) return trs_i_fail(TRS_E_ARG, "null argument");
    if (n < 0) return trs_i_fail(TRS_E_ARG, "n=%d", n);
    const bool full = p->kind == TRS_PILOT_CNN_2D_FULL_HOUSE;
    if ((full || p->kind == TRS_PILOT_CNN_2D_SPD_FTR) && !spd_feature_dev) return trs_i_fail(TRS_E_ARG, "this model needs the speed feature");
    if (full && !loc_feature_dev) return trs_i_fail(TRS_E_ARG, "the full-house model needs the loc/segment feature");
    if (((uintptr_t)frames_dev & 3) != 0) return trs_i_fail(TRS_E_ARG, "frames must be 4-byte aligned");
    TrsDeviceGuard guard(p->device);
    if (guard.err != cudaSuccess) return trs_i_cuda_fail(guard.err, "cudaSetDevice");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t frame_bytes = (size_t)p->h * p->w * 3;
    // head 0 reads feature_vec_input, head 1 current_spd_input (keras_train.py:213,226; keras_pilot.py:104)
    const float* feat0 = full ? loc_feature_dev : spd_feature_dev;
    const float* feat1 = spd_feature_dev;
    for (int done = 0; done < n; done += p->cap) {
        const int m = std::min(p->cap, n - done);
        const uint8_t* chunk = frames_dev + (size_t)done * frame_bytes;
        if (p->c1_rows && ((uintptr_t)chunk & 15) == 0) {
            // conv1 per input row: the patch of a tile is one 3-D TMA box over this chunk's frames (bytes of a row, rows, frames)
            const Layer& L = p->L[0];
            GemmGeom g = p->c1r;
            g.nf = m;
            const long long tiles = (long long)g.x_tiles * g.y_tiles * ((m + g.bn - 1) / g.bn);
            if (tiles > 0x7fffffffLL) return trs_i_fail(TRS_E_RANGE, "too many tiles in one launch: lower max_batch");
            g.tiles = (int)tiles;
            EncodeTiledFn enc = encode_fn();
            alignas(64) CUtensorMap map_in;
            const cuuint64_t dims[3] = {(cuuint64_t)p->w * 3, (cuuint64_t)p->h, (cuuint64_t)m};
            const cuuint64_t strides[2] = {(cuuint64_t)p->w * 3, (cuuint64_t)frame_bytes};
            const cuuint32_t box[3] = {(cuuint32_t)C1R_ROWB, (cuuint32_t)C1R_ROWS, (cuuint32_t)g.bn};
            const cuuint32_t estr[3] = {1, 1, 1};
            CUresult cr = enc(&map_in, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(chunk), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (cr != CUDA_SUCCESS) return trs_i_fail(TRS_E_STATE, "cuTensorMapEncodeTiled (frames %dx%d) failed: CUresult %d", p->h, p->w, (int)cr);
            const unsigned grid = (unsigned)std::min<long long>(tiles, (long long)p->sm_count);
            k_pilot_conv1r<<<grid, C1R_THREADS, c1r_smem_bytes(g.bn), st>>>(map_in, p->map_w1r, g, L.b_dev, static_cast<__half*>(L.out));
            CU(cudaGetLastError());
            trs_i_count_launches(1);
        } else {
            const Layer& L = p->L[0];
            GemmGeom g = L.g;
            g.nf = m;
            const long long tiles = (long long)g.x_tiles * g.y_tiles * m;
            if (tiles > 0x7fffffffLL) return trs_i_fail(TRS_E_RANGE, "too many tiles in one launch: lower max_batch");
            g.tiles = (int)tiles;
            Conv1Geom c = p->c1;
            c.total_bytes = (unsigned long long)m * frame_bytes;
            const unsigned grid = (unsigned)std::min<long long>(tiles, (long long)p->sm_count * 2);
            k_pilot_conv1<<<grid, C1_THREADS, c1_smem_bytes(), st>>>(L.map_b, g, c, chunk, L.b_dev, static_cast<__half*>(L.out));
            CU(cudaGetLastError());
            trs_i_count_launches(1);
        }
        for (int i = 1; i <= N_CONV; ++i) {
            const int rc = run_layer(p->L[i], m, i == N_CONV, p->sm_count, st);
            if (rc) return rc;
        }
        const unsigned hgrid = (unsigned)std::min((m + HEADS_WARPS - 1) / HEADS_WARPS, p->sm_count);
        k_pilot_heads<<<hgrid, HEADS_WARPS * 32, p->heads_smem, st>>>(p->heads, p->blob_dev, p->partial, feat0 ? feat0 + done : nullptr,
                                                                       feat1 ? feat1 + done : nullptr, out_dev + (size_t)done * 2, m);
        CU(cudaGetLastError());
        trs_i_count_launches(1);
        p->last_n = m;
    }
    return 0;
}

int trs_pilot_cap(trs_ctx* ctx, const float* model_out_dev, int n, int smooth_steering, double smooth_threshold, double* steering_dev,
                  double* throttle_dev, double* breaking_dev, void* stream)
{
    if (!ctx || !model_out_dev || !steering_dev || !throttle_dev || !breaking_dev) return trs_i_fail(TRS_E_ARG, "null argument");
    if (n < 0) return trs_i_fail(TRS_E_ARG, "n=%d", n);
    if (n == 0) return 0;
    TrsDeviceGuard guard(trs_i_ctx_device(ctx));
    if (guard.err != cudaSuccess) return trs_i_cuda_fail(guard.err, "cudaSetDevice");
    k_pilot_cap<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(model_out_dev, n, smooth_steering, smooth_threshold, steering_dev,
                                                                   throttle_dev, breaking_dev);
    CU(cudaGetLastError());
    trs_i_count_launches(1);
    return 0;
}

int trs_pilot_layer_shape(trs_pilot* p, int layer, int* ho, int* wo, int* c)
{
    if (!p || layer < 0 || layer > N_CONV + 1 || !ho || !wo || !c) return trs_i_fail(TRS_E_ARG, "bad layer %d", layer);
    if (layer == 0) return trs_i_fail(TRS_E_ARG, "layer 0 is the caller's u8 frame: conv1 reads it directly");
    if (layer <= N_CONV) { *ho = p->L[layer - 1].ho; *wo = p->L[layer - 1].wo; *c = p->L[layer - 1].cout; }
    else { *ho = 1; *wo = 1; *c = p->heads.ldp; }
    return 0;
}

int trs_pilot_debug_activation(trs_pilot* p, int layer, void* host_out, unsigned long long bytes, void* stream)
{
    if (!p || layer < 0 || layer > N_CONV + 1 || !host_out) return trs_i_fail(TRS_E_ARG, "bad layer %d", layer);
    TrsDeviceGuard guard(p->device);
    if (guard.err != cudaSuccess) return trs_i_cuda_fail(guard.err, "cudaSetDevice");
    CU(cudaStreamSynchronize((cudaStream_t)stream));
    if (layer == 0) return trs_i_fail(TRS_E_ARG, "layer 0 is the caller's u8 frame: conv1 reads it directly");
    const void* src = layer <= N_CONV ? p->L[layer - 1].out : (const void*)p->partial;
    CU(cudaMemcpy(host_out, src, bytes, cudaMemcpyDeviceToHost));
    return 0;
}

}  // extern "C"
