// Host build of the product's per-pixel arithmetic (csrc/pixel_math.cuh: pure host / device functions, the source the generic kernel compiles) for
// the CPU-side check against the oracle and against OpenCV itself (tests/test_host_logic.py).  Test infrastructure.
#include <math.h>

#include "../triton-racer-sim_b200/csrc/pixel_math.cuh"
using namespace trs;

// every colour of [first, first + n): (r, g, b) = (c >> 16, c >> 8, c) -> h, s, v bytes
extern "C" void math_rgb2hsv_range(unsigned first, unsigned n, uint8_t* hsv)
{
    int32_t sdiv[256], hdiv[256];
    for (int i = 0; i < 256; ++i) {
        sdiv[i] = i ? (int32_t)rint((255 << 12) / (double)i) : 0;
        hdiv[i] = i ? (int32_t)rint((180 << 12) / (6.0 * i)) : 0;
    }
    for (unsigned k = 0; k < n; ++k) {
        const unsigned c = first + k;
        int h, s, v;
        rgb2hsv_px((c >> 16) & 255, (c >> 8) & 255, c & 255, sdiv, hdiv, h, s, v);
        hsv[3 * k] = (uint8_t)h; hsv[3 * k + 1] = (uint8_t)s; hsv[3 * k + 2] = (uint8_t)v;
    }
}

extern "C" int math_in_range(int h, int s, int v, const int32_t* lo, const int32_t* hi)
{
    HsvRange r;
    for (int c = 0; c < 3; ++c) { r.lo[c] = lo[c]; r.hi[c] = hi[c]; }
    return in_range_px(h, s, v, r) ? 1 : 0;
}

// direction classes of all gradients (dx, dy) in [-m, m]^2, row-major in dy then dx
extern "C" void math_canny_dirs(int m, uint8_t* out)
{
    for (int dy = -m; dy <= m; ++dy)
        for (int dx = -m; dx <= m; ++dx) *out++ = (uint8_t)canny_dir(dx, dy);
}

extern "C" void math_adjust_table(int dynamic, float fdelta, float foff, float fratio, uint8_t* lut)
{
    for (int i = 0; i < 256; ++i) lut[i] = adjust_entry(i, dynamic != 0, fdelta, foff, fratio);
}

extern "C" double math_brightness_delta(unsigned long long s0, unsigned long long s1, unsigned long long s2, double npx, double baseline)
{
    return brightness_delta(s0, s1, s2, npx, baseline);
}

// Exhaustive check of the threshold tables the fast kernels use instead of computing saturation and hue (csrc/preproc_fast.cuh, init_tables):
// returns the number of (v, d, K) triples for which "s <= K" and "d <= T_K[v]" disagree, K over [k_lo, k_hi].
extern "C" long long math_check_sat_thresholds(int k_lo, int k_hi)
{
    long long bad = 0;
    for (int v = 0; v < 256; ++v) {
        const int sd = hsv_sdiv_entry(v);
        for (long long K = k_lo; K <= k_hi; ++K) {
            const uint32_t T = sat_threshold_entry(K, v, sd);
            const int t = T == 0x8001u ? -1 : (int)T;                       // the packed compare reads 0x8001 as -1
            for (int d = 0; d <= v; ++d) {
                const int s = (d * sd + 2048) >> 12;
                bad += (s <= K) != (d <= t);
            }
        }
    }
    return bad;
}

// ... and of the hue thresholds for one range [lo, hi] with hi <= 149: every delta d and every numerator h0 the three formulas can produce
// (g - b in [-d, d], b - r + 2d in [d, 3d], r - g + 4d in [3d, 5d]); the wrapped hue of a negative numerator is h + 180.
extern "C" long long math_check_hue_thresholds(int lo, int hi)
{
    long long bad = 0;
    for (int d = 0; d < 256; ++d) {
        const int hd = hsv_hdiv_entry(d);
        const uint32_t e = hue_threshold_entry(lo, hi, hd);
        const int a_lo = (int)(e & 0xffffu) - 2048, a_hi = (int)(e >> 16) - 2048;
        for (int h0 = -d; h0 <= 5 * d; ++h0) {
            int h = (h0 * hd + 2048) >> 12;
            h += h < 0 ? 180 : 0;
            bad += (lo <= h && h <= hi) != (a_lo <= h0 && h0 <= a_hi);
        }
    }
    return bad;
}
