/*
 * trs_oracle.c — CPU restatement of the reference's observation path.  TEST INFRASTRUCTURE ONLY.
 *
 * Nothing in the product (triton-racer-sim_b200/) may include, link or call this file; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do, as the checker.
 *
 * The reference is pure Python; its pixel arithmetic lives in OpenCV (opencv-python, unpinned in the
 * reference README.md:32; pinned here by the image: opencv-python-headless 4.13.0.92), numpy and
 * CPython's math module.  Each function below restates the arithmetic reached from one reference call
 * site (cited as file:line under /root/reference/TritonRacerSim/) in plain scalar C.  The restatement is
 * pinned by tests/test_oracle_golden.py against vectors produced by running the reference itself
 * (tests/golden/make_golden.py) — the reference ships no golden vectors of its own.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAX_HSV 4

typedef struct orc_preproc_params {
    double contrast_ratio, contrast_offset, brightness_baseline;
    int32_t dynamic_brightness, color_filter_enabled, n_hsv, edge_enabled;
    double hsv_lo[ORC_MAX_HSV][3], hsv_hi[ORC_MAX_HSV][3];
    int32_t color_dest[ORC_MAX_HSV];
    int32_t edge_dest;
    double canny_a, canny_b;
} orc_preproc_params;

/* ---------------------------------------------------------------------------------------------
 * components/img_preprocessing.py:81-102  __trim_brightness_contrast
 *   cv2.mean(img[40:119]) -> per-channel f64 mean; Python sum() left to right starting from int 0;
 *   then numpy float32 in-place ops with one rounding per op, clip, truncating cast to uint8.
 * Because a pixel takes 256 values the whole transform is a 256-entry table per frame.
 * ------------------------------------------------------------------------------------------- */
void orc_brightness_lut(const uint8_t* img, int h, int w, const orc_preproc_params* p, uint8_t lut[256],
                        uint64_t roi_sums[3])
{
    int y0 = 40 < h ? 40 : h, y1 = 119 < h ? 119 : h;
    uint64_t s[3] = {0, 0, 0};
    for (int y = y0; y < y1; ++y)
        for (int x = 0; x < w; ++x)
            for (int c = 0; c < 3; ++c) s[c] += img[((size_t)y * w + x) * 3 + c];
    if (roi_sums) { roi_sums[0] = s[0]; roi_sums[1] = s[1]; roi_sums[2] = s[2]; }
    double npx = (double)(y1 > y0 ? (y1 - y0) : 0) * (double)w;
    double m0 = npx > 0 ? (double)s[0] / npx : 0.0;
    double m1 = npx > 0 ? (double)s[1] / npx : 0.0;
    double m2 = npx > 0 ? (double)s[2] / npx : 0.0;
    double cur = ((m0 + m1) + m2) + 0.0;               /* 4th cv2.mean entry is 0.0 */
    double delta = (p->brightness_baseline - cur) / 3.0;
    volatile float fdelta = (float)delta, foff = (float)p->contrast_offset, fratio = (float)p->contrast_ratio;
    for (int i = 0; i < 256; ++i) {
        volatile float v = (float)i;                   /* volatile: one rounding per step, no fma */
        if (p->dynamic_brightness) v = v + fdelta;
        v = v - foff;
        v = v * fratio;
        v = v + foff;
        if (v < 0.0f) v = 0.0f;
        if (v > 255.0f) v = 255.0f;
        lut[i] = (uint8_t)(int)v;                      /* truncation toward zero; NaN cannot occur for finite params */
    }
}

/* ---------------------------------------------------------------------------------------------
 * components/img_preprocessing.py:66  cv2.cvtColor(img, COLOR_RGB2HSV), 8-bit: H in [0,179].
 * OpenCV's integer path: 12-bit fixed point with two reciprocal tables rounded half-to-even.
 * ------------------------------------------------------------------------------------------- */
static int32_t g_sdiv[256], g_hdiv[256];
static int g_tables_ready = 0;

static void orc_init_tables(void)
{
    if (g_tables_ready) return;
    g_sdiv[0] = g_hdiv[0] = 0;
    for (int i = 1; i < 256; ++i) {
        g_sdiv[i] = (int32_t)nearbyint((double)(255 << 12) / (1.0 * i));   /* default rounding mode = half-even */
        g_hdiv[i] = (int32_t)nearbyint((double)(180 << 12) / (6.0 * i));
    }
    g_tables_ready = 1;
}

void orc_hsv_tables(int32_t sdiv[256], int32_t hdiv[256])
{
    orc_init_tables();
    memcpy(sdiv, g_sdiv, sizeof g_sdiv);
    memcpy(hdiv, g_hdiv, sizeof g_hdiv);
}

static inline void orc_rgb2hsv_px(int r, int g, int b, int* ph, int* ps, int* pv)
{
    int v = r > g ? r : g; if (b > v) v = b;
    int vmin = r < g ? r : g; if (b < vmin) vmin = b;
    int d = v - vmin;
    int s = (d * g_sdiv[v] + (1 << 11)) >> 12;
    int h0;
    if (v == r) h0 = g - b;
    else if (v == g) h0 = b - r + 2 * d;
    else h0 = r - g + 4 * d;
    int hh = h0 * g_hdiv[d] + (1 << 11);
    int h = hh >> 12;                                   /* arithmetic shift: floor for negatives */
    if (h < 0) h += 180;
    *ph = h; *ps = s; *pv = v;
}

void orc_rgb2hsv(const uint8_t* rgb, size_t npx, uint8_t* hsv)
{
    orc_init_tables();
    for (size_t i = 0; i < npx; ++i) {
        int h, s, v;
        orc_rgb2hsv_px(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2], &h, &s, &v);
        hsv[3 * i] = (uint8_t)h; hsv[3 * i + 1] = (uint8_t)s; hsv[3 * i + 2] = (uint8_t)v;
    }
}

/* components/img_preprocessing.py:71  cv2.inRange(hsv, lo, hi) on an 8-bit image with scalar bounds.
 * OpenCV first converts each bound to int32 with cvRound (round-half-to-even; a double outside the int32
 * range becomes INT_MIN, as cvtsd2si does), then tests lo <= x <= hi on integers, inclusive.  (Probed in this
 * image: lo=40.5 admits 40, lo=41.5 excludes 41, hi=100.5 excludes 101, (-1e10, 1e10) is empty.) */
int32_t orc_round_bound(double v)
{
    if (!(v > -2147483648.5 && v < 2147483647.5)) return INT32_MIN;
    return (int32_t)nearbyint(v);
}

void orc_inrange(const uint8_t* hsv, size_t npx, const double lo[3], const double hi[3], uint8_t* mask)
{
    int32_t ilo[3], ihi[3];
    for (int c = 0; c < 3; ++c) { ilo[c] = orc_round_bound(lo[c]); ihi[c] = orc_round_bound(hi[c]); }
    for (size_t i = 0; i < npx; ++i) {
        int ok = 1;
        for (int c = 0; c < 3; ++c) {
            int32_t x = hsv[3 * i + c];
            ok &= (ilo[c] <= x) & (x <= ihi[c]);
        }
        mask[i] = ok ? 255 : 0;
    }
}

/* ---------------------------------------------------------------------------------------------
 * components/img_preprocessing.py:79  cv2.Canny(img_rgb, a, b): 3-channel u8, aperture 3, L1 norm.
 *   Sobel with replicated borders per channel; per pixel keep the channel with the largest L1
 *   magnitude (first wins ties); non-maximum suppression against a zero-padded magnitude plane with
 *   15-bit fixed-point tan(22.5) / tan(67.5) sector tests; 8-connected hysteresis.
 * mag_out (h*w u16) and map_out (h*w u8: 0 none, 1 candidate, 2 strong) are optional taps.
 * ------------------------------------------------------------------------------------------- */
void orc_canny3(const uint8_t* rgb, int h, int w, double thr_a, double thr_b, uint8_t* edges,
                uint16_t* mag_out, uint8_t* map_out)
{
    if (h <= 0 || w <= 0) return;
    if (thr_a > thr_b) { double t = thr_a; thr_a = thr_b; thr_b = t; }
    int low = (int)floor(thr_a), high = (int)floor(thr_b);
    size_t npx = (size_t)h * w;
    int16_t* gx = (int16_t*)malloc(npx * sizeof(int16_t));
    int16_t* gy = (int16_t*)malloc(npx * sizeof(int16_t));
    int32_t* mg = (int32_t*)malloc(npx * sizeof(int32_t));
    uint8_t* map = (uint8_t*)calloc(npx, 1);
    int32_t* stack = (int32_t*)malloc(npx * sizeof(int32_t));
    size_t sp = 0;

#define PX(yy, xx, cc) ((int)rgb[((size_t)(yy) * w + (xx)) * 3 + (cc)])
    for (int y = 0; y < h; ++y) {
        int ym = y > 0 ? y - 1 : 0, yp = y < h - 1 ? y + 1 : h - 1;
        for (int x = 0; x < w; ++x) {
            int xm = x > 0 ? x - 1 : 0, xp = x < w - 1 ? x + 1 : w - 1;
            int bdx = 0, bdy = 0, bm = -1;
            for (int c = 0; c < 3; ++c) {
                int dx = (PX(ym, xp, c) - PX(ym, xm, c)) + 2 * (PX(y, xp, c) - PX(y, xm, c)) + (PX(yp, xp, c) - PX(yp, xm, c));
                int dy = (PX(yp, xm, c) - PX(ym, xm, c)) + 2 * (PX(yp, x, c) - PX(ym, x, c)) + (PX(yp, xp, c) - PX(ym, xp, c));
                int m = abs(dx) + abs(dy);
                if (m > bm) { bm = m; bdx = dx; bdy = dy; }          /* strict >: lowest channel wins ties */
            }
            size_t i = (size_t)y * w + x;
            gx[i] = (int16_t)bdx; gy[i] = (int16_t)bdy; mg[i] = bm;
        }
    }
#undef PX
#define MAG(yy, xx) (((yy) < 0 || (yy) >= h || (xx) < 0 || (xx) >= w) ? 0 : mg[(size_t)(yy) * w + (xx)])
    for (int y = 0; y < h; ++y) {
        for (int x = 0; x < w; ++x) {
            size_t i = (size_t)y * w + x;
            int m = mg[i];
            if (m <= low) continue;
            int xs = gx[i], ys = gy[i];
            int ax = abs(xs), ay = abs(ys) << 15;
            int tg22x = ax * 13573;
            int is_max;
            if (ay < tg22x) {
                is_max = (m > MAG(y, x - 1)) && (m >= MAG(y, x + 1));
            } else {
                int tg67x = tg22x + (ax << 16);
                if (ay > tg67x) {
                    is_max = (m > MAG(y - 1, x)) && (m >= MAG(y + 1, x));
                } else {
                    int s = ((xs ^ ys) < 0) ? -1 : 1;
                    is_max = (m > MAG(y - 1, x - s)) && (m > MAG(y + 1, x + s));
                }
            }
            if (!is_max) continue;
            if (m > high) { map[i] = 2; stack[sp++] = (int32_t)i; }
            else map[i] = 1;
        }
    }
#undef MAG
    if (mag_out) for (size_t i = 0; i < npx; ++i) mag_out[i] = (uint16_t)mg[i];
    if (map_out) memcpy(map_out, map, npx);
    /* hysteresis: flood from strong pixels through candidates, 8-connected */
    memset(edges, 0, npx);
    for (size_t k = 0; k < sp; ++k) edges[stack[k]] = 255;
    while (sp > 0) {
        int32_t i = stack[--sp];
        int y = i / w, x = i % w;
        for (int dy = -1; dy <= 1; ++dy) {
            int yy = y + dy; if (yy < 0 || yy >= h) continue;
            for (int dx = -1; dx <= 1; ++dx) {
                int xx = x + dx; if (xx < 0 || xx >= w) continue;
                size_t j = (size_t)yy * w + xx;
                if (map[j] == 1 && !edges[j]) { edges[j] = 255; stack[sp++] = (int32_t)j; }
            }
        }
    }
    free(gx); free(gy); free(mg); free(map); free(stack);
}

/* ---------------------------------------------------------------------------------------------
 * components/img_preprocessing.py:37-62  __process + __merge for one frame.
 * ------------------------------------------------------------------------------------------- */
void orc_process_frame(const uint8_t* in, int h, int w, const orc_preproc_params* p, uint8_t* out)
{
    size_t npx = (size_t)h * w;
    uint8_t lut[256];
    orc_brightness_lut(in, h, w, p, lut, NULL);
    uint8_t* adj = (uint8_t*)malloc(npx * 3);
    for (size_t i = 0; i < npx * 3; ++i) adj[i] = lut[in[i]];
    memcpy(out, adj, npx * 3);
    uint8_t* layer = (uint8_t*)malloc(npx);
    if (p->color_filter_enabled && p->n_hsv > 0) {
        uint8_t* hsv = (uint8_t*)malloc(npx * 3);
        orc_rgb2hsv(adj, npx, hsv);
        for (int k = 0; k < p->n_hsv; ++k) {
            orc_inrange(hsv, npx, p->hsv_lo[k], p->hsv_hi[k], layer);
            int ch = p->color_dest[k];
            for (size_t i = 0; i < npx; ++i) out[3 * i + ch] = layer[i];
        }
        free(hsv);
    }
    if (p->edge_enabled) {
        orc_canny3(adj, h, w, p->canny_a, p->canny_b, layer, NULL, NULL);   /* edges come from the adjusted image, not the merged one */
        int ch = p->edge_dest;
        for (size_t i = 0; i < npx; ++i) out[3 * i + ch] = layer[i];
    }
    free(layer); free(adj);
}

/* components/keras_pilot.py:49-50 / keras_train.py:41-42: float32(p) / float32(255), correctly rounded. */
void orc_normalise(const uint8_t* in, size_t n, float* out)
{
    for (size_t i = 0; i < n; ++i) { volatile float v = (float)in[i]; out[i] = v / 255.0f; }
}

/* N frames, optional f32 output, OpenMP over frames (the CPU baseline uses all cores this way). */
void orc_process_batch(const uint8_t* in, int n, int h, int w, const orc_preproc_params* p,
                       uint8_t* out_u8, float* out_f32, int nthreads)
{
    size_t fsz = (size_t)h * w * 3;
    orc_init_tables();
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel
#endif
    {
        uint8_t* tmp = out_u8 ? NULL : (uint8_t*)malloc(fsz);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 4)
#endif
        for (int i = 0; i < n; ++i) {
            uint8_t* o = out_u8 ? out_u8 + (size_t)i * fsz : tmp;
            orc_process_frame(in + (size_t)i * fsz, h, w, p, o);
            if (out_f32) orc_normalise(o, fsz, out_f32 + (size_t)i * fsz);
        }
        free(tmp);
    }
}

/* components/camera.py:36 (nearest resize) + plain window crop + normalise. */
void orc_crop_resize(const uint8_t* in, int n, int h_in, int w_in, int y0, int y1, int x0, int x1,
                     int h_out, int w_out, uint8_t* out_u8, float* out_f32)
{
    int hs = y1 - y0, ws = x1 - x0;
    for (int f = 0; f < n; ++f)
        for (int y = 0; y < h_out; ++y) {
            int sy = y0 + (int)(((long long)y * hs) / h_out);
            for (int x = 0; x < w_out; ++x) {
                int sx = x0 + (int)(((long long)x * ws) / w_out);
                for (int c = 0; c < 3; ++c) {
                    uint8_t v = in[(((size_t)f * h_in + sy) * w_in + sx) * 3 + c];
                    size_t o = (((size_t)f * h_out + y) * w_out + x) * 3 + c;
                    if (out_u8) out_u8[o] = v;
                    if (out_f32) { volatile float fv = (float)v; out_f32[o] = fv / 255.0f; }
                }
            }
        }
}

/* ---------------------------------------------------------------------------------------------
 * components/track_data_process.py:89-107  __find_closest / __distance / __map, float64, N cars.
 * ------------------------------------------------------------------------------------------- */
void orc_locate(const double* wp, int n_wp, double min_map, double max_map, const double* xyz, int n,
                int32_t* idx_out, double* seg_out, int nthreads)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(static)
#endif
    for (int k = 0; k < n; ++k) {
        double px = xyz[3 * k], py = xyz[3 * k + 1], pz = xyz[3 * k + 2];
        int sel = 0;
        volatile double last = 100.0;
        for (int i = 0; i < n_wp; ++i) {
            volatile double d = fabs(px - wp[3 * i]) + fabs(py - wp[3 * i + 1]);
            d = d + fabs(pz - wp[3 * i + 2]);
            if (d < last) { sel = i; last = d; }
        }
        if (idx_out) idx_out[k] = sel;
        if (seg_out) {
            volatile double q = (double)sel / (double)n_wp;
            q = q * (max_map - min_map);
            q = q + min_map;
            seg_out[k] = q;
        }
    }
}

/* ---------------------------------------------------------------------------------------------
 * utils/mapping.py:23-35 calcThrottle / calcBreak and the call sequence of
 * components/keras_pilot.py:80-95 (identical at 99-118), plus __cap (142-145) and
 * __smooth_steering (147-153).  Types follow numpy >= 2 promotion (NEP 50; numpy 2.3.5 in this image),
 * checked against the reference functions in tests/golden/make_golden.py: the model outputs are
 * np.float32 scalars, python ints/floats are weak, so `model_spd * 20`, `predicted * threshold`,
 * `predicted_spd - current_spd` and `delta * 2` are all float32 (current_spd is rounded to float32
 * first); math.atan then works in float64 and everything after it is float64.
 * ------------------------------------------------------------------------------------------- */
typedef struct orc_spd_params {
    double threshold, reverse_multiplier, break_multiplier;
    int32_t use_break, smooth_steering;
    double smooth_threshold;
    int32_t numpy_legacy_promotion, reserved;   /* 1: NumPy 1.x promotion, np.float32 * 20 is float64 and so is everything after it */
} orc_spd_params;

void orc_speed_control(const double* cur, const float* model_spd, const float* model_steer, int n,
                       const orc_spd_params* p, double* steer_out, double* thr_out, double* brk_out,
                       float* spd_feature)
{
    const double half_pi = 3.141592653589793 / 2.0;
    for (int k = 0; k < n; ++k) {
        double real_spd = cur[k];
        volatile float steer_f = model_steer[k];
        double steering = steer_f;
        if (steer_f < -1.0f) steering = -1.0; else if (steer_f > 1.0f) steering = 1.0;
        double delta, delta2;
        int faster;
        if (p->numpy_legacy_promotion) {
            volatile double predicted = (double)model_spd[k] * 20.0;    /* NumPy 1.x: np.float32 scalar * python int -> float64 */
            volatile double target = predicted * p->threshold;
            volatile double d = target - real_spd;
            volatile double d2 = d * 2;
            volatile double gap = predicted - real_spd;
            delta = d; delta2 = d2; faster = gap > 0.0;
        } else {
            volatile float predicted = model_spd[k] * 20.0f;             /* np.float32 * int */
            volatile float target_f = predicted * (float)p->threshold;  /* np.float32 * python float -> float32 */
            volatile float real_f = (float)real_spd;
            volatile float delta_f = target_f - real_f;
            volatile float delta2_f = delta_f * 2.0f;
            volatile float gap = predicted - real_f;
            delta = delta_f; delta2 = delta2_f; faster = gap > 0.0f;
        }
        double throttle = p->reverse_multiplier * atan(delta2) / half_pi;
        if (-0.2 < throttle && throttle < 0.0) throttle = 0.0;
        double breaking = 0.0;
        if (p->use_break) {
            throttle = faster ? 1.0 : 0.0;
            breaking = -1.0 * p->break_multiplier * atan(delta * 1.0) / half_pi;
            if (breaking < 0.4) breaking = 0.0;
        }
        if (p->smooth_steering) {
            if (steering > p->smooth_threshold) steering = 1.0;
            else if (steering < p->smooth_threshold * -1.0) steering = -1.0;
        }
        steer_out[k] = steering; thr_out[k] = throttle; brk_out[k] = breaking;
        if (spd_feature) spd_feature[k] = (float)(real_spd / 20.0);    /* np.asarray(real_spd/20, float32) */
    }
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
