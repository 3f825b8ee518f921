// Sanitizer fuzz of the product's grid walk (csrc/loc_grid.h through host_locate_check.cpp) against the reference's loop: 300 random centre lines
// (circles, flat and vertical lines, heavy repeats; scales 1e-4 .. 1e4, shifts up to 1e7) x 400 cars each (near, far, NaN, infinite, huge), built with
// -fsanitize=address,undefined: every index must match and no cell index may leave its array or overflow.  (tests/test_host_logic.py)
#include <stdio.h>
#include <random>
#include <vector>
#include "host_locate_check.cpp"
int main()
{
    std::mt19937_64 rng(7);
    std::uniform_real_distribution<double> U(0, 1);
    long total = 0;
    for (int trial = 0; trial < 300; ++trial) {
        const int n_wp = 64 + rng() % 400;
        const double scale = std::pow(10.0, (int)(rng() % 9) - 4), shift = (U(rng) - 0.5) * std::pow(10.0, (int)(rng() % 8));
        std::vector<double> wp(3 * n_wp);
        const int shape = trial % 4;
        for (int i = 0; i < n_wp; ++i) {
            const double t = 6.283185307 * i / n_wp;
            double x = std::cos(t) * 40, z = std::sin(t) * (shape == 1 ? 0.0 : 30), y = 0.5 + 0.01 * U(rng);
            if (shape == 2) { x = 0; }                       // a vertical line
            if (shape == 3) { x = (double)(rng() % 5); z = (double)(rng() % 5); }   // many repeats
            wp[3 * i] = x * scale + shift; wp[3 * i + 1] = y; wp[3 * i + 2] = z * scale - shift;
        }
        const int n = 400;
        std::vector<double> xyz(3 * n);
        for (int k = 0; k < n; ++k) {
            const int j = rng() % n_wp;
            const double spread = std::pow(10.0, (int)(rng() % 7) - 3) * scale;
            xyz[3 * k] = wp[3 * j] + (U(rng) - 0.5) * spread * 100; xyz[3 * k + 1] = wp[3 * j + 1] + (U(rng) - 0.5) * 3; xyz[3 * k + 2] = wp[3 * j + 2] + (U(rng) - 0.5) * spread * 100;
            if (k % 97 == 0) xyz[3 * k] = std::nan("");
            if (k % 89 == 0) xyz[3 * k + 2] = (k & 1) ? INFINITY : -INFINITY;
            if (k % 83 == 0) xyz[3 * k] = 1e308;
        }
        std::vector<int32_t> idx(n), ref(n);
        long long stats[6];
        locate_grid_host(wp.data(), n_wp, xyz.data(), n, idx.data(), stats);
        // brute force, the reference's loop
        for (int k = 0; k < n; ++k) {
            double best = 100.0; int r = 0;
            for (int i = 0; i < n_wp; ++i) {
                const double d = (std::fabs(xyz[3 * k] - wp[3 * i]) + std::fabs(xyz[3 * k + 1] - wp[3 * i + 1])) + std::fabs(xyz[3 * k + 2] - wp[3 * i + 2]);
                if (d < best) { best = d; r = i; }
            }
            if (r != idx[k]) { printf("MISMATCH trial %d car %d: %d vs %d\n", trial, k, idx[k], r); return 1; }
        }
        total += n;
    }
    printf("ok %ld cars\n", total);
    return 0;
}
