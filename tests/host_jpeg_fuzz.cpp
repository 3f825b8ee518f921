// AddressSanitizer fuzz of the product's JPEG arithmetic (csrc/jpeg_core.cuh + csrc/jpeg_host.h, the source the kernels compile) through the
// record loop of host_jpeg_check.cpp: telemetry JPEGs are external input, a malformed file must end in an error code or in garbage pixels,
// never in a read or write outside its buffers.  usage: host_jpeg_fuzz <h> <w> <iterations> <file.jpg>...   (tests/test_jpeg_host.py)
#include <stdio.h>
#include <string.h>

#include <random>
#include <vector>

#include "host_jpeg_check.cpp"

int main(int argc, char** argv)
{
    if (argc < 5) return 2;
    const int h = atoi(argv[1]), w = atoi(argv[2]), iters = atoi(argv[3]);
    std::mt19937 rng(12345);
    long runs = 0, ok = 0;
    std::vector<uint8_t> rgb((size_t)h * w * 3);
    for (int a = 4; a < argc; ++a) {
        FILE* f = fopen(argv[a], "rb");
        if (!f) return 3;
        std::vector<uint8_t> data;
        uint8_t buf[65536];
        size_t n;
        while ((n = fread(buf, 1, sizeof buf, f)) > 0) data.insert(data.end(), buf, buf + n);
        fclose(f);
        for (int it = 0; it < iters; ++it) {
            std::vector<uint8_t> m = data;
            switch (it % 5) {
            case 0: { const int k = 1 + rng() % 4; for (int i = 0; i < k; ++i) m[rng() % m.size()] = (uint8_t)rng(); break; }      // a few random bytes
            case 1: m.resize(rng() % m.size()); break;                                                                              // truncation
            case 2: { const size_t p = rng() % m.size(), l = 1 + rng() % 64; for (size_t i = p; i < p + l && i < m.size(); ++i) m[i] = 0xff; break; }   // marker bytes
            case 3: { const size_t hdr = m.size() < 700 ? m.size() : 700; m[rng() % hdr] = (uint8_t)rng(); break; }               // tables / frame header
            default: { const size_t p = rng() % m.size(); m.insert(m.begin() + p, (size_t)(rng() % 32), (uint8_t)rng()); }         // inserted bytes
            }
            uint8_t* heap = (uint8_t*)malloc(m.size() ? m.size() : 1);      // exact size: the sanitizer sees any read past the end
            memcpy(heap, m.data(), m.size());
            const int rc = jpg_decode_rgb_host(heap, m.size(), h, w, rgb.data());
            free(heap);
            ++runs;
            ok += rc == 0;
        }
    }
    printf("runs %ld decoded %ld\n", runs, ok);
    return 0;
}
