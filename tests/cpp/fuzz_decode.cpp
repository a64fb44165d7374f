// Mutation fuzzer for tw_decode_gray (PNG / JPEG / PGM) under AddressSanitizer + UBSan: exact-size heap copies of mutated golden
// files; any out-of-bounds read, overflow or undefined shift aborts.  Built and run by tests/test_abi_cpu.py::test_decode_fuzz_sanitizers.
// g++ -O1 -g -fsanitize=address,undefined -fno-sanitize-recover=undefined -std=c++17 -Iinclude tests/cpp/fuzz_decode.cpp \
//     tidal-wave_b200/csrc/tw_jpeg.cpp tidal-wave_b200/csrc/tw_decode.cpp -lz -o /tmp/fuzz_decode && /tmp/fuzz_decode tests/golden/jpg/*.jpg tests/golden/png/*.png
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <random>
extern "C" int tw_decode_gray(const uint8_t *bytes, size_t n, uint8_t *out, size_t cap, int *w, int *h);
int main(int argc, char **argv)
{
    std::mt19937 rng(5);
    long runs = 0, ok = 0;
    for (int a = 1; a < argc; a++) {
        FILE *f = fopen(argv[a], "rb"); if (!f) continue;
        std::vector<uint8_t> d; int c; while ((c = fgetc(f)) != EOF) d.push_back((uint8_t)c); fclose(f);
        for (int it = 0; it < (argc > 0 ? 400 : 0); it++) {
            // exact-size heap copies so that ASan sees any read past the end
            std::vector<uint8_t> m(d);
            int k = 1 + rng() % 8;
            for (int j = 0; j < k; j++) m[2 + rng() % (m.size() - 2)] = (uint8_t)rng();
            if (it % 4 == 0) m.resize(4 + rng() % (m.size() - 4));
            uint8_t *buf = (uint8_t *)malloc(m.size()); for (size_t i = 0; i < m.size(); i++) buf[i] = m[i];
            int w = 0, h = 0;
            int rc = tw_decode_gray(buf, m.size(), nullptr, 0, &w, &h);
            if (rc == 0 && (long)w * h > 0 && (long)w * h <= (1 << 22)) {
                uint8_t *out = (uint8_t *)malloc((size_t)w * h);
                rc = tw_decode_gray(buf, m.size(), out, (size_t)w * h, &w, &h);
                if (rc == 0) ok++;
                free(out);
            }
            free(buf); runs++;
        }
    }
    printf("runs %ld decoded %ld\n", runs, ok);
    return 0;
}
