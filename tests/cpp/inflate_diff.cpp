// Differential test of tw_inflate.h against zlib (tests/test_abi_cpu.py::test_inflate_matches_zlib, built with ASan + UBSan): random
// contents x every deflate level / strategy / window / flush pattern, exact / shorter / longer output sizes, truncated streams, and
// single-bit mutations (no crash; equal bytes whenever both decoders accept).  usage: inflate_diff [seed] [iterations]
#include "../../tidal-wave_b200/csrc/tw_inflate.h"
#include <zlib.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <random>
#include <chrono>
int main(int argc, char **argv) {
    std::mt19937 rng(argc > 1 ? atoi(argv[1]) : 1);
    int bad = 0, cases = 0;
    const int iters = argc > 2 ? atoi(argv[2]) : 3000;
    for (int it = 0; it < iters; it++) {
        size_t n = 1 + rng() % (it % 10 == 0 ? 300000 : 5000);
        std::vector<uint8_t> data(n);
        int kind = rng() % 6;
        for (size_t i = 0; i < n; i++) {
            switch (kind) {
                case 0: data[i] = rng(); break;
                case 1: data[i] = (rng() % 16 == 0) ? rng() : (i ? data[i - 1] : 0); break;           // runs
                case 2: data[i] = i >= 3 && rng() % 32 ? data[i - 3] : rng(); break;                   // period 3
                case 3: data[i] = (uint8_t)(rng() % 4); break;                                          // small alphabet
                case 4: data[i] = i >= 1000 && rng() % 64 ? data[i - 1000 + (rng() % 3)] : rng() % 7; break; // long distances
                default: data[i] = (uint8_t)(i / 7 + (rng() % 3 == 0)); break;
            }
        }
        int level = rng() % 10, strategy = rng() % 5; // Z_DEFAULT_STRATEGY..Z_FIXED
        z_stream zs{}; deflateInit2(&zs, level, Z_DEFLATED, 8 + rng() % 8, 1 + rng() % 9, strategy);
        std::vector<uint8_t> comp(deflateBound(&zs, n) + 64);
        zs.next_in = data.data(); zs.avail_in = n; zs.next_out = comp.data(); zs.avail_out = comp.size();
        // a few full flushes in the middle make extra (stored-empty) blocks
        if (n > 100 && rng() % 3 == 0) { zs.avail_in = n / 2; deflate(&zs, rng() % 2 ? Z_FULL_FLUSH : Z_SYNC_FLUSH); zs.avail_in = n - n / 2; }
        deflate(&zs, Z_FINISH);
        size_t cn = zs.total_out; deflateEnd(&zs);
        std::vector<uint8_t> out(n + 16, 0xAA);
        bool ok = tw_inflate::inflate_exact(comp.data(), cn, out.data(), n);
        cases++;
        if (!ok || memcmp(out.data(), data.data(), n) || out[n] != 0xAA) { bad++; printf("MISMATCH it %d n %zu kind %d level %d strat %d ok %d\n", it, n, kind, level, strategy, ok); }
        // asking for fewer bytes than the stream holds = "data past the last scanline": still true, prefix equal
        if (n > 10) {
            size_t m = rng() % n;
            std::vector<uint8_t> o2(m + 16, 0xBB);
            bool ok2 = tw_inflate::inflate_exact(comp.data(), cn, o2.data(), m);
            if (!ok2 || memcmp(o2.data(), data.data(), m) || o2[m] != 0xBB) { bad++; printf("PREFIX MISMATCH it %d\n", it); }
            // asking for more than it holds must fail; so must a truncated stream
            std::vector<uint8_t> o3(n + 64);
            if (tw_inflate::inflate_exact(comp.data(), cn, o3.data(), n + 5)) { bad++; printf("LONG ACCEPTED it %d\n", it); }
            if (cn > 8 && n > 50 && kind == 0 && tw_inflate::inflate_exact(comp.data(), cn / 2, o3.data(), n)) { bad++; printf("TRUNCATED ACCEPTED it %d\n", it); }
        }
        // mutations must not crash (run under ASan) and must agree with zlib whenever we say true
        for (int mu = 0; mu < 4 && cn > 4; mu++) {
            std::vector<uint8_t> c2(comp.begin(), comp.begin() + cn);
            c2[2 + rng() % (cn - 2)] ^= 1 << (rng() % 8);
            std::vector<uint8_t> o4(n + 16), o5(n + 16);
            bool okm = tw_inflate::inflate_exact(c2.data(), cn, o4.data(), n);
            if (okm) {
                z_stream is{}; inflateInit(&is); is.next_in = c2.data(); is.avail_in = cn; is.next_out = o5.data(); is.avail_out = n;
                int zr = inflate(&is, Z_NO_FLUSH); bool full = is.avail_out == 0; inflateEnd(&is);
                bool zok = full && (zr == Z_OK || zr == Z_STREAM_END || zr == Z_BUF_ERROR);
                if (zok && memcmp(o4.data(), o5.data(), n)) { bad++; printf("MUTATION DIFF it %d\n", it); }
                // (zlib may reject on the Adler-32 what we accept: not a difference in the bytes)
            }
        }
    }
    printf("cases %d bad %d\n", cases, bad);
    return bad != 0;
}
