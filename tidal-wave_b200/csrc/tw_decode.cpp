// tw_decode.cpp -- cv::imread(path, IMREAD_GRAYSCALE) for the formats this build can decode bit-exactly
// (/root/reference/src/opticalflow.cpp:37,44).
//
//   PNG  every colour type (gray, RGB, palette, gray+alpha, RGBA) at every bit depth the format allows (1 / 2 / 4 / 8 / 16),
//        Adam7-interlaced or not; critical-chunk CRCs are checked.
//        inflate (tw_inflate.h: whole-buffer decoder with a 64-bit bit buffer and two-level tables; zlib decides whatever it
//        declines) + the five PNG row filters; alpha / tRNS is dropped; gray below 8 bits is scaled to 0..255; 16-bit samples
//        keep their high byte (16-bit RGB is weighted at 16 bits first, with rounding); colour -> gray exactly as OpenCV's decoder does
//        it through libpng (png_set_rgb_to_gray with 0.299 / 0.587): gray = (9797*R + 19234*G + 3737*B) >> 15,
//        truncating (libpng turns 0.299 / 0.587 into the integers 29900*32768/100000 and 58700*32768/100000).  Verified bit-identical to cv2.imread(..., IMREAD_GRAYSCALE) on the reference's PNG fixtures.
//   PGM  binary P5, maxval 255.
//   JPEG baseline and progressive Huffman, gray or YCbCr: tw_jpeg.cpp (luma-only decode with libjpeg's ISLOW inverse DCT,
//        bit-identical to cv2.imread(..., IMREAD_GRAYSCALE) on the reference's progressive scenario1 fixtures).
//   Anything else: TW_BAD_IMAGE_FORMAT, which the callers report as "Can't open <path>" like a failed imread.
#include "../../include/tidalwave_b200.h"
#include "tw_inflate.h"

#include <cstdlib>
#include <cstring>
#include <vector>
#include <zlib.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace {

uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

int decode_pgm(const uint8_t *b, size_t n, uint8_t *out, size_t cap, int *w, int *h)
{
    size_t pos = 2;
    int vals[3], got = 0;
    while (got < 3 && pos < n) {
        while (pos < n && (b[pos] == ' ' || b[pos] == '\n' || b[pos] == '\r' || b[pos] == '\t')) pos++;
        if (pos < n && b[pos] == '#') { while (pos < n && b[pos] != '\n') pos++; continue; }
        int v = 0, digits = 0;
        while (pos < n && b[pos] >= '0' && b[pos] <= '9') {
            if (++digits > 7) return TW_BAD_IMAGE_FORMAT; // bounded: no signed overflow on a long digit string
            v = v * 10 + (b[pos] - '0'); pos++;
        }
        if (!digits) return TW_BAD_IMAGE_FORMAT;
        vals[got++] = v;
    }
    if (got < 3 || vals[2] != 255 || vals[0] <= 0 || vals[1] <= 0) return TW_BAD_IMAGE_FORMAT;
    pos++; // single whitespace after maxval
    size_t need = (size_t)vals[0] * vals[1];
    if (pos + need > n) return TW_BAD_IMAGE_FORMAT;
    *w = vals[0]; *h = vals[1];
    if (out) {
        if (cap < need) return TW_BAD_PARAMETER;
        memcpy(out, b + pos, need);
    }
    return TW_OK;
}

// ---- PNG ----
// Filter type 1 (Sub) is a running sum per byte lane along the row.  Sixteen (B = 1, 4) or twelve (B = 3) bytes at a time: a
// log-step prefix sum inside the register (byte shifts by whole pixels), plus the carried last pixel of the previous group
// broadcast to every pixel -- byte additions wrap exactly like the scalar form.  Returns the number of bytes done.
#if defined(__x86_64__)
__attribute__((target("ssse3"))) size_t unfilter_sub_ssse3(uint8_t *line, size_t n, int B)
{
    size_t i = 0;
    if (B == 4) {
        __m128i carry = _mm_setzero_si128();
        for (; i + 16 <= n; i += 16) {
            __m128i x = _mm_loadu_si128(reinterpret_cast<const __m128i *>(line + i));
            x = _mm_add_epi8(x, _mm_slli_si128(x, 4));
            x = _mm_add_epi8(x, _mm_slli_si128(x, 8));
            x = _mm_add_epi8(x, carry);
            _mm_storeu_si128(reinterpret_cast<__m128i *>(line + i), x);
            carry = _mm_shuffle_epi32(x, 0xFF);
        }
    } else if (B == 1) {
        __m128i carry = _mm_setzero_si128();
        const __m128i last = _mm_set1_epi8(15);
        for (; i + 16 <= n; i += 16) {
            __m128i x = _mm_loadu_si128(reinterpret_cast<const __m128i *>(line + i));
            x = _mm_add_epi8(x, _mm_slli_si128(x, 1));
            x = _mm_add_epi8(x, _mm_slli_si128(x, 2));
            x = _mm_add_epi8(x, _mm_slli_si128(x, 4));
            x = _mm_add_epi8(x, _mm_slli_si128(x, 8));
            x = _mm_add_epi8(x, carry);
            _mm_storeu_si128(reinterpret_cast<__m128i *>(line + i), x);
            carry = _mm_shuffle_epi8(x, last);
        }
    } else if (B == 3) {
        __m128i carry = _mm_setzero_si128();
        const __m128i keep12 = _mm_setr_epi8(-1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, 0, 0, 0, 0);
        const __m128i bcast = _mm_setr_epi8(9, 10, 11, 9, 10, 11, 9, 10, 11, 9, 10, 11, -1, -1, -1, -1);
        for (; i + 16 <= n; i += 12) { // four pixels; the 16-byte load reads 4 bytes of the next group, the stores write 12
            __m128i x = _mm_and_si128(_mm_loadu_si128(reinterpret_cast<const __m128i *>(line + i)), keep12);
            x = _mm_add_epi8(x, _mm_slli_si128(x, 3));
            x = _mm_add_epi8(x, _mm_slli_si128(x, 6));
            x = _mm_add_epi8(x, carry);
            _mm_storel_epi64(reinterpret_cast<__m128i *>(line + i), x);
            const int hi = _mm_cvtsi128_si32(_mm_srli_si128(x, 8));
            memcpy(line + i + 8, &hi, 4);
            carry = _mm_shuffle_epi8(x, bcast);
        }
    }
    return i;
}
// Filter type 4 (Paeth) is a serial chain per byte, but on screenshot content it mostly degenerates, and both degenerate forms can
// be recognised for a whole group of pixels (16 bytes, 12 for B = 3) before any of its bytes is computed:
//   (A) the row above is locally flat, b == c at every byte of the group: then p = a, |p - a| = 0 and the predictor is a -- the group
//       is a running sum, done in the register like filter type 1;
//   (B) the row above continues into this one: speculate predictor b (out = x + b) and check that every left neighbour then equals
//       its upper-left, a == c -- there |p - b| = 0 and the predictor is b unless |p - a| = 0 too, where a == b == c anyway.
// Groups that are neither run the scalar predictor.  Same bytes as the scalar form by construction.
template <int B> __attribute__((target("ssse3"))) void unfilter_paeth_ssse3(uint8_t *line, const uint8_t *prev, size_t n)
{
    const size_t G = B == 3 ? 12 : 16;
    const __m128i keep = B == 3 ? _mm_setr_epi8(-1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, 0, 0, 0, 0) : _mm_set1_epi8(-1);
    const int want = B == 3 ? 0x0FFF : 0xFFFF;
    const __m128i lowB = B == 1 ? _mm_setr_epi8(-1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0)
                       : B == 3 ? _mm_setr_epi8(-1, -1, -1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0)
                                : _mm_setr_epi8(-1, -1, -1, -1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0);
    const __m128i bcast = B == 1 ? _mm_setzero_si128()
                        : B == 3 ? _mm_setr_epi8(0, 1, 2, 0, 1, 2, 0, 1, 2, 0, 1, 2, -1, -1, -1, -1)
                                 : _mm_setr_epi8(0, 1, 2, 3, 0, 1, 2, 3, 0, 1, 2, 3, 0, 1, 2, 3);
    for (int k = 0; k < B && (size_t)k < n; k++) line[k] = (uint8_t)(line[k] + prev[k]); // first pixel: a = c = 0, the predictor is b
    size_t i = (size_t)B;
    __m128i lastpix = _mm_setzero_si128(); // the pixel left of the group in the low B bytes
    bool have_last = false;
    while (i + 16 <= n) {
        if (!have_last) {
            int w = 0;
            memcpy(&w, line + i - B, (size_t)B);
            lastpix = _mm_cvtsi32_si128(w);
        }
        const __m128i P = _mm_loadu_si128(reinterpret_cast<const __m128i *>(prev + i));
        const __m128i Pm = _mm_loadu_si128(reinterpret_cast<const __m128i *>(prev + i - B));
        const __m128i X = _mm_and_si128(_mm_loadu_si128(reinterpret_cast<const __m128i *>(line + i)), keep);
        __m128i out;
        bool simd = true;
        if ((_mm_movemask_epi8(_mm_cmpeq_epi8(P, Pm)) & want) == want) { // (A) running sum
            out = _mm_add_epi8(X, _mm_slli_si128(X, B == 1 ? 1 : B == 3 ? 3 : 4));
            out = _mm_add_epi8(out, _mm_slli_si128(out, B == 1 ? 2 : B == 3 ? 6 : 8));
            if (B == 1) { out = _mm_add_epi8(out, _mm_slli_si128(out, 4)); out = _mm_add_epi8(out, _mm_slli_si128(out, 8)); }
            out = _mm_add_epi8(out, _mm_shuffle_epi8(lastpix, bcast));
        } else {
            out = _mm_add_epi8(X, _mm_and_si128(P, keep)); // (B) speculate predictor b
            const __m128i left = _mm_or_si128(B == 1 ? _mm_slli_si128(out, 1) : B == 3 ? _mm_slli_si128(out, 3) : _mm_slli_si128(out, 4), _mm_and_si128(lastpix, lowB));
            simd = (_mm_movemask_epi8(_mm_cmpeq_epi8(left, Pm)) & want) == want;
        }
        if (simd) {
            if (B == 3) {
                _mm_storel_epi64(reinterpret_cast<__m128i *>(line + i), out);
                const int hi = _mm_cvtsi128_si32(_mm_srli_si128(out, 8));
                memcpy(line + i + 8, &hi, 4);
                lastpix = _mm_srli_si128(out, 9);
            } else {
                _mm_storeu_si128(reinterpret_cast<__m128i *>(line + i), out);
                lastpix = B == 1 ? _mm_srli_si128(out, 15) : _mm_srli_si128(out, 12);
            }
            have_last = true;
        } else {
            for (size_t j = i; j < i + G; j++) {
                const int av = line[j - B], bv = prev[j], cv = prev[j - B];
                const int pa0 = bv - cv, pb0 = av - cv;
                const int pa = abs(pa0), pb = abs(pb0), pc = abs(pa0 + pb0);
                line[j] = (uint8_t)(line[j] + ((pa <= pb && pa <= pc) ? av : (pb <= pc ? bv : cv)));
            }
            have_last = false;
        }
        i += G;
    }
    for (; i < n; i++) {
        const int av = line[i - B], bv = prev[i], cv = prev[i - B];
        const int pa0 = bv - cv, pb0 = av - cv;
        const int pa = abs(pa0), pb = abs(pb0), pc = abs(pa0 + pb0);
        line[i] = (uint8_t)(line[i] + ((pa <= pb && pa <= pc) ? av : (pb <= pc ? bv : cv)));
    }
}
const bool kHaveSsse3Filter = (__builtin_cpu_init(), __builtin_cpu_supports("ssse3"));
#endif

// Undo the filter of one scanline in place (PNG spec 9.2).  B = bytes per pixel (1 for sub-byte depths); prev = the unfiltered
// line above, or nullptr for the first line of an image / pass (all zero).  The left neighbour a and the upper-left c of each of
// the B byte lanes live in registers: going through line[i - B] would put a store-to-load forward on every byte's critical path.
template <int B> bool unfilter_line(int f, uint8_t *line, const uint8_t *prev, size_t n)
{
    uint8_t a[B], c[B];
    for (int k = 0; k < B; k++) a[k] = c[k] = 0;
    size_t i = 0;
    switch (f) {
        case 0: return true;
        case 1:
#if defined(__x86_64__)
            if ((B == 1 || B == 3 || B == 4) && kHaveSsse3Filter && n >= 32) {
                i = unfilter_sub_ssse3(line, n, B); // a multiple of B bytes
                for (int k = 0; k < B; k++) a[k] = line[i - B + k];
            }
#endif
            for (; i + B <= n; i += B)
                for (int k = 0; k < B; k++) line[i + k] = a[k] = (uint8_t)(line[i + k] + a[k]);
            for (int k = 0; i < n; i++, k++) line[i] = (uint8_t)(line[i] + a[k]);
            return true;
        case 2:
            if (prev) for (; i < n; i++) line[i] = (uint8_t)(line[i] + prev[i]);
            return true;
        case 3:
            for (; i + B <= n; i += B)
                for (int k = 0; k < B; k++) line[i + k] = a[k] = (uint8_t)(line[i + k] + ((a[k] + (prev ? prev[i + k] : 0)) >> 1));
            for (int k = 0; i < n; i++, k++) line[i] = (uint8_t)(line[i] + ((a[k] + (prev ? prev[i] : 0)) >> 1));
            return true;
        case 4:
            if (!prev) return unfilter_line<B>(1, line, prev, n); // b = c = 0: the predictor is a
#if defined(__x86_64__)
            if ((B == 1 || B == 3 || B == 4) && kHaveSsse3Filter && n >= 32) { unfilter_paeth_ssse3<(B == 1 || B == 3 || B == 4) ? B : 1>(line, prev, n); return true; }
#endif
            for (; i + B <= n; i += B)
                for (int k = 0; k < B; k++) {
                    const int av = a[k], bv = prev[i + k], cv = c[k];
                    const int pa0 = bv - cv, pb0 = av - cv; // p - a, p - b with p = a + b - c
                    const int pa = abs(pa0), pb = abs(pb0), pc = abs(pa0 + pb0);
                    const int pred = (pa <= pb && pa <= pc) ? av : (pb <= pc ? bv : cv);
                    c[k] = (uint8_t)bv;
                    line[i + k] = a[k] = (uint8_t)(line[i + k] + pred);
                }
            for (int k = 0; i < n; i++, k++) { // a partial last pixel cannot occur in a valid stream; keep the bytes defined
                const int av = a[k], bv = prev[i], cv = c[k];
                const int pa0 = bv - cv, pb0 = av - cv;
                const int pa = abs(pa0), pb = abs(pb0), pc = abs(pa0 + pb0);
                line[i] = (uint8_t)(line[i] + ((pa <= pb && pa <= pc) ? av : (pb <= pc ? bv : cv)));
            }
            return true;
        default: return false;
    }
}

// `rows` scanlines, each = filter byte + rowbytes of data
__attribute__((noinline)) bool unfilter(uint8_t *data, int rows, size_t rowbytes, int bpp)
{
    const uint8_t *prev = nullptr;
    for (int y = 0; y < rows; y++) {
        uint8_t *line = data + (rowbytes + 1) * (size_t)y;
        const int f = *line++;
        bool ok;
        switch (bpp) {
            case 1: ok = unfilter_line<1>(f, line, prev, rowbytes); break;
            case 2: ok = unfilter_line<2>(f, line, prev, rowbytes); break;
            case 3: ok = unfilter_line<3>(f, line, prev, rowbytes); break;
            case 4: ok = unfilter_line<4>(f, line, prev, rowbytes); break;
            case 6: ok = unfilter_line<6>(f, line, prev, rowbytes); break;
            case 8: ok = unfilter_line<8>(f, line, prev, rowbytes); break;
            default: return false;
        }
        if (!ok) return false;
        prev = line;
    }
    return true;
}

// (9797 R + 19234 G + 3737 B) >> 15 for four pixels at a time: one shuffle spreads (R, G) into 16-bit pairs for a multiply-add
// against (9797, 19234), another B against 3737 -- exact integers (< 2^23), so the bytes are those of the scalar form.  Returns the
// number of pixels done (the caller finishes the row; 16-byte loads stay inside it).
#if defined(__x86_64__)
__attribute__((target("ssse3"))) int rgb_to_gray_ssse3(const uint8_t *row, int npix, uint8_t *dst, int ch)
{
    const __m128i rg3 = _mm_setr_epi8(0, -1, 1, -1, 3, -1, 4, -1, 6, -1, 7, -1, 9, -1, 10, -1), b3 = _mm_setr_epi8(2, -1, -1, -1, 5, -1, -1, -1, 8, -1, -1, -1, 11, -1, -1, -1);
    const __m128i rg4 = _mm_setr_epi8(0, -1, 1, -1, 4, -1, 5, -1, 8, -1, 9, -1, 12, -1, 13, -1), b4 = _mm_setr_epi8(2, -1, -1, -1, 6, -1, -1, -1, 10, -1, -1, -1, 14, -1, -1, -1);
    const __m128i mrg = ch == 3 ? rg3 : rg4, mb = ch == 3 ? b3 : b4;
    const __m128i krg = _mm_setr_epi16(9797, 19234, 9797, 19234, 9797, 19234, 9797, 19234), kb = _mm_setr_epi16(3737, 0, 3737, 0, 3737, 0, 3737, 0);
    int x = 0;
    for (; (size_t)(x + 4) * ch + (ch == 3 ? 4 : 0) <= (size_t)npix * ch; x += 4) { // a 16-byte load of 12 needs 4 more bytes in the row
        const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i *>(row + (size_t)x * ch));
        const __m128i s = _mm_add_epi32(_mm_madd_epi16(_mm_shuffle_epi8(v, mrg), krg), _mm_madd_epi16(_mm_shuffle_epi8(v, mb), kb));
        const __m128i g = _mm_srli_epi32(s, 15);
        const __m128i p = _mm_packus_epi16(_mm_packs_epi32(g, g), g);
        const int out4 = _mm_cvtsi128_si32(p);
        memcpy(dst + x, &out4, 4);
    }
    return x;
}
const bool kHaveSsse3 = (__builtin_cpu_init(), __builtin_cpu_supports("ssse3"));
inline int rgb_to_gray_simd(const uint8_t *row, int npix, uint8_t *dst, int ch) { return kHaveSsse3 ? rgb_to_gray_ssse3(row, npix, dst, ch) : 0; }
#else
inline int rgb_to_gray_simd(const uint8_t *, int, uint8_t *, int) { return 0; }
#endif

struct PngFormat {
    int depth, ctype, ch;
    const std::vector<uint8_t> *plte;
};

// One unfiltered scanline of npix pixels -> gray bytes at dst[j * step], as cv::imread(IMREAD_GRAYSCALE) produces them through libpng:
// palette -> RGB, gray 1 / 2 / 4 bits scaled to 8, alpha (and tRNS) dropped, RGB -> gray with OpenCV's coefficients
// ((9797 R + 19234 G + 3737 B) >> 15 at 8 bits; at 16 bits with + 16384 rounding, libpng's 16-bit form), and 16-bit samples
// reduced to their high byte afterwards (png_set_strip_16).
__attribute__((noinline)) void to_gray(const PngFormat &F, const uint8_t *row, int npix, uint8_t *dst, size_t step)
{
    const int depth = F.depth, ch = F.ch;
    if (depth == 8) {
        switch (F.ctype) {
            case 0: case 4:
                for (int x = 0; x < npix; x++) dst[x * step] = row[(size_t)x * ch];
                break;
            case 3: {
                const std::vector<uint8_t> &pl = *F.plte;
                for (int x = 0; x < npix; x++) {
                    const size_t idx = (size_t)row[x] * 3;
                    int r = 0, g = 0, bl = 0;
                    if (idx + 2 < pl.size()) { r = pl[idx]; g = pl[idx + 1]; bl = pl[idx + 2]; }
                    dst[x * step] = (uint8_t)((9797 * r + 19234 * g + 3737 * bl) >> 15);
                }
                break;
            }
            default:
                if (step == 1 && ch == 3) {
                    int x = rgb_to_gray_simd(row, npix, dst, 3);
                    for (; x < npix; x++) dst[x] = (uint8_t)((9797 * row[3 * x] + 19234 * row[3 * x + 1] + 3737 * row[3 * x + 2]) >> 15);
                } else if (step == 1) {
                    int x = rgb_to_gray_simd(row, npix, dst, 4);
                    for (; x < npix; x++) dst[x] = (uint8_t)((9797 * row[4 * x] + 19234 * row[4 * x + 1] + 3737 * row[4 * x + 2]) >> 15);
                } else {
                    for (int x = 0; x < npix; x++) {
                        const uint8_t *q = row + (size_t)x * ch;
                        dst[x * step] = (uint8_t)((9797 * q[0] + 19234 * q[1] + 3737 * q[2]) >> 15);
                    }
                }
        }
    } else if (depth == 16) {
        if (F.ctype == 0 || F.ctype == 4) {
            for (int x = 0; x < npix; x++) dst[x * step] = row[(size_t)x * ch * 2];
        } else {
            for (int x = 0; x < npix; x++) {
                const uint8_t *q = row + (size_t)x * ch * 2;
                const uint32_t r = (uint32_t)(q[0] << 8 | q[1]), g = (uint32_t)(q[2] << 8 | q[3]), bl = (uint32_t)(q[4] << 8 | q[5]);
                dst[x * step] = (uint8_t)(((9797u * r + 19234u * g + 3737u * bl + 16384u) >> 15) >> 8);
            }
        }
    } else { // 1, 2, 4 bits: gray (scaled to 0..255) or palette index, most significant bits first
        const int per = 8 / depth, mask = (1 << depth) - 1, scale = 255 / mask;
        const std::vector<uint8_t> *pl = F.ctype == 3 ? F.plte : nullptr;
        for (int x = 0; x < npix; x++) {
            const int v = (row[x / per] >> ((per - 1 - x % per) * depth)) & mask;
            if (!pl) { dst[x * step] = (uint8_t)(v * scale); continue; }
            const size_t idx = (size_t)v * 3;
            int r = 0, g = 0, bl = 0;
            if (idx + 2 < pl->size()) { r = (*pl)[idx]; g = (*pl)[idx + 1]; bl = (*pl)[idx + 2]; }
            dst[x * step] = (uint8_t)((9797 * r + 19234 * g + 3737 * bl) >> 15);
        }
    }
}

// The reduced images of the stream: the whole image, or Adam7's seven passes (PNG spec 8.2); empty passes carry no data.
struct Pass { int x0, y0, dx, dy, pw, ph; size_t rowbytes; };
__attribute__((noinline)) int png_passes(int W, int H, int bits, bool interlace, Pass *passes, size_t *total)
{
    static const int kAdam7[7][4] = {{0, 0, 8, 8}, {4, 0, 8, 8}, {0, 4, 4, 8}, {2, 0, 4, 4}, {0, 2, 2, 4}, {1, 0, 2, 2}, {0, 1, 1, 2}};
    int np = 0;
    *total = 0;
    if (!interlace) {
        passes[0] = Pass{0, 0, 1, 1, W, H, ((size_t)W * bits + 7) / 8};
        *total = (passes[0].rowbytes + 1) * (size_t)H;
        return 1;
    }
    for (int i = 0; i < 7; i++) {
        const int x0 = kAdam7[i][0], y0 = kAdam7[i][1], dx = kAdam7[i][2], dy = kAdam7[i][3];
        const int pw = W > x0 ? (W - x0 + dx - 1) / dx : 0, ph = H > y0 ? (H - y0 + dy - 1) / dy : 0;
        if (pw > 0 && ph > 0) {
            passes[np] = Pass{x0, y0, dx, dy, pw, ph, ((size_t)pw * bits + 7) / 8};
            *total += (passes[np].rowbytes + 1) * (size_t)ph;
            np++;
        }
    }
    return np;
}

// inflate the concatenated IDAT chunks into `raw`; data past the last scanline is ignored (libpng: a warning), too little fails.
// tw_inflate.h does the work; whatever it declines goes through zlib, which has the last word.
bool inflate_idat(const std::vector<std::pair<const uint8_t *, uint32_t>> &idat, std::vector<uint8_t> &raw)
{
    static const bool use_fast = !getenv("TW_PNG_ZLIB"); // measurement / bisecting switch
    if (use_fast && !idat.empty()) {
        if (idat.size() == 1) {
            if (tw_inflate::inflate_exact(idat[0].first, idat[0].second, raw.data(), raw.size())) return true;
        } else {
            size_t total = 0;
            for (auto &c : idat) total += c.second;
            std::vector<uint8_t> joined(total);
            size_t o = 0;
            for (auto &c : idat) { memcpy(joined.data() + o, c.first, c.second); o += c.second; }
            if (tw_inflate::inflate_exact(joined.data(), total, raw.data(), raw.size())) return true;
        }
    }
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (inflateInit(&zs) != Z_OK) return false;
    zs.next_out = raw.data();
    zs.avail_out = (uInt)raw.size();
    int zr = Z_OK;
    size_t i = 0;
    while (i < idat.size() && zs.avail_out > 0 && zr == Z_OK) {
        zs.next_in = const_cast<Bytef *>(idat[i].first);
        zs.avail_in = idat[i].second;
        zr = inflate(&zs, Z_NO_FLUSH);
        i++;
    }
    const bool full = zs.avail_out == 0;
    inflateEnd(&zs);
    return full && (zr == Z_OK || zr == Z_STREAM_END || zr == Z_BUF_ERROR);
}

int decode_png(const uint8_t *b, size_t n, uint8_t *out, size_t cap, int *w, int *h)
{
    size_t pos = 8;
    int W = 0, H = 0, depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> plte;
    std::vector<std::pair<const uint8_t *, uint32_t>> idat;
    size_t idat_bytes = 0;
    bool have_ihdr = false;
    while (pos + 12 <= n) {
        const uint32_t len = be32(b + pos);
        const uint8_t *type = b + pos + 4, *data = b + pos + 8;
        if (pos + 12 + (size_t)len > n) return TW_BAD_IMAGE_FORMAT;
        const bool ihdr = !memcmp(type, "IHDR", 4), is_plte = !memcmp(type, "PLTE", 4), is_idat = !memcmp(type, "IDAT", 4);
        if (out && (ihdr || is_plte || is_idat) && // libpng fails the read on a critical chunk whose CRC is wrong
            (uint32_t)crc32(crc32(0L, type, 4), data, len) != be32(data + len)) return TW_BAD_IMAGE_FORMAT;
        if (ihdr && len >= 13) {
            if (data[10] != 0 || data[11] != 0) return TW_BAD_IMAGE_FORMAT; // compression / filter method
            W = (int)be32(data); H = (int)be32(data + 4); depth = data[8]; ctype = data[9]; interlace = data[12];
            have_ihdr = true;
        } else if (is_plte) {
            plte.assign(data, data + len);
        } else if (is_idat) {
            idat.emplace_back(data, len);
            idat_bytes += len;
        } else if (!memcmp(type, "IEND", 4)) {
            break;
        }
        pos += 12 + (size_t)len;
    }
    if (!have_ihdr || W <= 0 || H <= 0 || interlace > 1) return TW_BAD_IMAGE_FORMAT;
    int ch;
    bool ok_depth;
    switch (ctype) {
        case 0: ch = 1; ok_depth = depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16; break;
        case 2: ch = 3; ok_depth = depth == 8 || depth == 16; break;
        case 3: ch = 1; ok_depth = depth == 1 || depth == 2 || depth == 4 || depth == 8; break;
        case 4: ch = 2; ok_depth = depth == 8 || depth == 16; break;
        case 6: ch = 4; ok_depth = depth == 8 || depth == 16; break;
        default: return TW_BAD_IMAGE_FORMAT;
    }
    if (!ok_depth || (ctype == 3 && plte.size() < 3)) return TW_BAD_IMAGE_FORMAT;
    *w = W; *h = H;
    if (!out) return TW_OK;
    if (cap < (size_t)W * H) return TW_BAD_PARAMETER;
    const int bits = ch * depth, bpp = bits >= 8 ? bits / 8 : 1;
    Pass passes[7];
    size_t total = 0;
    const int np = png_passes(W, H, bits, interlace != 0, passes, &total);
    std::vector<uint8_t> raw(total);
    if (!inflate_idat(idat, raw)) return TW_BAD_IMAGE_FORMAT;
    (void)idat_bytes;
    const PngFormat F{depth, ctype, ch, &plte};
    uint8_t *data = raw.data();
    for (int i = 0; i < np; i++) {
        const Pass &P = passes[i];
        if (!unfilter(data, P.ph, P.rowbytes, bpp)) return TW_BAD_IMAGE_FORMAT;
        for (int y = 0; y < P.ph; y++)
            to_gray(F, data + (P.rowbytes + 1) * (size_t)y + 1, P.pw, out + (size_t)(P.y0 + y * P.dy) * W + P.x0, (size_t)P.dx);
        data += (P.rowbytes + 1) * P.ph;
    }
    return TW_OK;
}

} // namespace

int tw_decode_jpeg_gray(const uint8_t *bytes, size_t n, uint8_t *out, size_t cap, int *w, int *h); // tw_jpeg.cpp

extern "C" int tw_decode_gray(const uint8_t *bytes, size_t n, uint8_t *out, size_t cap, int *w, int *h)
{
    if (!bytes || !w || !h) return TW_BAD_PARAMETER;
    static const uint8_t png_sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
    try { // nothing may propagate across the C ABI (a hostile header can ask for gigabytes: std::bad_alloc)
        if (n >= 8 && !memcmp(bytes, png_sig, 8)) return decode_png(bytes, n, out, cap, w, h);
        if (n >= 2 && bytes[0] == 'P' && bytes[1] == '5') return decode_pgm(bytes, n, out, cap, w, h);
        if (n >= 2 && bytes[0] == 0xFF && bytes[1] == 0xD8) return tw_decode_jpeg_gray(bytes, n, out, cap, w, h);
    } catch (...) {
        return TW_BAD_IMAGE_FORMAT;
    }
    return TW_BAD_IMAGE_FORMAT;
}
