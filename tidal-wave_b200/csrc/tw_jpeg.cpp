// tw_jpeg.cpp -- the JPEG leg of cv::imread(path, IMREAD_GRAYSCALE) (/root/reference/src/opticalflow.cpp:37,44; SURVEY
// row f-1): the reference's scenario1 fixtures are progressive JPEGs.
//
// OpenCV's JPEG reader asks libjpeg for JCS_GRAYSCALE output, which for a YCbCr (or gray) file means: entropy-decode, then
// run the inverse DCT of the LUMA component only and copy it out -- chroma is never touched.  Bit-exactness with
// cv2.imread therefore needs exactly (a) the coefficients (Huffman, baseline or progressive with successive
// approximation, restart intervals) and (b) libjpeg's default "ISLOW" integer inverse DCT (Loeffler-Ligtenberg-Moschytz,
// 13-bit constants, 2 extra bits after the column pass) with its wrap-around range-limit table.  Both are restated here
// from the published algorithm (ITU-T T.81 Annex F/G; IJG jidctint.c / jdphuff.c describe the same arithmetic); no
// libjpeg source or header is used.  Block smoothing never applies to complete files (all coefficients reach Al = 0).
//
// Supported: 8-bit Huffman SOF0 / SOF1 / SOF2, 1 component or 3 components (YCbCr: JFIF, or Adobe transform 1, or
// component ids 1,2,3) with the luma component at the maximum sampling factors (4:4:4, 4:2:2, 4:2:0, 4:4:0, ...).
// Not supported (-> TW_BAD_IMAGE_FORMAT): arithmetic coding, 12-bit, lossless, CMYK / RGB-coded files.
// EXIF orientation is ignored, like the reference's OpenCV 2.4.9 does (newer OpenCV builds rotate on imread).
#include "../../include/tidalwave_b200.h"

#include <cstdint>
#include <cstring>
#if defined(__x86_64__)
#include <immintrin.h>
#endif
#include <vector>

namespace {

const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

constexpr int kLookBits = 10;

struct Huff {
    bool present = false;
    int mincode[17], maxcode[18], valptr[17];
    uint8_t vals[256];
    uint16_t look[1 << kLookBits]; // (length << 8) | symbol for codes of at most kLookBits bits, 0 = longer code
    // AC tables: where the code AND the magnitude bits that follow it fit in kLookBits bits, the decoded coefficient itself:
    // (value << 8) | (run << 4) | total bits; 0 = take the two-step path (value in [-128, 127], never 0 for a coefficient)
    int16_t fast_ac[1 << kLookBits];
    void build(const uint8_t *bits /* [1..16] */, const uint8_t *v, int n)
    {
        memset(vals, 0, sizeof vals);
        memcpy(vals, v, n);
        memset(look, 0, sizeof look);
        int code = 0, k = 0;
        for (int l = 1; l <= 16; l++) {
            valptr[l] = k;
            mincode[l] = code;
            if (l <= kLookBits)
                for (int i = 0; i < bits[l]; i++) {
                    const int first = (code + i) << (kLookBits - l);
                    if (first + (1 << (kLookBits - l)) > (1 << kLookBits)) break; // over-subscribed table: the slow path handles it
                    for (int j = 0; j < (1 << (kLookBits - l)); j++) look[first + j] = (uint16_t)((l << 8) | vals[(k + i) & 255]);
                }
            k += bits[l];
            code += bits[l];
            maxcode[l] = bits[l] ? code - 1 : -1;
            code <<= 1;
        }
        maxcode[17] = 0x7fffffff;
        for (int i = 0; i < (1 << kLookBits); i++) {
            fast_ac[i] = 0;
            const int e = look[i], len = e >> 8, run = (e >> 4) & 15, mag = e & 15;
            if (!e || !mag || len + mag > kLookBits) continue;
            int v = ((i << len) & ((1 << kLookBits) - 1)) >> (kLookBits - mag); // the mag bits after the code
            if (v < (1 << (mag - 1))) v += 1 - (1 << mag);                     // T.81 F.2.2.1 EXTEND
            if (v >= -128 && v <= 127) fast_ac[i] = (int16_t)(v * 256 + run * 16 + len + mag);
        }
        present = true;
    }
};

struct Component {
    int id = 0, h = 1, v = 1, tq = 0;
    int wblk = 0, hblk = 0;     // blocks covering the component's own size (non-interleaved scans)
    int wpad = 0, hpad = 0;     // blocks padded to whole MCUs (interleaved scans, storage)
    int dc_tbl = 0, ac_tbl = 0; // current scan
    int last_dc = 0;
    std::vector<int16_t> coef;  // [hpad][wpad][64], natural (row-major) order -- only kept for the luma component
};

const uint16_t kEndianProbe0 = 1;
const bool kLittleEndianHost = *reinterpret_cast<const uint8_t *>(&kEndianProbe0) == 1;

struct BitReader {
    const uint8_t *p, *end;
    uint64_t buf = 0; // bits left-aligned
    int cnt = 0;
    bool hit_marker = false;
    BitReader(const uint8_t *b, const uint8_t *e) : p(b), end(e) {}
    void fill()
    {
        if (!hit_marker && end - p >= 8 && cnt <= 56) { // eight bytes at once when none of them is 0xFF (stuffing / marker)
            uint64_t w;
            memcpy(&w, p, 8);
            w = __builtin_bswap64(w); // little-endian host (checked by kLittleEndian below; big-endian takes the byte loop)
            const uint64_t x = ~w;
            if (kLittleEndianHost && !((x - 0x0101010101010101ull) & ~x & 0x8080808080808080ull)) {
                const int nb = (64 - cnt) >> 3;
                const uint64_t m = nb == 8 ? ~0ull : ~(~0ull >> (nb * 8));
                buf |= (w & m) >> cnt;
                cnt += nb * 8;
                p += nb;
                return;
            }
        }
        while (cnt <= 56) {
            int c = 0;
            if (!hit_marker && p < end) {
                c = *p;
                if (c == 0xFF) {
                    if (p + 1 < end && p[1] == 0x00) p += 2;
                    else { hit_marker = true; c = 0; } // a marker: feed zeros from here on, like libjpeg does
                } else p++;
            }
            buf |= (uint64_t)c << (56 - cnt);
            cnt += 8;
        }
    }
    int bits(int n)
    {
        if (n == 0) return 0;
        if (cnt < n) fill();
        int v = (int)(buf >> (64 - n));
        buf <<= n;
        cnt -= n;
        return v;
    }
    int bit() { return bits(1); }
    int peek(int n) { if (cnt < n) fill(); return (int)(buf >> (64 - n)); }
    void skip(int n) { buf <<= n; cnt -= n; }
    // byte-align and consume an RSTn marker if one is next
    void restart()
    {
        buf = 0; cnt = 0; hit_marker = false;
        while (p + 1 < end && !(p[0] == 0xFF && p[1] >= 0xD0 && p[1] <= 0xD7)) {
            if (p[0] == 0xFF && p[1] != 0x00 && p[1] != 0xFF) return; // some other marker: leave it to the caller
            p++;
        }
        if (p + 1 < end) p += 2;
    }
};

inline int extend(int v, int s) { return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v; }

int decode_symbol(BitReader &br, const Huff &h)
{
    const int e = h.look[br.peek(kLookBits)];
    if (e) { br.skip(e >> 8); return e & 255; }
    const int code16 = br.peek(16); // canonical codes: the first length whose largest code is not below the prefix
    for (int l = 1; l <= 16; l++) {
        const int code = code16 >> (16 - l);
        if (code <= h.maxcode[l]) {
            br.skip(l);
            return h.vals[(h.valptr[l] + code - h.mincode[l]) & 255];
        }
    }
    br.skip(16);
    return 0; // corrupt data: libjpeg warns and returns 0
}

// ---- inverse DCT, "ISLOW" ----
const int kConstBits = 13, kPass1Bits = 2;
inline int64_t descale(int64_t x, int n) { return (x + ((int64_t)1 << (n - 1))) >> n; }

struct RangeLimit {
    uint8_t t[1024];
    RangeLimit()
    {
        for (int i = 0; i < 1024; i++) {
            if (i < 128) t[i] = (uint8_t)(i + 128);
            else if (i < 512) t[i] = 255;
            else if (i < 896) t[i] = 0;
            else t[i] = (uint8_t)(i - 896);
        }
    }
};
const RangeLimit kRange;
const uint16_t kEndianProbe = 1;
const bool kLittleEndian = *reinterpret_cast<const uint8_t *>(&kEndianProbe) == 1;

void idct_islow(const int16_t *in, const uint16_t *q, uint8_t *out, int stride)
{
    const int64_t F0_298 = 2446, F0_390 = 3196, F0_541 = 4433, F0_765 = 6270, F0_899 = 7373, F1_175 = 9633, F1_501 = 12299,
                  F1_847 = 15137, F1_961 = 16069, F2_053 = 16819, F2_562 = 20995, F3_072 = 25172;
    int64_t ws[64];
    for (int c = 0; c < 8; c++) { // pass 1: columns
        const int16_t *ip = in + c;
        const uint16_t *qp = q + c;
        int64_t *wp = ws + c;
        if (!ip[8] && !ip[16] && !ip[24] && !ip[32] && !ip[40] && !ip[48] && !ip[56]) {
            const int64_t dc = (int64_t)((int32_t)ip[0] * qp[0]) * (1 << kPass1Bits);
            for (int r = 0; r < 8; r++) wp[8 * r] = dc;
            continue;
        }
        int64_t z2 = ip[16] * qp[16], z3 = ip[48] * qp[48];
        int64_t z1 = (z2 + z3) * F0_541;
        int64_t tmp2 = z1 + z3 * (-F1_847), tmp3 = z1 + z2 * F0_765;
        z2 = ip[0] * qp[0]; z3 = ip[32] * qp[32];
        int64_t tmp0 = (z2 + z3) * (1 << kConstBits), tmp1 = (z2 - z3) * (1 << kConstBits);
        const int64_t tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
        tmp0 = ip[56] * qp[56]; tmp1 = ip[40] * qp[40]; tmp2 = ip[24] * qp[24]; tmp3 = ip[8] * qp[8];
        z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
        int64_t z4 = tmp1 + tmp3, z5 = (z3 + z4) * F1_175;
        tmp0 *= F0_298; tmp1 *= F2_053; tmp2 *= F3_072; tmp3 *= F1_501;
        z1 *= -F0_899; z2 *= -F2_562; z3 *= -F1_961; z4 *= -F0_390;
        z3 += z5; z4 += z5;
        tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
        const int n = kConstBits - kPass1Bits;
        wp[0] = descale(tmp10 + tmp3, n);  wp[56] = descale(tmp10 - tmp3, n);
        wp[8] = descale(tmp11 + tmp2, n);  wp[48] = descale(tmp11 - tmp2, n);
        wp[16] = descale(tmp12 + tmp1, n); wp[40] = descale(tmp12 - tmp1, n);
        wp[24] = descale(tmp13 + tmp0, n); wp[32] = descale(tmp13 - tmp0, n);
    }
    for (int r = 0; r < 8; r++) { // pass 2: rows
        const int64_t *wp = ws + 8 * r;
        uint8_t *o = out + (size_t)r * stride;
        const int n = kConstBits + kPass1Bits + 3;
        if (!wp[1] && !wp[2] && !wp[3] && !wp[4] && !wp[5] && !wp[6] && !wp[7]) {
            const uint8_t dc = kRange.t[descale(wp[0], kPass1Bits + 3) & 1023];
            for (int c = 0; c < 8; c++) o[c] = dc;
            continue;
        }
        int64_t z2 = wp[2], z3 = wp[6];
        int64_t z1 = (z2 + z3) * F0_541;
        int64_t tmp2 = z1 + z3 * (-F1_847), tmp3 = z1 + z2 * F0_765;
        int64_t tmp0 = (wp[0] + wp[4]) * (1 << kConstBits), tmp1 = (wp[0] - wp[4]) * (1 << kConstBits);
        const int64_t tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
        tmp0 = wp[7]; tmp1 = wp[5]; tmp2 = wp[3]; tmp3 = wp[1];
        z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
        int64_t z4 = tmp1 + tmp3, z5 = (z3 + z4) * F1_175;
        tmp0 *= F0_298; tmp1 *= F2_053; tmp2 *= F3_072; tmp3 *= F1_501;
        z1 *= -F0_899; z2 *= -F2_562; z3 *= -F1_961; z4 *= -F0_390;
        z3 += z5; z4 += z5;
        tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
        o[0] = kRange.t[descale(tmp10 + tmp3, n) & 1023]; o[7] = kRange.t[descale(tmp10 - tmp3, n) & 1023];
        o[1] = kRange.t[descale(tmp11 + tmp2, n) & 1023]; o[6] = kRange.t[descale(tmp11 - tmp2, n) & 1023];
        o[2] = kRange.t[descale(tmp12 + tmp1, n) & 1023]; o[5] = kRange.t[descale(tmp12 - tmp1, n) & 1023];
        o[3] = kRange.t[descale(tmp13 + tmp0, n) & 1023]; o[4] = kRange.t[descale(tmp13 - tmp0, n) & 1023];
    }
}

// The same inverse DCT on eight 32-bit lanes (AVX2; pass 1: lane = column, pass 2: lane = row).  Integer arithmetic is exact
// as long as nothing overflows, so wherever every dequantised coefficient and every pass-1 output lies within +-8191 (any
// block a real encoder produces: pass-1 outputs of 8-bit samples stay below 4096 * 1.5) the int32 lanes hold exactly the
// int64 values of idct_islow above; otherwise -- and on a host without AVX2 -- the function declines and the caller runs
// the scalar form.
#if defined(__x86_64__)
#define TW_JPEG_AVX2 1
#define TW_AVX2 __attribute__((target("avx2")))

TW_AVX2 inline void islow_lanes(const __m256i *in, __m256i *out, int shift)
{
#define K(c) _mm256_set1_epi32(c)
#define MUL(a, c) _mm256_mullo_epi32(a, K(c))
#define ADD _mm256_add_epi32
#define SUB _mm256_sub_epi32
    __m256i z2 = in[2], z3 = in[6];
    __m256i z1 = MUL(ADD(z2, z3), 4433);
    __m256i tmp2 = SUB(z1, MUL(z3, 15137)), tmp3 = ADD(z1, MUL(z2, 6270));
    __m256i tmp0 = _mm256_slli_epi32(ADD(in[0], in[4]), kConstBits), tmp1 = _mm256_slli_epi32(SUB(in[0], in[4]), kConstBits);
    const __m256i tmp10 = ADD(tmp0, tmp3), tmp13 = SUB(tmp0, tmp3), tmp11 = ADD(tmp1, tmp2), tmp12 = SUB(tmp1, tmp2);
    tmp0 = in[7]; tmp1 = in[5]; tmp2 = in[3]; tmp3 = in[1];
    z1 = ADD(tmp0, tmp3); z2 = ADD(tmp1, tmp2); z3 = ADD(tmp0, tmp2);
    __m256i z4 = ADD(tmp1, tmp3);
    const __m256i z5 = MUL(ADD(z3, z4), 9633);
    tmp0 = MUL(tmp0, 2446); tmp1 = MUL(tmp1, 16819); tmp2 = MUL(tmp2, 25172); tmp3 = MUL(tmp3, 12299);
    z1 = MUL(z1, -7373); z2 = MUL(z2, -20995); z3 = ADD(MUL(z3, -16069), z5); z4 = ADD(MUL(z4, -3196), z5);
    tmp0 = ADD(tmp0, ADD(z1, z3)); tmp1 = ADD(tmp1, ADD(z2, z4)); tmp2 = ADD(tmp2, ADD(z2, z3)); tmp3 = ADD(tmp3, ADD(z1, z4));
    const __m256i rnd = K(1 << (shift - 1));
    const __m128i sh = _mm_cvtsi32_si128(shift);
#define OUT(a, op, b) _mm256_sra_epi32(ADD(op(a, b), rnd), sh)
    out[0] = OUT(tmp10, ADD, tmp3); out[7] = OUT(tmp10, SUB, tmp3);
    out[1] = OUT(tmp11, ADD, tmp2); out[6] = OUT(tmp11, SUB, tmp2);
    out[2] = OUT(tmp12, ADD, tmp1); out[5] = OUT(tmp12, SUB, tmp1);
    out[3] = OUT(tmp13, ADD, tmp0); out[4] = OUT(tmp13, SUB, tmp0);
#undef OUT
#undef K
#undef MUL
#undef ADD
#undef SUB
}

TW_AVX2 inline void transpose8(const __m256i *r, __m256i *o)
{
    const __m256i t0 = _mm256_unpacklo_epi32(r[0], r[1]), t1 = _mm256_unpackhi_epi32(r[0], r[1]);
    const __m256i t2 = _mm256_unpacklo_epi32(r[2], r[3]), t3 = _mm256_unpackhi_epi32(r[2], r[3]);
    const __m256i t4 = _mm256_unpacklo_epi32(r[4], r[5]), t5 = _mm256_unpackhi_epi32(r[4], r[5]);
    const __m256i t6 = _mm256_unpacklo_epi32(r[6], r[7]), t7 = _mm256_unpackhi_epi32(r[6], r[7]);
    const __m256i u0 = _mm256_unpacklo_epi64(t0, t2), u1 = _mm256_unpackhi_epi64(t0, t2);
    const __m256i u2 = _mm256_unpacklo_epi64(t1, t3), u3 = _mm256_unpackhi_epi64(t1, t3);
    const __m256i u4 = _mm256_unpacklo_epi64(t4, t6), u5 = _mm256_unpackhi_epi64(t4, t6);
    const __m256i u6 = _mm256_unpacklo_epi64(t5, t7), u7 = _mm256_unpackhi_epi64(t5, t7);
    o[0] = _mm256_permute2x128_si256(u0, u4, 0x20); o[1] = _mm256_permute2x128_si256(u1, u5, 0x20);
    o[2] = _mm256_permute2x128_si256(u2, u6, 0x20); o[3] = _mm256_permute2x128_si256(u3, u7, 0x20);
    o[4] = _mm256_permute2x128_si256(u0, u4, 0x31); o[5] = _mm256_permute2x128_si256(u1, u5, 0x31);
    o[6] = _mm256_permute2x128_si256(u2, u6, 0x31); o[7] = _mm256_permute2x128_si256(u3, u7, 0x31);
}

TW_AVX2 bool idct_islow_avx2(const int16_t *in, const uint16_t *q, uint8_t *out, int stride)
{
    __m256i a[8], w[8];
    const __m256i big = _mm256_set1_epi32(~8191);
    __m256i m = _mm256_setzero_si256();
    for (int k = 0; k < 8; k++) {
        a[k] = _mm256_mullo_epi32(_mm256_cvtepi16_epi32(_mm_loadu_si128((const __m128i *)(in + 8 * k))),
                                  _mm256_cvtepu16_epi32(_mm_loadu_si128((const __m128i *)(q + 8 * k))));
        m = _mm256_or_si256(m, _mm256_abs_epi32(a[k]));
    }
    if (!_mm256_testz_si256(m, big)) return false;
    islow_lanes(a, w, kConstBits - kPass1Bits); // w[r] = row r of the workspace, lane = column
    m = _mm256_abs_epi32(w[0]);
    for (int r = 1; r < 8; r++) m = _mm256_or_si256(m, _mm256_abs_epi32(w[r]));
    if (!_mm256_testz_si256(m, big)) return false;
    transpose8(w, a);                           // a[k] = workspace column k, lane = row
    islow_lanes(a, w, kConstBits + kPass1Bits + 3); // w[c] = output column c, lane = row
    transpose8(w, a);                           // a[r] = output row r
    // the wrap-around range-limit table = ((v mod 1024, as a signed 10-bit number) + 128) clamped to a byte
    const __m256i c128 = _mm256_set1_epi32(128);
    for (int r = 0; r < 8; r += 2) {
        const __m256i v0 = _mm256_add_epi32(_mm256_srai_epi32(_mm256_slli_epi32(a[r], 22), 22), c128);
        const __m256i v1 = _mm256_add_epi32(_mm256_srai_epi32(_mm256_slli_epi32(a[r + 1], 22), 22), c128);
        const __m256i p16 = _mm256_packs_epi32(v0, v1);    // per 128-bit half: row r cols 0-3 | row r+1 cols 0-3 ; cols 4-7 likewise
        const __m256i p8 = _mm256_packus_epi16(p16, p16);  // bytes: [r c0-3][r+1 c0-3][..][..] | [r c4-7][r+1 c4-7][..][..]
        const __m128i lo = _mm256_castsi256_si128(p8), hi = _mm256_extracti128_si256(p8, 1);
        const __m128i rows = _mm_unpacklo_epi32(lo, hi);   // [r c0-3][r c4-7][r+1 c0-3][r+1 c4-7]
        _mm_storel_epi64((__m128i *)(out + (size_t)r * stride), rows);
        _mm_storel_epi64((__m128i *)(out + (size_t)(r + 1) * stride), _mm_srli_si128(rows, 8));
    }
    return true;
}
const bool kHaveAvx2 = (__builtin_cpu_init(), __builtin_cpu_supports("avx2")); // explicit init: this runs as a static constructor
inline bool idct_islow_lanes(const int16_t *in, const uint16_t *q, uint8_t *out, int stride)
{
    return kHaveAvx2 && idct_islow_avx2(in, q, out, stride);
}
#else
inline bool idct_islow_lanes(const int16_t *, const uint16_t *, uint8_t *, int) { return false; }
#endif

struct Decoder {
    const uint8_t *b;
    size_t n;
    int W = 0, H = 0, ncomp = 0, maxh = 1, maxv = 1;
    bool progressive = false, have_sof = false;
    int restart_interval = 0;
    int adobe_transform = -1;
    uint16_t qt[4][64];
    bool have_qt[4] = {false, false, false, false};
    Huff dc[4], ac[4];
    Component comp[4];
    int luma = 0;
    int eobrun = 0;
    uint8_t *out_ = nullptr;   // baseline files: every luma block goes through the inverse DCT as soon as it is decoded
    bool luma_scanned = false; // (no coefficient store; a block is coded exactly once)

    uint16_t be16(size_t pos) const { return (uint16_t)((b[pos] << 8) | b[pos + 1]); }

    int16_t *block(Component &c, int bx, int by) { return c.coef.data() + ((size_t)by * c.wpad + bx) * 64; }

    // ---- per-block entropy decoding; `blk` == nullptr: decode and discard (chroma) ----
    // returns the zigzag index of the last coefficient stored (0 = DC only)
    int baseline_block(BitReader &br, Component &c, int16_t *blk)
    {
        int s = decode_symbol(br, dc[c.dc_tbl]);
        int diff = s ? extend(br.bits(s), s) : 0;
        c.last_dc += diff;
        if (blk) blk[0] = (int16_t)c.last_dc;
        const Huff &h = ac[c.ac_tbl];
        int last = 0;
        for (int k = 1; k < 64; k++) {
            if (br.cnt < 32) br.fill(); // one refill covers a code (<= 16 bits) and its magnitude bits (<= 15)
            const int idx = (int)(br.buf >> (64 - kLookBits)), f = h.fast_ac[idx];
            if (f) { // code + magnitude resolved by one lookup
                k += (f >> 4) & 15;
                br.skip(f & 15);
                if (blk && k < 64) { blk[kZigzag[k]] = (int16_t)(f >> 8); last = k; }
                continue;
            }
            int rs = decode_symbol(br, h), r = rs >> 4;
            s = rs & 15;
            if (s) {
                k += r;
                int v = extend(br.bits(s), s);
                if (blk && k < 64) { blk[kZigzag[k]] = (int16_t)v; last = k; }
            } else {
                if (r != 15) break;
                k += 15;
            }
        }
        return last;
    }
    void dc_first(BitReader &br, Component &c, int16_t *blk, int Al)
    {
        int s = decode_symbol(br, dc[c.dc_tbl]);
        int diff = s ? extend(br.bits(s), s) : 0;
        c.last_dc += diff;
        if (blk) blk[0] = (int16_t)((uint32_t)c.last_dc << Al);
    }
    void dc_refine(BitReader &br, int16_t *blk, int Al)
    {
        int bit = br.bit();
        if (blk && bit) blk[0] |= (int16_t)(1 << Al);
    }
    void ac_first(BitReader &br, Component &c, int16_t *blk, int Ss, int Se, int Al)
    {
        if (eobrun > 0) { eobrun--; return; }
        const Huff &h = ac[c.ac_tbl];
        for (int k = Ss; k <= Se; k++) {
            int rs = decode_symbol(br, h), r = rs >> 4, s = rs & 15;
            if (s) {
                k += r;
                int v = extend(br.bits(s), s);
                if (k <= 63) blk[kZigzag[k]] = (int16_t)((uint32_t)v << Al);
            } else {
                if (r == 15) k += 15;
                else {
                    eobrun = 1 << r;
                    if (r) eobrun += br.bits(r);
                    eobrun--;
                    break;
                }
            }
        }
    }
    void ac_refine(BitReader &br, Component &c, int16_t *blk, int Ss, int Se, int Al)
    {
        const int p1 = 1 << Al, m1 = -(1 << Al);
        const Huff &h = ac[c.ac_tbl];
        int k = Ss;
        if (eobrun == 0) {
            for (; k <= Se; k++) {
                int rs = decode_symbol(br, h), r = rs >> 4, s = rs & 15;
                if (s) {
                    s = br.bit() ? p1 : m1; // the size must be 1
                } else if (r != 15) {
                    eobrun = 1 << r;
                    if (r) eobrun += br.bits(r);
                    break; // end of band: the remaining coefficients are handled below
                }
                do { // advance over already-nonzero coefficients (each takes a correction bit) and r zeroes
                    int16_t *cf = blk + kZigzag[k];
                    if (*cf != 0) {
                        if (br.bit() && (*cf & p1) == 0) *cf = (int16_t)(*cf >= 0 ? *cf + p1 : *cf + m1);
                    } else if (--r < 0) break;
                    k++;
                } while (k <= Se);
                if (s && k <= 63) blk[kZigzag[k]] = (int16_t)s;
            }
        }
        if (eobrun > 0) {
            for (; k <= Se; k++) {
                int16_t *cf = blk + kZigzag[k];
                if (*cf != 0 && br.bit() && (*cf & p1) == 0) *cf = (int16_t)(*cf >= 0 ? *cf + p1 : *cf + m1);
            }
            eobrun--;
        }
    }

    int decode_scan(size_t &pos)
    {
        const int len = be16(pos);
        if (pos + len > n || len < 6) return TW_BAD_IMAGE_FORMAT;
        const int ns = b[pos + 2];
        if (ns < 1 || ns > ncomp || len != 6 + 2 * ns) return TW_BAD_IMAGE_FORMAT;
        Component *sc[4];
        for (int i = 0; i < ns; i++) {
            const int id = b[pos + 3 + 2 * i], tb = b[pos + 4 + 2 * i];
            sc[i] = nullptr;
            for (int c = 0; c < ncomp; c++) if (comp[c].id == id) sc[i] = &comp[c];
            if (!sc[i]) return TW_BAD_IMAGE_FORMAT;
            sc[i]->dc_tbl = (tb >> 4) & 3; sc[i]->ac_tbl = tb & 3;
        }
        const int Ss = b[pos + 3 + 2 * ns], Se = b[pos + 4 + 2 * ns], Ah = b[pos + 5 + 2 * ns] >> 4, Al = b[pos + 5 + 2 * ns] & 15;
        pos += len;
        // entropy-coded segment: up to the next marker that is neither a stuffed FF00 nor RSTn
        size_t e = pos;
        while (e + 1 < n && !(b[e] == 0xFF && b[e + 1] != 0x00 && !(b[e + 1] >= 0xD0 && b[e + 1] <= 0xD7) && b[e + 1] != 0xFF)) e++;
        if (e + 1 >= n) e = n;
        const size_t seg_begin = pos;
        pos = e;
        if (progressive) {
            if (Ss > Se || Se > 63 || (Ss == 0 && Se != 0) || (Ss > 0 && ns != 1) || Al > 13) return TW_BAD_IMAGE_FORMAT;
        } else if (Ss != 0 || Se != 63 || Ah != 0 || Al != 0) {
            return TW_BAD_IMAGE_FORMAT;
        }
        const bool touches_luma = [&] { for (int i = 0; i < ns; i++) if (sc[i] == &comp[luma]) return true; return false; }();
        if (!touches_luma) return TW_OK; // chroma-only scan: nothing the gray output needs
        for (int i = 0; i < ns; i++) {
            if ((!progressive || Ss == 0) && Ah == 0 && !dc[sc[i]->dc_tbl].present) return TW_BAD_IMAGE_FORMAT;
            if ((!progressive || Ss > 0) && !ac[sc[i]->ac_tbl].present) return TW_BAD_IMAGE_FORMAT;
            sc[i]->last_dc = 0;
        }
        eobrun = 0;
        BitReader br(b + seg_begin, b + e);
        luma_scanned = true;
        const uint16_t *lq = qt[comp[luma].tq];
        if (!progressive && !have_qt[comp[luma].tq]) return TW_BAD_IMAGE_FORMAT;
        auto one_block = [&](Component &c, int bx, int by) {
            if (!progressive) {
                if (&c != &comp[luma] || bx >= c.wblk || by >= c.hblk) { baseline_block(br, c, nullptr); return; }
                alignas(32) int16_t cf[64];
                memset(cf, 0, sizeof cf);
                const int last = baseline_block(br, c, cf);
                emit_block(cf, last == 0, lq, bx, by);
                return;
            }
            int16_t *blk = (&c == &comp[luma]) ? block(c, bx, by) : nullptr;
            if (Ss == 0) { if (Ah == 0) dc_first(br, c, blk, Al); else dc_refine(br, blk, Al); }
            else if (Ah == 0) ac_first(br, c, blk, Ss, Se, Al);
            else ac_refine(br, c, blk, Ss, Se, Al);
        };
        int mcus_x, mcus_y;
        if (ns == 1) { mcus_x = sc[0]->wblk; mcus_y = sc[0]->hblk; }
        else { mcus_x = (W + 8 * maxh - 1) / (8 * maxh); mcus_y = (H + 8 * maxv - 1) / (8 * maxv); }
        int until_restart = restart_interval;
        for (int my = 0; my < mcus_y; my++) {
            for (int mx = 0; mx < mcus_x; mx++) {
                if (restart_interval && until_restart == 0) {
                    br.restart();
                    for (int i = 0; i < ns; i++) sc[i]->last_dc = 0;
                    eobrun = 0;
                    until_restart = restart_interval;
                }
                if (ns == 1) one_block(*sc[0], mx, my);
                else
                    for (int i = 0; i < ns; i++)
                        for (int v = 0; v < sc[i]->v; v++)
                            for (int hh = 0; hh < sc[i]->h; hh++) one_block(*sc[i], mx * sc[i]->h + hh, my * sc[i]->v + v);
                until_restart--;
            }
        }
        return TW_OK;
    }

    int parse(uint8_t *out, size_t cap, int *w, int *h)
    {
        size_t pos = 2;
        bool saw_scan = false;
        while (pos + 4 <= n) {
            if (b[pos] != 0xFF) { pos++; continue; }
            const int m = b[pos + 1];
            if (m == 0xFF) { pos++; continue; }
            pos += 2;
            if (m == 0xD9) break;
            if (m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;
            if (pos + 2 > n) break;
            const int len = be16(pos);
            if (len < 2 || pos + len > n) return TW_BAD_IMAGE_FORMAT;
            const uint8_t *d = b + pos + 2;
            const int dl = len - 2;
            switch (m) {
                case 0xC0: case 0xC1: case 0xC2: {
                    if (have_sof || dl < 6 || d[0] != 8) return TW_BAD_IMAGE_FORMAT;
                    progressive = (m == 0xC2);
                    H = (d[1] << 8) | d[2]; W = (d[3] << 8) | d[4]; ncomp = d[5];
                    if (W <= 0 || H <= 0 || (ncomp != 1 && ncomp != 3) || dl < 6 + 3 * ncomp) return TW_BAD_IMAGE_FORMAT;
                    for (int c = 0; c < ncomp; c++) {
                        comp[c].id = d[6 + 3 * c]; comp[c].h = d[7 + 3 * c] >> 4; comp[c].v = d[7 + 3 * c] & 15; comp[c].tq = d[8 + 3 * c] & 3;
                        if (comp[c].h < 1 || comp[c].h > 4 || comp[c].v < 1 || comp[c].v > 4) return TW_BAD_IMAGE_FORMAT;
                        if (comp[c].h > maxh) maxh = comp[c].h;
                        if (comp[c].v > maxv) maxv = comp[c].v;
                    }
                    have_sof = true;
                    break;
                }
                case 0xC3: case 0xC5: case 0xC6: case 0xC7: case 0xC9: case 0xCA: case 0xCB: case 0xCD: case 0xCE: case 0xCF:
                    return TW_BAD_IMAGE_FORMAT; // lossless / hierarchical / arithmetic
                case 0xC4: {
                    int o = 0;
                    while (o + 17 <= dl) {
                        const int tc = d[o] >> 4, th = d[o] & 15;
                        uint8_t bits[17] = {0};
                        int cnt = 0;
                        for (int i = 1; i <= 16; i++) { bits[i] = d[o + i]; cnt += bits[i]; }
                        if (th > 3 || tc > 1 || cnt > 256 || o + 17 + cnt > dl) return TW_BAD_IMAGE_FORMAT;
                        if (tc == 0) // DC symbols are magnitude categories: libjpeg rejects tables with any > 15
                            for (int i = 0; i < cnt; i++) if (d[o + 17 + i] > 15) return TW_BAD_IMAGE_FORMAT;
                        (tc ? ac[th] : dc[th]).build(bits, d + o + 17, cnt);
                        o += 17 + cnt;
                    }
                    break;
                }
                case 0xDB: {
                    int o = 0;
                    while (o < dl) {
                        const int pq = d[o] >> 4, tq = d[o] & 15;
                        if (tq > 3 || o + 1 + (pq ? 128 : 64) > dl) return TW_BAD_IMAGE_FORMAT;
                        for (int i = 0; i < 64; i++) qt[tq][kZigzag[i]] = pq ? (uint16_t)((d[o + 1 + 2 * i] << 8) | d[o + 2 + 2 * i]) : d[o + 1 + i];
                        have_qt[tq] = true;
                        o += 1 + (pq ? 128 : 64);
                    }
                    break;
                }
                case 0xDD:
                    if (dl < 2) return TW_BAD_IMAGE_FORMAT;
                    restart_interval = (d[0] << 8) | d[1];
                    break;
                case 0xEE:
                    if (dl >= 12 && !memcmp(d, "Adobe", 5)) adobe_transform = d[11];
                    break;
                case 0xDA: {
                    if (!have_sof) return TW_BAD_IMAGE_FORMAT;
                    if (!saw_scan) {
                        int rc = begin_image(out, cap, w, h);
                        if (rc != TW_OK || !out) return rc;
                        saw_scan = true;
                    }
                    int rc = decode_scan(pos);
                    if (rc != TW_OK) return rc;
                    continue; // pos already sits on the next marker
                }
                default: break;
            }
            pos += len;
        }
        if (!have_sof) return TW_BAD_IMAGE_FORMAT;
        if (!saw_scan) {
            int rc = begin_image(out, cap, w, h);
            if (rc != TW_OK || !out) return rc;
            return TW_BAD_IMAGE_FORMAT; // a frame without scans
        }
        return finish(out);
    }

    int begin_image(uint8_t *out, size_t cap, int *w, int *h)
    {
        if (ncomp == 3) {
            // YCbCr only: Adobe transform 0 or component ids 'R','G','B' mean an RGB-coded file
            if (adobe_transform == 0 || (comp[0].id == 'R' && comp[1].id == 'G' && comp[2].id == 'B')) return TW_BAD_IMAGE_FORMAT;
        }
        luma = 0;
        if (comp[0].h != maxh || comp[0].v != maxv) return TW_BAD_IMAGE_FORMAT; // luma would need upsampling
        for (int c = 0; c < ncomp; c++) {
            Component &k = comp[c];
            const int cw = (W * k.h + maxh - 1) / maxh, ch = (H * k.v + maxv - 1) / maxv;
            k.wblk = (cw + 7) / 8; k.hblk = (ch + 7) / 8;
            k.wpad = (W + 8 * maxh - 1) / (8 * maxh) * k.h; k.hpad = (H + 8 * maxv - 1) / (8 * maxv) * k.v;
        }
        *w = W; *h = H;
        if (!out) return TW_OK;
        if (cap < (size_t)W * H) return TW_BAD_PARAMETER;
        out_ = out;
        if (progressive) comp[luma].coef.assign((size_t)comp[luma].wpad * comp[luma].hpad * 64, 0);
        return TW_OK;
    }

    // inverse DCT of one luma block into the output image (clipped at the right / bottom edge)
    void emit_block(const int16_t *blk, bool dc_only, const uint16_t *q, int bx, int by)
    {
        const int ys = by * 8, xs = bx * 8;
        const bool inside = xs + 8 <= W && ys + 8 <= H;
        uint8_t px[64];
        uint8_t *dst = inside ? out_ + (size_t)ys * W + xs : px;
        const int stride = inside ? W : 8;
        if (dc_only) {
            // DC-only block (most of a screenshot): both passes of the inverse DCT take their zero-AC shortcuts, the block is
            // the constant range_limit(descale(dc * q0 << PASS1_BITS, PASS1_BITS + 3)) -- same value, 64 multiplies saved
            const int64_t dc = (int64_t)((int32_t)blk[0] * q[0]) * (1 << kPass1Bits);
            const uint8_t v = kRange.t[descale(dc, kPass1Bits + 3) & 1023];
            for (int y = 0; y < 8; y++) memset(dst + (size_t)y * stride, v, 8);
        } else if (!idct_islow_lanes(blk, q, dst, stride)) {
            idct_islow(blk, q, dst, stride);
        }
        if (!inside)
            for (int y = 0; y < 8 && ys + y < H; y++) {
                const int cols = W - xs < 8 ? W - xs : 8;
                memcpy(out_ + (size_t)(ys + y) * W + xs, px + 8 * y, cols);
            }
    }

    int finish(uint8_t *out)
    {
        Component &c = comp[luma];
        if (!have_qt[c.tq]) return TW_BAD_IMAGE_FORMAT;
        if (!progressive) {
            if (!luma_scanned) memset(out, 128, (size_t)W * H); // no luma scan at all: all-zero coefficients
            return TW_OK;
        }
        for (int by = 0; by < c.hblk; by++) {
            for (int bx = 0; bx < c.wblk; bx++) {
                const int16_t *blk = block(c, bx, by);
                uint64_t ac[16];
                memcpy(ac, blk, sizeof ac);
                uint64_t any = ac[0] & ~(uint64_t)0xFFFF; // little-endian: coefficient 0 is the low half-word
                for (int i = 1; i < 16; i++) any |= ac[i];
                emit_block(blk, !any && kLittleEndian, qt[c.tq], bx, by);
            }
        }
        return TW_OK;
    }
};

} // namespace

// Declared in tw_decode.cpp's dispatch; kept out of the public header (tw_decode_gray is the entry point).
int tw_decode_jpeg_gray(const uint8_t *bytes, size_t n, uint8_t *out, size_t cap, int *w, int *h)
{
    if (n < 4 || bytes[0] != 0xFF || bytes[1] != 0xD8) return TW_BAD_IMAGE_FORMAT;
    Decoder d;
    d.b = bytes; d.n = n;
    return d.parse(out, cap, w, h);
}
