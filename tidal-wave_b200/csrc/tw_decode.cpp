// tw_decode.cpp -- cv::imread(path, IMREAD_GRAYSCALE) for the formats this build can decode bit-exactly
// (/root/reference/src/opticalflow.cpp:37,44).
//
//   PNG  non-interlaced, 8 bits per sample, colour types 0 (gray), 2 (RGB), 3 (palette), 4 (gray+alpha), 6 (RGBA).
//        zlib inflate + the five PNG row filters; alpha is dropped; colour -> gray exactly as OpenCV's decoder does
//        it through libpng (png_set_rgb_to_gray with 0.299 / 0.587): gray = (9797*R + 19234*G + 3737*B) >> 15,
//        truncating (libpng turns 0.299 / 0.587 into the integers 29900*32768/100000 and 58700*32768/100000).  Verified bit-identical to cv2.imread(..., IMREAD_GRAYSCALE) on the reference's PNG fixtures.
//   PGM  binary P5, maxval 255.
//   JPEG baseline and progressive Huffman, gray or YCbCr: tw_jpeg.cpp (luma-only decode with libjpeg's ISLOW inverse DCT,
//        bit-identical to cv2.imread(..., IMREAD_GRAYSCALE) on the reference's progressive scenario1 fixtures).
//   Anything else: TW_BAD_IMAGE_FORMAT, which the callers report as "Can't open <path>" like a failed imread.
#include "../../include/tidalwave_b200.h"

#include <cstdlib>
#include <cstring>
#include <vector>
#include <zlib.h>

namespace {

uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

int paeth(int a, int b, int c)
{
    int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

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

int decode_png(const uint8_t *b, size_t n, uint8_t *out, size_t cap, int *w, int *h)
{
    size_t pos = 8;
    int W = 0, H = 0, depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat, plte;
    bool have_ihdr = false;
    while (pos + 12 <= n) {
        uint32_t len = be32(b + pos);
        const uint8_t *type = b + pos + 4, *data = b + pos + 8;
        if (pos + 12 + (size_t)len > n) return TW_BAD_IMAGE_FORMAT;
        if (!memcmp(type, "IHDR", 4) && len >= 13) {
            W = (int)be32(data); H = (int)be32(data + 4); depth = data[8]; ctype = data[9]; interlace = data[12];
            have_ihdr = true;
        } else if (!memcmp(type, "PLTE", 4)) {
            plte.assign(data, data + len);
        } else if (!memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), data, data + len);
        } else if (!memcmp(type, "IEND", 4)) {
            break;
        }
        pos += 12 + (size_t)len;
    }
    if (!have_ihdr || W <= 0 || H <= 0 || depth != 8 || interlace != 0) return TW_BAD_IMAGE_FORMAT;
    int ch;
    switch (ctype) {
        case 0: ch = 1; break;
        case 2: ch = 3; break;
        case 3: ch = 1; break;
        case 4: ch = 2; break;
        case 6: ch = 4; break;
        default: return TW_BAD_IMAGE_FORMAT;
    }
    if (ctype == 3 && plte.size() < 3) return TW_BAD_IMAGE_FORMAT;
    *w = W; *h = H;
    if (!out) return TW_OK;
    if (cap < (size_t)W * H) return TW_BAD_PARAMETER;
    const size_t stride = (size_t)W * ch;
    std::vector<uint8_t> raw((stride + 1) * H);
    uLongf rawlen = (uLongf)raw.size();
    if (uncompress(raw.data(), &rawlen, idat.data(), (uLong)idat.size()) != Z_OK || rawlen != raw.size()) return TW_BAD_IMAGE_FORMAT;
    std::vector<uint8_t> prev(stride, 0), cur(stride);
    for (int y = 0; y < H; y++) {
        const uint8_t *line = raw.data() + (stride + 1) * y;
        const int f = line[0];
        line++;
        for (size_t i = 0; i < stride; i++) {
            const int a = i >= (size_t)ch ? cur[i - ch] : 0, up = prev[i], c = i >= (size_t)ch ? prev[i - ch] : 0;
            int v;
            switch (f) {
                case 0: v = line[i]; break;
                case 1: v = line[i] + a; break;
                case 2: v = line[i] + up; break;
                case 3: v = line[i] + ((a + up) >> 1); break;
                case 4: v = line[i] + paeth(a, up, c); break;
                default: return TW_BAD_IMAGE_FORMAT;
            }
            cur[i] = (uint8_t)v;
        }
        uint8_t *o = out + (size_t)y * W;
        for (int x = 0; x < W; x++) {
            int r, g, bl;
            if (ctype == 0 || ctype == 4) { o[x] = cur[(size_t)x * ch]; continue; }
            if (ctype == 3) {
                size_t idx = (size_t)cur[x] * 3;
                if (idx + 2 >= plte.size()) { r = g = bl = 0; } else { r = plte[idx]; g = plte[idx + 1]; bl = plte[idx + 2]; }
            } else {
                r = cur[(size_t)x * ch]; g = cur[(size_t)x * ch + 1]; bl = cur[(size_t)x * ch + 2];
            }
            // libpng rgb_to_gray with OpenCV's coefficients; gray pixels (r == g == b) pass through unchanged
            o[x] = (r == g && g == bl) ? (uint8_t)r : (uint8_t)((9797 * r + 19234 * g + 3737 * bl) >> 15);
        }
        prev.swap(cur);
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
