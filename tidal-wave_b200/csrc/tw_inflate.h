// tw_inflate.h -- a zlib-stream (RFC 1950 / 1951) decompressor for the PNG leg of tw_decode_gray (cv::imread,
// /root/reference/src/opticalflow.cpp:37,44): inflate is 55-75 % of a screenshot PNG's decode time with the system
// zlib, and file -> result throughput is bound by the host decode.  Written from the RFCs for this one use: the whole input and the
// exact output size are known up front, so there is no streaming state -- a 64-bit bit buffer refilled eight bytes at a time,
// two-level decode tables (11 / 8 root bits), run and word-wise match copies, no Adler-32 pass (libpng does not fail a decoded image on
// it either).  Output past `out_cap` is dropped (PNG: data after the last scanline is ignored).  Any inconsistency returns false and
// the caller falls back to zlib for the verdict, so a stream is never rejected on this decoder's word alone.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>

namespace tw_inflate {

constexpr int kLitBits = 11, kDistBits = 8;
// decode-table entry: bits 0..4 = bits to consume at this level, 5..7 = kind, 8..11 = extra bits (or sub-table index bits), 16..31 = value
enum : uint32_t { kInvalid = 0, kLiteral = 1, kLength = 2, kEnd = 3, kSub = 4, kDist = 5 };
inline uint32_t entry(uint32_t kind, uint32_t nbits, uint32_t extra, uint32_t value) { return nbits | (kind << 5) | (extra << 8) | (value << 16); }

struct Tables {
    uint32_t lit[(1 << kLitBits) + 1024];  // root + sub-tables (zlib's ENOUGH bound for 286 symbols / 15 bits is below this)
    uint32_t dist[(1 << kDistBits) + 512];
};

inline uint32_t bitrev(uint32_t c, int n)
{
    uint32_t r = 0;
    for (int i = 0; i < n; i++) { r = (r << 1) | (c & 1); c >>= 1; }
    return r;
}

// Canonical Huffman decode table from code lengths (RFC 1951 3.2.2).  sym_entry(sym, len_at_this_level) makes the leaf entry.
// Rejects over-subscribed sets and incomplete ones, except -- as zlib does -- a single code of length 1.
template <class F>
bool build(const uint8_t *lens, int n, uint32_t *table, int root_bits, int table_cap, F sym_entry)
{
    int count[16] = {0};
    for (int i = 0; i < n; i++) count[lens[i]]++;
    int maxlen = 15;
    while (maxlen > 0 && !count[maxlen]) maxlen--;
    for (int i = 0; i < (1 << root_bits); i++) table[i] = 0;
    if (maxlen == 0) return true; // no codes: every lookup is invalid
    int left = 1;
    for (int l = 1; l <= 15; l++) {
        left = (left << 1) - count[l];
        if (left < 0) return false; // over-subscribed
    }
    if (left > 0 && maxlen != 1) return false; // incomplete
    uint16_t offs[16], sorted[288 + 32];
    offs[1] = 0;
    for (int l = 1; l < 15; l++) offs[l + 1] = (uint16_t)(offs[l] + count[l]);
    for (int i = 0; i < n; i++)
        if (lens[i]) sorted[offs[lens[i]]++] = (uint16_t)i;
    const uint32_t root_mask = (1u << root_bits) - 1;
    uint32_t code = 0;
    int si = 0, next_free = 1 << root_bits;
    for (int l = 1; l <= maxlen && l <= root_bits; l++) {
        for (int k = 0; k < count[l]; k++, si++, code++) {
            const uint32_t e = sym_entry(sorted[si], l), r = bitrev(code, l);
            for (uint32_t i = r; i <= root_mask; i += 1u << l) table[i] = e;
        }
        code <<= 1;
    }
    // codes longer than the root: sub-tables, one per root prefix (canonical order keeps a prefix's codes together)
    uint32_t cur_prefix = ~0u;
    int sub_start = 0, sub_bits = 0;
    for (int l = root_bits + 1; l <= maxlen; l++) {
        for (int k = 0; k < count[l]; k++, si++, code++) {
            const uint32_t r = bitrev(code, l), prefix = r & root_mask;
            if (prefix != cur_prefix) {
                // size of this sub-table: enough index bits for the longest code that shares the prefix
                sub_bits = l - root_bits;
                int room = 1 << sub_bits, ll = l, used = count[l] - k;
                while (used < room && ll < maxlen) { // the remaining codes of length ll do not fill it: longer codes follow under this prefix
                    room = (room - used) << 1;
                    ll++; sub_bits++;
                    used = count[ll];
                }
                sub_start = next_free;
                next_free += 1 << sub_bits;
                if (next_free > table_cap) return false;
                for (int i = sub_start; i < next_free; i++) table[i] = 0;
                table[prefix] = entry(kSub, (uint32_t)root_bits, (uint32_t)sub_bits, (uint32_t)sub_start);
                cur_prefix = prefix;
            }
            const uint32_t e = sym_entry(sorted[si], l - root_bits);
            for (uint32_t i = r >> root_bits; i < (1u << sub_bits); i += 1u << (l - root_bits)) table[sub_start + i] = e;
        }
        code <<= 1;
    }
    return true;
}

inline uint32_t lit_entry(int sym, int nbits)
{
    static const uint16_t base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    static const uint8_t extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    if (sym < 256) return entry(kLiteral, (uint32_t)nbits, 0, (uint32_t)sym);
    if (sym == 256) return entry(kEnd, (uint32_t)nbits, 0, 0);
    if (sym > 285) return entry(kInvalid, (uint32_t)nbits, 0, 0);
    return entry(kLength, (uint32_t)nbits, extra[sym - 257], base[sym - 257]);
}
inline uint32_t dist_entry(int sym, int nbits)
{
    static const uint16_t base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    static const uint8_t extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    if (sym > 29) return entry(kInvalid, (uint32_t)nbits, 0, 0);
    return entry(kDist, (uint32_t)nbits, extra[sym], base[sym]);
}

struct Bits {
    const uint8_t *in, *end;
    uint64_t buf = 0;
    int cnt = 0; // valid bits in buf; goes negative when the stream is read past its end
    inline void refill()
    {
        if (end - in >= 8) { // bytes that do not fit entirely are ORed in again by the next refill (same bits: harmless)
            uint64_t w;
            memcpy(&w, in, 8);
            buf |= w << cnt;
            in += (63 - cnt) >> 3;
            cnt |= 56;
        } else {
            while (cnt <= 56 && in < end) { buf |= (uint64_t)*in++ << cnt; cnt += 8; }
        }
    }
    inline uint32_t peek(int n) const { return (uint32_t)(buf & ((1ull << n) - 1)); }
    inline void drop(int n) { buf >>= n; cnt -= n; }
    inline uint32_t take(int n) { const uint32_t v = peek(n); drop(n); return v; }
};

// zlib stream `src` -> exactly `out_cap` bytes (true), or false: malformed / truncated / shorter than out_cap.  Little-endian hosts.
inline bool inflate_exact(const uint8_t *src, size_t n, uint8_t *out, size_t out_cap)
{
    if (n < 2 || (src[0] & 15) != 8 || (src[0] >> 4) > 7 || ((src[0] << 8) | src[1]) % 31 != 0 || (src[1] & 0x20)) return false;
    Bits b{src + 2, src + n};
    uint8_t *o = out, *const oend = out + out_cap;
    Tables *T = new Tables; // 14 KB: off the stack of the pool's decoder threads
    struct Free { Tables *t; ~Free() { delete t; } } guard{T};
    bool fixed_built = false, dyn_current = false;
    for (;;) {
        b.refill();
        const uint32_t last = b.take(1), type = b.take(2);
        if (b.cnt < 0) return false;
        if (type == 0) { // stored: back to the byte boundary, LEN / NLEN, raw bytes
            b.drop(b.cnt & 7);
            b.in -= b.cnt >> 3;
            b.buf = 0; b.cnt = 0;
            if (b.end - b.in < 4) return false;
            const uint32_t len = b.in[0] | (b.in[1] << 8), nlen = b.in[2] | (b.in[3] << 8);
            if ((len ^ 0xFFFFu) != nlen) return false;
            b.in += 4;
            if ((size_t)(b.end - b.in) < len) return false;
            const size_t take = len < (size_t)(oend - o) ? len : (size_t)(oend - o);
            memcpy(o, b.in, take);
            o += take; b.in += len;
            if (o == oend) return true;
        } else if (type == 1 || type == 2) {
            if (type == 1) {
                if (!fixed_built || dyn_current) {
                    uint8_t lens[288 + 32];
                    for (int i = 0; i < 144; i++) lens[i] = 8;
                    for (int i = 144; i < 256; i++) lens[i] = 9;
                    for (int i = 256; i < 280; i++) lens[i] = 7;
                    for (int i = 280; i < 288; i++) lens[i] = 8;
                    if (!build(lens, 288, T->lit, kLitBits, (int)(sizeof T->lit / 4), lit_entry)) return false;
                    for (int i = 0; i < 32; i++) lens[i] = 5;
                    if (!build(lens, 32, T->dist, kDistBits, (int)(sizeof T->dist / 4), dist_entry)) return false;
                    fixed_built = true; dyn_current = false;
                }
            } else {
                const int hlit = (int)b.take(5) + 257, hdist = (int)b.take(5) + 1, hclen = (int)b.take(4) + 4;
                if (hlit > 286 || hdist > 30) return false;
                static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
                uint8_t cl[19] = {0};
                b.refill();
                for (int i = 0; i < hclen; i++) { if (b.cnt < 3) b.refill(); cl[order[i]] = (uint8_t)b.take(3); }
                if (b.cnt < 0) return false;
                uint32_t pre[128];
                if (!build(cl, 19, pre, 7, 128, [](int sym, int nbits) { return entry(kLiteral, (uint32_t)nbits, 0, (uint32_t)sym); })) return false;
                uint8_t lens[286 + 30 + 138];
                int i = 0;
                while (i < hlit + hdist) {
                    b.refill();
                    const uint32_t e = pre[b.peek(7)];
                    if ((e >> 5 & 7) != kLiteral) return false;
                    b.drop((int)(e & 31));
                    const int sym = (int)(e >> 16);
                    if (sym < 16) { lens[i++] = (uint8_t)sym; }
                    else {
                        int rep, val = 0;
                        if (sym == 16) { if (i == 0) return false; val = lens[i - 1]; rep = 3 + (int)b.take(2); }
                        else if (sym == 17) rep = 3 + (int)b.take(3);
                        else rep = 11 + (int)b.take(7);
                        if (i + rep > hlit + hdist) return false;
                        while (rep--) lens[i++] = (uint8_t)val;
                    }
                    if (b.cnt < 0) return false;
                }
                if (lens[256] == 0) return false; // no end-of-block code
                if (!build(lens, hlit, T->lit, kLitBits, (int)(sizeof T->lit / 4), lit_entry)) return false;
                if (!build(lens + hlit, hdist, T->dist, kDistBits, (int)(sizeof T->dist / 4), dist_entry)) return false;
                dyn_current = true;
            }
            // ---- the block's symbols ----
            for (;;) {
                b.refill();
                uint32_t e = T->lit[b.peek(kLitBits)];
                if ((e >> 5 & 7) == kSub) { b.drop(kLitBits); e = T->lit[(e >> 16) + b.peek((int)(e >> 8 & 15))]; }
                b.drop((int)(e & 31));
                if (b.cnt < 0) return false; // read past the end of the input
                uint32_t kind = e >> 5 & 7;
                if (kind == kLiteral) {
                    if (o == oend) return true; // data past the last scanline
                    *o++ = (uint8_t)(e >> 16);
                    // a second and third literal without another refill: <= 15 bits each, >= 56 - 15 were left
                    e = T->lit[b.peek(kLitBits)];
                    if ((e >> 5 & 7) != kLiteral || b.cnt < 40) continue;
                    b.drop((int)(e & 31));
                    if (o == oend) return true;
                    *o++ = (uint8_t)(e >> 16);
                    e = T->lit[b.peek(kLitBits)];
                    if ((e >> 5 & 7) != kLiteral || b.cnt < 40) continue;
                    b.drop((int)(e & 31));
                    if (o == oend) return true;
                    *o++ = (uint8_t)(e >> 16);
                    continue;
                }
                if (kind == kEnd) break;
                if (kind != kLength) return false;
                const uint32_t len = (e >> 16) + b.take((int)(e >> 8 & 15));
                uint32_t d = T->dist[b.peek(kDistBits)];
                if ((d >> 5 & 7) == kSub) { b.drop(kDistBits); d = T->dist[(d >> 16) + b.peek((int)(d >> 8 & 15))]; }
                b.drop((int)(d & 31));
                if ((d >> 5 & 7) != kDist) return false;
                const uint32_t dist = (d >> 16) + b.take((int)(d >> 8 & 15));
                if (b.cnt < 0 || dist > (size_t)(o - out)) return false;
                size_t m = len < (size_t)(oend - o) ? len : (size_t)(oend - o);
                const uint8_t *s = o - dist;
                if (dist == 1) { memset(o, *s, m); o += m; }
                else {
                    // a short period (RGB runs: distance 3 or 4) is doubled until it spans a word: after `period` bytes have been
                    // copied the data is periodic with twice the period, seen from the same source
                    size_t period = dist;
                    while (period < 8 && m) {
                        const size_t c = period < m ? period : m;
                        for (size_t i = 0; i < c; i++) o[i] = s[i];
                        o += c; m -= c; period *= 2;
                    }
                    if (m) {
                        if ((size_t)(oend - o) >= m + 8) { // eight bytes at a time; may write up to 7 bytes past the match, inside the buffer
                            uint8_t *q = o;
                            o += m;
                            do { memcpy(q, s, 8); q += 8; s += 8; } while (q < o);
                        } else {
                            while (m--) *o++ = *s++;
                        }
                    }
                }
                if (o == oend) return true;
            }
        } else {
            return false;
        }
        if (o == oend) return true;
        if (last) return false; // the stream ended before out_cap bytes were produced
    }
}

} // namespace tw_inflate
