// deflate.cuh -- method id 5, DeflateCompression (advanced_compression.py:71-107): the reference calls
// zlib.compress(data, level=9) / zlib.decompress(data).  zlib itself is not part of the reference's sources; what
// is restated here is the published format -- RFC 1950 (zlib wrapper: CMF / FLG, Adler-32) and RFC 1951 (stored,
// fixed and dynamic Huffman blocks, LZ77 lengths 3..258 at distances 1..32768) -- with two sides:
//   inflate_zlib   decodes ANY conforming zlib stream (so packages of type 5 in files the reference wrote decode
//                  here); errors follow the reference's decompress(): zlib raises -> original_length zero bytes;
//   deflate_fixed  writes a conforming zlib stream with one fixed-Huffman block from a greedy hash-head LZ77 parse
//                  (stock zlib decodes it; it is NOT byte-identical to zlib's own level-9 output, which is why this
//                  codec is reported separately and stays out of the chunk trial).
// Both are serial per stream (one lane); throughput is not a goal of this row.
#pragma once
#include "common.cuh"

__constant__ uint16_t c_len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
__constant__ uint8_t c_len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__constant__ uint16_t c_dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
__constant__ uint8_t c_dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
__constant__ uint8_t c_clen_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

#define ADLER_MOD 65521u

// ---- inflate -------------------------------------------------------------------------------------------------
struct InfBits {
    const uint8_t *p;
    long len, pos;
    uint32_t buf;
    int cnt;
    bool err; // ran past the end of the input
};
__device__ __forceinline__ uint32_t inf_bits(InfBits &b, int n) // n <= 16, LSB first (RFC 1951 3.1.1)
{
    while (b.cnt < n) {
        if (b.pos >= b.len) { b.err = true; return 0; }
        b.buf |= (uint32_t)b.p[b.pos++] << b.cnt;
        b.cnt += 8;
    }
    const uint32_t v = b.buf & ((1u << n) - 1u);
    b.buf >>= n;
    b.cnt -= n;
    return v;
}
// canonical Huffman code given as count[len] and the symbols ordered by (len, symbol) (RFC 1951 3.2.2)
struct InfCode { uint16_t count[16]; uint16_t symbol[288]; };
__device__ inline int inf_symbol(InfBits &b, const InfCode &h)
{
    int code = 0, first = 0, index = 0;
    for (int len = 1; len <= 15; len++) {
        code |= (int)inf_bits(b, 1);
        if (b.err) return -1;
        const int cnt = h.count[len];
        if (code - cnt < first) return h.symbol[index + (code - first)];
        index += cnt;
        first = (first + cnt) << 1;
        code <<= 1;
    }
    return -1; // no symbol has this code
}
// returns < 0 for an over-subscribed set of lengths, 0 for a complete code, > 0 for an incomplete one
__device__ inline int inf_construct(InfCode &h, const uint8_t *length, int n)
{
    uint16_t offs[16];
    for (int l = 0; l <= 15; l++) h.count[l] = 0;
    for (int s = 0; s < n; s++) h.count[length[s]]++;
    if (h.count[0] == n) return 0; // no codes at all: complete (and unusable)
    int left = 1;
    for (int l = 1; l <= 15; l++) {
        left <<= 1;
        left -= h.count[l];
        if (left < 0) return left;
    }
    offs[1] = 0;
    for (int l = 1; l < 15; l++) offs[l + 1] = offs[l] + h.count[l];
    for (int s = 0; s < n; s++)
        if (length[s]) h.symbol[offs[length[s]]++] = (uint16_t)s;
    return left;
}

// in[0 .. len) = a zlib stream.  Writes the first min(total, cap) decompressed bytes to out and returns the total
// number of decompressed bytes, or -1 where zlib.decompress raises (bad header, invalid block, truncated stream,
// Adler-32 mismatch).  Bytes behind the stream are ignored, as zlib.decompress ignores them.  One lane.
// A stream may decompress to more than cap bytes (the reference truncates, :88-89).  With ring == true, out has
// 32768 more bytes behind cap that serve as the LZ77 window of the bytes behind cap, and the stream is checked to
// its end like any other.  Without (the container path writes straight into the file's output, there is no room
// behind a package), bytes behind cap are not kept, so matches reaching behind cap and the checksum cannot be
// verified for such a stream (the reference's own packages never do that: original_length is the decompressed
// length).
__device__ inline long inflate_zlib(const uint8_t *in, long len, uint8_t *out, long cap, bool ring, InfCode *lencode, InfCode *distcode)
{
#define INF_AT(q) ((q) < cap ? (q) : cap + (((q) - cap) & 32767))
    if (len < 2) return -1;
    const uint32_t cmf = in[0], flg = in[1];
    if ((cmf & 15u) != 8u || (cmf >> 4) > 7u || ((cmf << 8) | flg) % 31u != 0u || (flg & 0x20u)) return -1;
    InfBits b;
    b.p = in; b.len = len; b.pos = 2; b.buf = 0; b.cnt = 0; b.err = false;
    long o = 0;
    uint32_t s1 = 1, s2 = 0; // Adler-32 of the output (RFC 1950)
    uint8_t lengths[320];
    for (;;) {
        const uint32_t last = inf_bits(b, 1), type = inf_bits(b, 2);
        if (b.err) return -1;
        if (type == 0) { // stored
            b.buf = 0; b.cnt = 0;
            if (b.pos + 4 > b.len) return -1;
            const uint32_t n = in[b.pos] | (in[b.pos + 1] << 8), nn = in[b.pos + 2] | (in[b.pos + 3] << 8);
            if (n != (~nn & 0xFFFFu)) return -1;
            b.pos += 4;
            if (b.pos + n > b.len) return -1;
            for (uint32_t k = 0; k < n; k++) {
                const uint8_t v = in[b.pos++];
                if (o < cap || ring) out[INF_AT(o)] = v;
                s1 += v; if (s1 >= ADLER_MOD) s1 -= ADLER_MOD;
                s2 += s1; if (s2 >= ADLER_MOD) s2 -= ADLER_MOD;
                o++;
            }
        } else if (type == 3) return -1;
        else {
            if (type == 1) { // fixed code (RFC 1951 3.2.6)
                for (int s = 0; s < 144; s++) lengths[s] = 8;
                for (int s = 144; s < 256; s++) lengths[s] = 9;
                for (int s = 256; s < 280; s++) lengths[s] = 7;
                for (int s = 280; s < 288; s++) lengths[s] = 8;
                inf_construct(*lencode, lengths, 288);
                for (int s = 0; s < 30; s++) lengths[s] = 5;
                inf_construct(*distcode, lengths, 30);
            } else { // dynamic code (3.2.7)
                const int nlen = (int)inf_bits(b, 5) + 257, ndist = (int)inf_bits(b, 5) + 1, ncode = (int)inf_bits(b, 4) + 4;
                if (b.err || nlen > 286 || ndist > 30) return -1;
                for (int k = 0; k < 19; k++) lengths[k] = 0;
                for (int k = 0; k < ncode; k++) lengths[c_clen_order[k]] = (uint8_t)inf_bits(b, 3);
                if (b.err) return -1;
                if (inf_construct(*lencode, lengths, 19) != 0) return -1; // the code-length code must be complete
                int idx = 0;
                while (idx < nlen + ndist) {
                    int sym = inf_symbol(b, *lencode);
                    if (sym < 0) return -1;
                    if (sym < 16) lengths[idx++] = (uint8_t)sym;
                    else {
                        int rep, val = 0;
                        if (sym == 16) {
                            if (idx == 0) return -1;
                            val = lengths[idx - 1];
                            rep = 3 + (int)inf_bits(b, 2);
                        } else if (sym == 17) rep = 3 + (int)inf_bits(b, 3);
                        else rep = 11 + (int)inf_bits(b, 7);
                        if (b.err || idx + rep > nlen + ndist) return -1;
                        while (rep--) lengths[idx++] = (uint8_t)val;
                    }
                }
                if (lengths[256] == 0) return -1; // no end-of-block code
                uint8_t dl[30];
                for (int k = 0; k < ndist; k++) dl[k] = lengths[nlen + k];
                int r = inf_construct(*lencode, lengths, nlen);
                if (r < 0 || (r > 0 && nlen - lencode->count[0] != 1)) return -1; // incomplete only with a single code
                r = inf_construct(*distcode, dl, ndist);
                if (r < 0 || (r > 0 && ndist - distcode->count[0] != 1)) return -1;
            }
            for (;;) {
                int sym = inf_symbol(b, *lencode);
                if (sym < 0) return -1;
                if (sym < 256) {
                    if (o < cap || ring) out[INF_AT(o)] = (uint8_t)sym;
                    s1 += (uint32_t)sym; if (s1 >= ADLER_MOD) s1 -= ADLER_MOD;
                    s2 += s1; if (s2 >= ADLER_MOD) s2 -= ADLER_MOD;
                    o++;
                } else if (sym == 256) break;
                else {
                    sym -= 257;
                    if (sym >= 29) return -1;
                    const int mlen = c_len_base[sym] + (int)inf_bits(b, c_len_extra[sym]);
                    const int ds = inf_symbol(b, *distcode);
                    if (ds < 0 || ds >= 30) return -1;
                    const long dist = c_dist_base[ds] + (long)inf_bits(b, c_dist_extra[ds]);
                    if (b.err || dist > o) return -1; // reaches before the start of the output
                    for (int k = 0; k < mlen; k++) {
                        const uint8_t v = (o - dist < cap || ring) ? out[INF_AT(o - dist)] : 0;
                        if (o < cap || ring) out[INF_AT(o)] = v;
                        s1 += v; if (s1 >= ADLER_MOD) s1 -= ADLER_MOD;
                        s2 += s1; if (s2 >= ADLER_MOD) s2 -= ADLER_MOD;
                        o++;
                    }
                }
            }
        }
        if (last) break;
    }
    // Adler-32, big-endian, on the next byte boundary
    if (b.pos + 4 > b.len) return -1;
    const uint32_t want = ((uint32_t)in[b.pos] << 24) | ((uint32_t)in[b.pos + 1] << 16) | ((uint32_t)in[b.pos + 2] << 8) | in[b.pos + 3];
    if ((o <= cap || ring) && want != ((s2 << 16) | s1)) return -1;
    return o;
#undef INF_AT
}

// ---- deflate ---------------------------------------------------------------------------------------------------
struct DefBits { uint8_t *p; int o, cap; unsigned long long acc; int nb; };
__device__ __forceinline__ void def_put(DefBits &w, uint32_t v, int n) // LSB first
{
    w.acc |= (unsigned long long)v << w.nb;
    w.nb += n;
    while (w.nb >= 8) {
        if (w.o < w.cap) w.p[w.o] = (uint8_t)w.acc;
        w.o++;
        w.acc >>= 8;
        w.nb -= 8;
    }
}
__device__ __forceinline__ void def_code(DefBits &w, uint32_t code, int n) { def_put(w, __brev(code) >> (32 - n), n); } // Huffman codes go MSB first
__device__ __forceinline__ void def_litlen(DefBits &w, int sym) // fixed code, RFC 1951 3.2.6
{
    if (sym < 144) def_code(w, 0x30u + (uint32_t)sym, 8);
    else if (sym < 256) def_code(w, 0x190u + (uint32_t)(sym - 144), 9);
    else if (sym < 280) def_code(w, (uint32_t)(sym - 256), 7);
    else def_code(w, 0xC0u + (uint32_t)(sym - 280), 8);
}

#define DEF_HB 12
// in: n bytes in shared memory; head: (1 << DEF_HB) u16 slots in shared memory, all 0xFFFF.  Writes the zlib stream
// to out (global, cap bytes) and returns its length (> cap: it did not fit).  adler = Adler-32 of the input.  One lane.
__device__ inline int deflate_fixed(const uint8_t *in, int n, uint8_t *out, int cap, uint16_t *head, uint32_t adler)
{
    DefBits w;
    w.p = out; w.o = 0; w.cap = cap; w.acc = 0; w.nb = 0;
    def_put(w, 0x78, 8);  // CMF: deflate, 32 KiB window
    def_put(w, 0x9C, 8);  // FLG: no dictionary, check bits
    def_put(w, 1, 1);     // BFINAL
    def_put(w, 1, 2);     // BTYPE = 01, fixed Huffman
    int p = 0;
    while (p < n) {
        int best = 0, bdist = 0;
        if (p + 3 <= n) {
            const uint32_t h = (((uint32_t)in[p] | ((uint32_t)in[p + 1] << 8) | ((uint32_t)in[p + 2] << 16)) * 2654435761u) >> (32 - DEF_HB);
            const int cand = head[h];
            head[h] = (uint16_t)p;
            if (cand != 0xFFFF) {
                const int maxl = min(258, n - p);
                int l = 0;
                while (l < maxl && in[cand + l] == in[p + l]) l++;
                if (l >= 3) { best = l; bdist = p - cand; }
            }
        }
        if (best) {
            int ls = 28;
            while (c_len_base[ls] > best) ls--;
            def_litlen(w, 257 + ls);
            def_put(w, (uint32_t)(best - c_len_base[ls]), c_len_extra[ls]);
            int ds = 29;
            while (c_dist_base[ds] > bdist) ds--;
            def_code(w, (uint32_t)ds, 5);
            def_put(w, (uint32_t)(bdist - c_dist_base[ds]), c_dist_extra[ds]);
            for (int q = p + 1; q < p + best && q + 3 <= n; q++) // the skipped positions stay findable
                head[(((uint32_t)in[q] | ((uint32_t)in[q + 1] << 8) | ((uint32_t)in[q + 2] << 16)) * 2654435761u) >> (32 - DEF_HB)] = (uint16_t)q;
            p += best;
        } else {
            def_litlen(w, in[p]);
            p++;
        }
    }
    def_litlen(w, 256); // end of block
    if (w.nb) def_put(w, 0, 8 - w.nb);
    def_put(w, adler >> 24, 8); def_put(w, (adler >> 16) & 0xFF, 8); def_put(w, (adler >> 8) & 0xFF, 8); def_put(w, adler & 0xFF, 8);
    return w.o;
}
