/*
 * oracle/ambc_oracle.c -- CPU restatement of the reference's chunked
 * encode / select / decode path.  TEST INFRASTRUCTURE ONLY.
 *
 * Nothing in the product (adaptive_compression_b200/, main.py) may import, link
 * or execute this file.  It is used by tests/, __graft_entry__.smoke() and by
 * bench.py's cpu_baseline / --impl reference legs as the checker / CPU baseline.
 *
 * Parity status: PINNED.  The reference ships no golden vectors,
 * so the pin is the unmodified Python reference itself, run in the build
 * container by oracle/make_golden.py (fixtures in tests/golden/) and compared
 * to this file by tests/test_oracle_golden.py.
 *
 * Every function cites the reference file:line (relative to /root/reference)
 * whose behaviour it restates.  Plain C99, no dependencies beyond libm.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define ORC_API __attribute__((visibility("default")))

/* error codes of the codec functions (mirror the Python exception raised) */
#define ORC_ERR_INDEX (-1) /* IndexError  */
#define ORC_ERR_VALUE (-2) /* ValueError  */

/* ------------------------------------------------------------------ */
/* RLE  (compression_methods.py:70-180)                                 */
/* ------------------------------------------------------------------ */

/* compression_methods.py:78-114.  out capacity: 2*n. */
ORC_API long orc_rle_compress(const uint8_t *d, long n, uint8_t *out)
{
    if (n <= 0) return 0;
    long o = 0;
    uint8_t cur = d[0];
    unsigned cnt = 1;
    for (long i = 1; i < n; i++) {
        if (d[i] == cur && cnt < 255) {
            cnt++;
        } else {
            out[o++] = cur;
            out[o++] = (uint8_t)cnt;
            cur = d[i];
            cnt = 1;
        }
    }
    out[o++] = cur;
    out[o++] = (uint8_t)cnt;
    return o;
}

/* compression_methods.py:116-152.  out capacity: max(orig_len, 255*len/2).
 * Returns the number of bytes produced (0 for empty input, else orig_len). */
ORC_API long orc_rle_decompress(const uint8_t *d, long len, long orig_len, uint8_t *out)
{
    if (len <= 0) return 0; /* :127-128 returns b'' */
    long o = 0;
    for (long i = 0; i + 1 < len; i += 2) { /* :132-136, odd trailing byte ignored */
        uint8_t b = d[i];
        long c = d[i + 1];
        for (long k = 0; k < c && o < orig_len; k++) out[o++] = b;
        /* bytes past orig_len are cut by :142-144; no need to materialise them */
    }
    while (o < orig_len) out[o++] = 0; /* :145-150 zero pad */
    return orig_len;
}

/* compression_methods.py:154-180 */
ORC_API int orc_rle_should_use(const uint8_t *d, long n)
{
    if (n < 4) return 0;
    long s = n < 1000 ? n : 1000;
    long step = n / s; if (step < 1) step = 1;
    long rep = 0;
    for (long i = 0; i < n - 1; i += step)
        if (d[i] == d[i + 1]) rep++;
    double ratio = (double)rep / (double)(s - 1);
    return ratio > 0.3;
}

/* ------------------------------------------------------------------ */
/* Dictionary / greedy LZ77  (compression_methods.py:183-343)          */
/* ------------------------------------------------------------------ */

#define LZ_WINDOW 4096
#define LZ_LOOKAHEAD 32

/* compression_methods.py:283-313: scan every window position in ascending
 * order, keep the first strictly-longer match ("earliest longest"). */
static void lz_longest_naive(const uint8_t *d, long n, long pos, long *mpos, long *mlen)
{
    long start = pos - LZ_WINDOW; if (start < 0) start = 0;
    long la = n - pos; if (la > LZ_LOOKAHEAD) la = LZ_LOOKAHEAD;
    long bp = 0, bl = 0;
    for (long i = start; i < pos; i++) {
        long l = 0;
        while (l < la && d[i + l] == d[pos + l]) l++;
        if (l > bl) { bp = i; bl = l; }
    }
    *mpos = bp; *mlen = bl;
}

/* Same result as lz_longest_naive for the only thing the caller uses (the
 * match when its length is > 2): candidates restricted to earlier positions
 * that share the first three bytes, visited oldest-first through a per-hash
 * list.  Exists so that tests can run the oracle on tens of MiB; proven equal
 * to the naive scan by tests/test_oracle_golden.py. */
typedef struct { int *first, *next, *last; } lz_index;

static unsigned lz_hash3(const uint8_t *p)
{
    unsigned t = p[0] | (p[1] << 8) | (p[2] << 16);
    return (t * 2654435761u) >> 18; /* 14 bits */
}

/* compression_methods.py:195-234.  out capacity: 2*n.  fast!=0 selects the
 * indexed search. */
ORC_API long orc_lz_compress_ex(const uint8_t *d, long n, uint8_t *out, int fast)
{
    if (n <= 0) return 0;
    long o = 0, pos = 0;
    int *first = NULL, *next = NULL, *last = NULL;
    long indexed = 0; /* positions [0, indexed) are in the lists */
    if (fast) {
        first = (int *)malloc(sizeof(int) * 16384);
        last = (int *)malloc(sizeof(int) * 16384);
        next = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
        for (int i = 0; i < 16384; i++) first[i] = last[i] = -1;
    }
    while (pos < n) {
        long mp = 0, ml = 0;
        if (!fast) {
            lz_longest_naive(d, n, pos, &mp, &ml);
        } else {
            /* bring the index up to date: all trigram starts < pos */
            for (; indexed < pos && indexed + 2 < n; indexed++) {
                unsigned h = lz_hash3(d + indexed);
                next[indexed] = -1;
                if (last[h] < 0) first[h] = (int)indexed; else next[last[h]] = (int)indexed;
                last[h] = (int)indexed;
            }
            long la = n - pos; if (la > LZ_LOOKAHEAD) la = LZ_LOOKAHEAD;
            if (la >= 3) {
                long start = pos - LZ_WINDOW; if (start < 0) start = 0;
                unsigned h = lz_hash3(d + pos);
                /* drop list heads that left the window for good */
                while (first[h] >= 0 && first[h] < start) first[h] = next[first[h]];
                if (first[h] < 0) last[h] = -1;
                for (int i = first[h]; i >= 0 && i < pos; i = next[i]) {
                    long l = 0;
                    while (l < la && d[i + l] == d[pos + l]) l++;
                    if (l > ml) { mp = i; ml = l; if (l == la) break; }
                }
            }
        }
        if (ml > 2) { /* :215-227 */
            long dist = pos - mp;
            out[o++] = 1;
            out[o++] = (uint8_t)(dist & 0xFF);
            out[o++] = (uint8_t)((dist >> 8) & 0xFF);
            out[o++] = (uint8_t)ml;
            pos += ml;
        } else { /* :228-232 */
            out[o++] = 0;
            out[o++] = d[pos];
            pos += 1;
        }
    }
    if (fast) { free(first); free(last); free(next); }
    return o;
}

ORC_API long orc_lz_compress(const uint8_t *d, long n, uint8_t *out)
{
    return orc_lz_compress_ex(d, n, out, 0);
}

/* compression_methods.py:236-281.  out capacity: orig_len + 256.
 * Returns bytes produced (<= orig_len, may be SHORTER: no padding, :281) or
 * ORC_ERR_INDEX where Python raises IndexError (:275 / :278). */
ORC_API long orc_lz_decompress(const uint8_t *d, long len, long orig_len, uint8_t *out)
{
    if (len <= 0) return 0;
    long pos = 0, o = 0;
    while (pos < len && o < orig_len) {
        uint8_t flag = d[pos++];
        if (flag == 0) {
            if (pos < len) out[o++] = d[pos++];
        } else {
            if (pos + 2 < len) {
                long dist = d[pos] | (d[pos + 1] << 8);
                pos += 2;
                long length = d[pos++];
                long start = o - dist;
                for (long i = 0; i < length; i++) {
                    if (start + i < o) {
                        long idx = start + i;
                        if (idx < 0) idx += o; /* Python negative index */
                        if (idx < 0) return ORC_ERR_INDEX;
                        out[o] = out[idx];
                        o++;
                    } else {
                        if (o == 0) return ORC_ERR_INDEX; /* decompressed[-1] on empty */
                        out[o] = out[o - 1];
                        o++;
                    }
                }
            }
        }
    }
    return o < orig_len ? o : orig_len;
}

/* compression_methods.py:315-343 */
ORC_API int orc_lz_should_use(const uint8_t *d, long n)
{
    if (n < 100) return 0;
    long s = n < 1000 ? n : 1000;
    long cnt = n - 3 < s ? n - 3 : s;
    /* distinct 3-byte slices among the first cnt positions */
    uint32_t *keys = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(cnt > 0 ? cnt : 1));
    long distinct = 0;
    /* tiny open-addressing set, 4096 slots */
    uint32_t *tab = (uint32_t *)calloc(4096, sizeof(uint32_t));
    for (long i = 0; i < cnt; i++) {
        uint32_t t = (d[i] | (d[i + 1] << 8) | (d[i + 2] << 16)) + 1u; /* +1: 0 = empty */
        uint32_t h = (t * 2654435761u) >> 20;
        for (;;) {
            if (tab[h] == 0) { tab[h] = t; distinct++; break; }
            if (tab[h] == t) break;
            h = (h + 1) & 4095;
        }
    }
    free(tab); free(keys);
    double ratio = (double)distinct / (double)s;
    return ratio < 0.8;
}

/* ------------------------------------------------------------------ */
/* Huffman  (compression_methods.py:346-574)                           */
/* ------------------------------------------------------------------ */

typedef struct {
    int k;              /* number of distinct symbols, table order */
    int sym[256];       /* symbol of table entry j */
    long long w[256];   /* weight of table entry j */
    /* tree: nodes 0..k-1 leaves (table order), k..2k-2 internal */
    int child0[512], child1[512];
    int root;
    uint32_t code[256]; /* indexed by symbol */
    int clen[256];      /* indexed by symbol; 0 = absent */
} huff_t;

/* compression_methods.py:472-500 (+502-549).  The heap holds lists
 * [weight, [sym, code], ...] compared lexicographically, i.e. by
 * (weight, first member's symbol); a merged node's first member is lo's.
 * lo (smaller key) is prefixed '0', hi '1'  (:486-494).
 * Returns 0, or ORC_ERR_INDEX for k<=1 (heappop on empty heap / code[-1] of
 * an empty code, :497 / :527). */
static int huff_build(huff_t *h)
{
    int k = h->k;
    if (k <= 1) return ORC_ERR_INDEX;
    long long weight[512];
    int leader[512], alive[512];
    int nn = k;
    for (int j = 0; j < k; j++) { weight[j] = h->w[j]; leader[j] = h->sym[j]; alive[j] = 1; }
    int live = k;
    while (live > 1) {
        int a = -1, b = -1; /* a = smallest, b = second smallest */
        for (int j = 0; j < nn; j++) {
            if (!alive[j]) continue;
            if (a < 0 || weight[j] < weight[a] || (weight[j] == weight[a] && leader[j] < leader[a])) {
                b = a; a = j;
            } else if (b < 0 || weight[j] < weight[b] || (weight[j] == weight[b] && leader[j] < leader[b])) {
                b = j;
            }
        }
        alive[a] = alive[b] = 0;
        weight[nn] = weight[a] + weight[b];
        leader[nn] = leader[a];
        h->child0[nn] = a;
        h->child1[nn] = b;
        alive[nn] = 1;
        nn++;
        live--;
    }
    h->root = nn - 1;
    /* codes: walk down from the root */
    for (int s = 0; s < 256; s++) { h->clen[s] = 0; h->code[s] = 0; }
    uint32_t ncode[512]; int nlen[512];
    ncode[h->root] = 0; nlen[h->root] = 0;
    for (int j = nn - 1; j >= k; j--) {
        int c0 = h->child0[j], c1 = h->child1[j];
        ncode[c0] = (ncode[j] << 1);     nlen[c0] = nlen[j] + 1;
        ncode[c1] = (ncode[j] << 1) | 1; nlen[c1] = nlen[j] + 1;
    }
    for (int j = 0; j < k; j++) { h->code[h->sym[j]] = ncode[j]; h->clen[h->sym[j]] = nlen[j]; }
    return 0;
}

/* compression_methods.py:354-405.  out capacity: 1 + 5*256 + 4 + 4*n.
 * Returns payload length, ORC_ERR_VALUE for 256 distinct symbols
 * (bytearray.append(256), :382) or ORC_ERR_INDEX for 1 distinct symbol. */
ORC_API long orc_huff_compress(const uint8_t *d, long n, uint8_t *out)
{
    if (n <= 0) return 0;
    huff_t *h = (huff_t *)malloc(sizeof(huff_t));
    int slot[256];
    for (int s = 0; s < 256; s++) slot[s] = -1;
    h->k = 0;
    for (long i = 0; i < n; i++) { /* Counter in first-occurrence order, :368-370 */
        int s = d[i];
        if (slot[s] < 0) { slot[s] = h->k; h->sym[h->k] = s; h->w[h->k] = 0; h->k++; }
        h->w[slot[s]]++;
    }
    int rc = huff_build(h); /* :373, raises before :382 */
    if (rc < 0) { free(h); return rc; }
    if (h->k == 256) { free(h); return ORC_ERR_VALUE; }
    long o = 0;
    out[o++] = (uint8_t)h->k;
    for (int j = 0; j < h->k; j++) {
        out[o++] = (uint8_t)h->sym[j];
        uint32_t c = (uint32_t)h->w[j];
        out[o++] = c & 0xFF; out[o++] = (c >> 8) & 0xFF; out[o++] = (c >> 16) & 0xFF; out[o++] = (c >> 24) & 0xFF;
    }
    unsigned long long nbits = 0;
    for (long i = 0; i < n; i++) nbits += (unsigned)h->clen[d[i]];
    uint32_t nb = (uint32_t)nbits;
    out[o++] = nb & 0xFF; out[o++] = (nb >> 8) & 0xFF; out[o++] = (nb >> 16) & 0xFF; out[o++] = (nb >> 24) & 0xFF;
    long nbytes = (long)((nbits + 7) / 8);
    memset(out + o, 0, (size_t)nbytes);
    unsigned long long bp = 0;
    for (long i = 0; i < n; i++) { /* MSB-first, :389-403 */
        uint32_t c = h->code[d[i]]; int l = h->clen[d[i]];
        for (int b = l - 1; b >= 0; b--, bp++)
            if ((c >> b) & 1) out[o + (long)(bp >> 3)] |= (uint8_t)(0x80 >> (bp & 7));
    }
    free(h);
    return o + nbytes;
}

/* compression_methods.py:407-470.  out capacity: orig_len (+1).
 * Returns bytes produced (may be SHORTER than orig_len; no padding) or
 * ORC_ERR_INDEX where Python raises IndexError. */
ORC_API long orc_huff_decompress(const uint8_t *d, long len, long orig_len, uint8_t *out)
{
    if (len <= 0) return 0; /* :418-419 */
    long pos = 0;
    int ne = d[pos++];
    huff_t *h = (huff_t *)malloc(sizeof(huff_t));
    int slot[256];
    for (int s = 0; s < 256; s++) slot[s] = -1;
    h->k = 0;
    for (int e = 0; e < ne; e++) { /* :429-436, dict keeps first position, last value */
        if (pos >= len) { free(h); return ORC_ERR_INDEX; } /* data[pos] */
        int s = d[pos++];
        uint32_t c = 0;
        for (int b = 0; b < 4; b++) if (pos + b < len) c |= (uint32_t)d[pos + b] << (8 * b);
        /* int.from_bytes of a short slice: only the bytes present */
        pos += 4;
        if (slot[s] < 0) { slot[s] = h->k; h->sym[h->k] = s; h->k++; }
        h->w[slot[s]] = c;
    }
    int rc = huff_build(h);
    if (rc < 0) { free(h); return rc; }
    uint32_t nb = 0;
    for (int b = 0; b < 4; b++) if (pos + b < len) nb |= (uint32_t)d[pos + b] << (8 * b);
    pos += 4;
    long avail = pos < len ? (len - pos) * 8 : 0;
    long nbits = (long)nb < avail ? (long)nb : avail;
    long o = 0;
    int node = h->root;
    for (long bp = 0; bp < nbits; bp++) { /* :457-468 */
        int bit = (d[pos + (bp >> 3)] >> (7 - (bp & 7))) & 1;
        node = bit ? h->child1[node] : h->child0[node];
        if (node < h->k) {
            out[o++] = (uint8_t)h->sym[node];
            node = h->root;
            if (o >= orig_len) break;
        }
    }
    free(h);
    return o;
}

/* compression_methods.py:551-574: Python-float running sum in the Counter's
 * first-occurrence order. */
ORC_API double orc_entropy(const uint8_t *d, long n)
{
    long cnt[256]; int order[256]; int k = 0;
    for (int s = 0; s < 256; s++) cnt[s] = 0;
    for (long i = 0; i < n; i++) { if (cnt[d[i]]++ == 0) order[k++] = d[i]; }
    double e = 0.0;
    for (int j = 0; j < k; j++) {
        double p = (double)cnt[order[j]] / (double)n;
        e -= p * log2(p);
    }
    return e;
}

ORC_API int orc_huff_should_use(const uint8_t *d, long n)
{
    if (n < 100) return 0;
    return orc_entropy(d, n) < 7.0;
}

/* ------------------------------------------------------------------ */
/* Delta  (compression_methods.py:577-667)                             */
/* ------------------------------------------------------------------ */

/* :585-608 */
ORC_API long orc_delta_compress(const uint8_t *d, long n, uint8_t *out)
{
    if (n <= 0) return 0;
    out[0] = d[0];
    for (long i = 1; i < n; i++) out[i] = (uint8_t)(d[i] - d[i - 1]);
    return n;
}

/* :610-638: returns min(len, orig_len) bytes */
ORC_API long orc_delta_decompress(const uint8_t *d, long len, long orig_len, uint8_t *out)
{
    if (len <= 0) return 0;
    long m = len < orig_len ? len : orig_len;
    uint8_t prev = 0;
    for (long i = 0; i < m; i++) { prev = (uint8_t)(i ? prev + d[i] : d[0]); out[i] = prev; }
    return m;
}

/* :640-667 */
ORC_API int orc_delta_should_use(const uint8_t *d, long n)
{
    if (n < 4) return 0;
    long s = n < 1000 ? n : 1000;
    long step = n / s; if (step < 1) step = 1;
    long small = 0;
    for (long i = 0; i < n - 1; i += step) {
        int dl = (int)d[i] - (int)d[i + 1]; if (dl < 0) dl = -dl;
        if (dl < 32) small++;
    }
    double ratio = (double)small / (double)(s - 1);
    return ratio > 0.5;
}

/* ------------------------------------------------------------------ */
/* NoCompression (compression_methods.py:670-713)                      */
/* ------------------------------------------------------------------ */
ORC_API long orc_raw_decompress(const uint8_t *d, long len, long orig_len, uint8_t *out)
{
    long m = len < orig_len ? len : orig_len;
    if (m > 0) memcpy(out, d, (size_t)m);
    if (m < orig_len) memset(out + m, 0, (size_t)(orig_len - m));
    return orig_len;
}

/* ------------------------------------------------------------------ */
/* generic method dispatch                                             */
/* ------------------------------------------------------------------ */

/* adaptive_compressor.py:114-127 */
static void method_range(int id, long *lo, long *hi)
{
    switch (id) {
    case 1: *lo = 32; *hi = 4096; break;
    case 2: *lo = 128; *hi = 8192; break;
    case 3: *lo = 32; *hi = 8192; break;
    case 4: *lo = 32; *hi = 4096; break;
    default: *lo = 1; *hi = 999999999; break;
    }
}

ORC_API int orc_should_use(int id, const uint8_t *d, long n)
{
    switch (id) {
    case 1: return orc_rle_should_use(d, n);
    case 2: return orc_lz_should_use(d, n);
    case 3: return orc_huff_should_use(d, n);
    case 4: return orc_delta_should_use(d, n);
    default: return 1;
    }
}

static int g_lz_fast = 0;
/* test knob: 1 = indexed LZ search (same bytes, proven in tests), 0 = the
 * reference's own O(n*window) scan */
ORC_API void orc_set_lz_fast(int on) { g_lz_fast = on; }

ORC_API long orc_compress(int id, const uint8_t *d, long n, uint8_t *out)
{
    switch (id) {
    case 1: return orc_rle_compress(d, n, out);
    case 2: return orc_lz_compress_ex(d, n, out, g_lz_fast);
    case 3: return orc_huff_compress(d, n, out);
    case 4: return orc_delta_compress(d, n, out);
    default: if (n > 0) memcpy(out, d, (size_t)n); return n;
    }
}

ORC_API long orc_decompress(int id, const uint8_t *d, long len, long orig_len, uint8_t *out)
{
    switch (id) {
    case 1: return orc_rle_decompress(d, len, orig_len, out);
    case 2: return orc_lz_decompress(d, len, orig_len, out);
    case 3: return orc_huff_decompress(d, len, orig_len, out);
    case 4: return orc_delta_decompress(d, len, orig_len, out);
    case 255: return orc_raw_decompress(d, len, orig_len, out);
    default: return -100; /* not a known method */
    }
}

/* ------------------------------------------------------------------ */
/* Container body  (adaptive_compressor.py:363-454, 537-700)            */
/* ------------------------------------------------------------------ */

typedef struct {
    const int *methods;  int n_methods;      /* trial order, ids != 255 */
    const long *cands;   int n_cands;        /* descending, as CHUNK_SIZE_CANDIDATES */
    const uint8_t *marker; int marker_bytes; /* marker_bytes_aligned */
    int per_chunk_raw;   /* 0 = reference rule (rest of file raw); 1 = labelled
                            extension: a losing chunk is its own raw package */
} orc_cfg;

static void put_u32(uint8_t *p, uint32_t v) { p[0] = v; p[1] = v >> 8; p[2] = v >> 16; p[3] = v >> 24; }
static uint32_t get_u32(const uint8_t *p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }

/* adaptive_compressor.py:537-590 */
static void pick_best(const orc_cfg *c, const uint8_t *data, long total, long position,
                      uint8_t *scratch, long *out_csize, int *out_method)
{
    long remain = total - position;
    long best_csize = remain; int best_method = 255; double best_ratio = 1.0;
    long overhead = c->marker_bytes + 14; /* :623-629 */
    for (int ci = 0; ci < c->n_cands; ci++) {
        long cand = c->cands[ci];
        if (cand > remain) cand = remain;
        if (cand <= 0) break;
        const uint8_t *chunk = data + position;
        double local_ratio = 1.0; int local_method = 255;
        for (int mi = 0; mi < c->n_methods; mi++) {
            int id = c->methods[mi];
            if (id == 255) continue;
            long lo, hi; method_range(id, &lo, &hi);
            if (!(lo <= cand && cand <= hi)) continue;
            if (!orc_should_use(id, chunk, cand)) continue;
            long cl = orc_compress(id, chunk, cand, scratch);
            if (cl < 0) continue; /* exception swallowed, :578-579 */
            double ratio = (double)(cl + overhead) / (double)cand;
            if (ratio < local_ratio) { local_ratio = ratio; local_method = id; }
        }
        if (local_ratio < best_ratio) { best_ratio = local_ratio; best_csize = cand; best_method = local_method; }
    }
    if (best_method == 255 && best_csize == remain) {
        *out_csize = remain; *out_method = 255;
        if (c->per_chunk_raw && c->n_cands > 0) {
            /* extension: smallest candidate as its own raw package */
            long cand = c->cands[c->n_cands - 1];
            if (cand > remain) cand = remain;
            if (cand > 0) *out_csize = cand;
        }
    } else {
        *out_csize = best_csize; *out_method = best_method;
    }
}

/* adaptive_compressor.py:609-621 */
static long put_package(const orc_cfg *c, uint8_t *out, int type, uint32_t used, uint32_t orig,
                        const uint8_t *payload, uint32_t comp)
{
    long o = 0;
    memcpy(out, c->marker, (size_t)c->marker_bytes); o += c->marker_bytes;
    out[o++] = (uint8_t)type;
    out[o++] = 0;
    put_u32(out + o, used); o += 4;
    put_u32(out + o, orig); o += 4;
    put_u32(out + o, comp); o += 4;
    if (comp) memcpy(out + o, payload, comp);
    return o + comp;
}

/*
 * adaptive_compressor.py:363-394 (+631-700).  out capacity:
 * total + (total/min_cand + 2) * (marker_bytes+14) + 16.
 * map_* (optional, capacity = number of packages): per-package type, orig, comp.
 * Returns body length (incl. END package); *n_pkgs = data packages written.
 */
ORC_API long orc_compress_body(const uint8_t *data, long total,
                               const long *cands, int n_cands,
                               const int *methods, int n_methods,
                               const uint8_t *marker, int marker_bytes,
                               int per_chunk_raw,
                               uint8_t *out,
                               int *map_type, long *map_orig, long *map_comp, long map_cap,
                               long *n_pkgs)
{
    orc_cfg c = { methods, n_methods, cands, n_cands, marker, marker_bytes, per_chunk_raw };
    long maxc = 0;
    for (int i = 0; i < n_cands; i++) if (cands[i] > maxc) maxc = cands[i];
    if (maxc > total) maxc = total;
    uint8_t *scratch = (uint8_t *)malloc((size_t)(4 * maxc + 2048));
    uint8_t *payload = (uint8_t *)malloc((size_t)(4 * maxc + 2048));
    long o = 0, position = 0, np = 0;
    long overhead = marker_bytes + 14;
    while (position < total) {
        long csize; int mid;
        pick_best(&c, data, total, position, scratch, &csize, &mid);
        const uint8_t *chunk = data + position;
        int type = 255; long comp = csize; const uint8_t *pl = chunk;
        if (mid != 255) { /* :658-688: compress again, keep only if beneficial */
            long cl = orc_compress(mid, chunk, csize, payload);
            if (cl >= 0 && cl + overhead < csize) { type = mid; comp = cl; pl = payload; }
        }
        o += put_package(&c, out + o, type, (uint32_t)csize, (uint32_t)csize, pl, (uint32_t)comp);
        if (np < map_cap) {
            if (map_type) map_type[np] = type;
            if (map_orig) map_orig[np] = csize;
            if (map_comp) map_comp[np] = comp;
        }
        np++;
        position += csize;
    }
    /* END package, adaptive_compressor.py:595-607: marker, 0, 0, u16 0, u32 0, u32 0 */
    memcpy(out + o, marker, (size_t)marker_bytes); o += marker_bytes;
    memset(out + o, 0, 12); o += 12;
    free(scratch); free(payload);
    if (n_pkgs) *n_pkgs = np;
    return o;
}

/*
 * adaptive_compressor.py:396-454.  known[] lists the type ids present in
 * method_lookup (anything else is copied through as raw, :432-435).
 * out capacity: orig_size (+ slack handled internally).
 * Returns orig_size, or -3 for "Marker mismatch in chunk header." (:406-407).
 */
ORC_API long orc_decompress_body(const uint8_t *body, long blen, long orig_size,
                                 const uint8_t *marker, int marker_bytes,
                                 const int *known, int n_known,
                                 uint8_t *out)
{
    long pos = 0, o = 0;
    long cap = orig_size + 4096;
    uint8_t *acc = (uint8_t *)malloc((size_t)cap);
    while (pos < blen) {
        long needed = marker_bytes + 14;
        if (pos + needed > blen) break;
        if (memcmp(body + pos, marker, (size_t)marker_bytes) != 0) { free(acc); return -3; }
        pos += marker_bytes;
        int type = body[pos++];
        pos++; /* k_value */
        pos += 4; /* used */
        long orig_len = get_u32(body + pos); pos += 4;
        long comp_len = get_u32(body + pos); pos += 4;
        if (type == 0) break;
        if (pos + comp_len > blen) break;
        const uint8_t *payload = body + pos;
        pos += comp_len;
        int is_known = 0;
        for (int i = 0; i < n_known; i++) if (known[i] == type) is_known = 1;
        long need = is_known ? orig_len + 512 : comp_len;
        if (o + need + 512 > cap) {
            cap = (o + need + 512) * 2;
            acc = (uint8_t *)realloc(acc, (size_t)cap);
        }
        if (!is_known) {
            memcpy(acc + o, payload, (size_t)comp_len); o += comp_len;
        } else {
            long r = orc_decompress(type, payload, comp_len, orig_len, acc + o);
            if (r < 0) { memset(acc + o, 0, (size_t)orig_len); r = orig_len; } /* :440-442 */
            o += r;
        }
        if (o >= orig_size) break;
    }
    long m = o < orig_size ? o : orig_size;
    memcpy(out, acc, (size_t)m);
    if (m < orig_size) memset(out + m, 0, (size_t)(orig_size - m));
    free(acc);
    return orig_size;
}

/* ------------------------------------------------------------------ */
/* Marker search  (marker_finder.py:22-123)                            */
/* ------------------------------------------------------------------ */

/*
 * Smallest L in 1..max_len such that some L-bit value is absent from all
 * nbits-L+1 windows of the MSB-first bit stream; smallest such value.
 * marker_out (4 bytes): value left-aligned, zero padded (:100-110).
 * The optional sampling branch (:38-51) is applied by the caller.
 * Returns L, or 0 when none exists up to max_len (ValueError, :123).
 */
ORC_API int orc_find_marker(const uint8_t *d, long n, int max_len, uint8_t *marker_out)
{
    long nbits = n * 8;
    for (int L = 1; L <= max_len && L <= 32; L++) {
        uint64_t space = 1ull << L;
        uint8_t *found = (uint8_t *)calloc((size_t)((space + 7) / 8), 1);
        if (nbits >= L) {
            uint64_t mask = (L == 64) ? ~0ull : (space - 1);
            uint64_t w = 0;
            for (long b = 0; b < nbits; b++) {
                w = ((w << 1) | ((d[b >> 3] >> (7 - (b & 7))) & 1)) & mask;
                if (b >= L - 1) found[w >> 3] |= (uint8_t)(1u << (w & 7));
            }
        }
        for (uint64_t v = 0; v < space; v++) {
            if (!(found[v >> 3] & (1u << (v & 7)))) {
                int nb = (L + 7) / 8;
                uint64_t aligned = v << (nb * 8 - L);
                for (int k = 0; k < nb; k++) marker_out[k] = (uint8_t)(aligned >> (8 * (nb - 1 - k)));
                free(found);
                return L;
            }
        }
        free(found);
    }
    return 0;
}

/* marker_finder.py:38-48: every step-th byte, concatenated, cut to sample_size */
ORC_API long orc_marker_sample(const uint8_t *d, long n, long sample_size, uint8_t *out)
{
    if (!(sample_size > 0 && n > sample_size)) { memcpy(out, d, (size_t)n); return n; }
    long step = n / sample_size, o = 0;
    for (long i = 0; i < n && o < sample_size; i += step) out[o++] = d[i];
    return o;
}

/* ------------------------------------------------------------------ */
/* MD5 (RFC 1321) -- the header checksum, adaptive_compressor.py:234     */
/* ------------------------------------------------------------------ */
static const uint32_t md5_k[64] = {
    0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501,
    0x698098d8, 0x8b44f7af, 0xffff5bb1, 0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821,
    0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453, 0xd8a1e681, 0xe7d3fbc8,
    0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a,
    0xfffa3942, 0x8771f681, 0x6d9d6122, 0xfde5380c, 0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70,
    0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05, 0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665,
    0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d, 0x85845dd1,
    0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1, 0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391 };
static const int md5_s[64] = { 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22,
                               5, 9, 14, 20, 5, 9, 14, 20, 5, 9, 14, 20, 5, 9, 14, 20,
                               4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23,
                               6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21 };
static void md5_block(uint32_t st[4], const uint8_t *p)
{
    uint32_t m[16];
    for (int i = 0; i < 16; i++) m[i] = get_u32(p + 4 * i);
    uint32_t a = st[0], b = st[1], c = st[2], d = st[3];
    for (int i = 0; i < 64; i++) {
        uint32_t f; int g;
        if (i < 16) { f = (b & c) | (~b & d); g = i; }
        else if (i < 32) { f = (d & b) | (~d & c); g = (5 * i + 1) & 15; }
        else if (i < 48) { f = b ^ c ^ d; g = (3 * i + 5) & 15; }
        else { f = c ^ (b | ~d); g = (7 * i) & 15; }
        uint32_t t = d; d = c; c = b;
        uint32_t x = a + f + md5_k[i] + m[g];
        b = b + ((x << md5_s[i]) | (x >> (32 - md5_s[i])));
        a = t;
    }
    st[0] += a; st[1] += b; st[2] += c; st[3] += d;
}
ORC_API void orc_md5(const uint8_t *d, long n, uint8_t *digest)
{
    uint32_t st[4] = { 0x67452301, 0xefcdab89, 0x98badcfe, 0x10325476 };
    long i = 0;
    for (; i + 64 <= n; i += 64) md5_block(st, d + i);
    uint8_t tail[128]; long r = n - i;
    memset(tail, 0, sizeof tail);
    if (r) memcpy(tail, d + i, (size_t)r);
    tail[r] = 0x80;
    long tl = (r < 56) ? 64 : 128;
    uint64_t bits = (uint64_t)n * 8;
    for (int k = 0; k < 8; k++) tail[tl - 8 + k] = (uint8_t)(bits >> (8 * k));
    md5_block(st, tail);
    if (tl == 128) md5_block(st, tail + 64);
    for (int k = 0; k < 4; k++) put_u32(digest + 4 * k, st[k]);
}

/* ------------------------------------------------------------------ */
/* Whole .ambc file  (adaptive_compressor.py:221-255, 312-358)          */
/* ------------------------------------------------------------------ */

/*
 * Header layout (:312-325): 'AMBC', version 2, header_size u32, marker_len
 * bits u8, marker bytes, checksum_type 1, md5[16], orig u64, comp u64.
 * out capacity: header + body capacity of orc_compress_body.
 * Returns the file length.  *stored_raw = 1 when header+body > total and the
 * input is written verbatim (:241-247).
 */
ORC_API long orc_compress_file(const uint8_t *data, long total,
                               const long *cands, int n_cands,
                               const int *methods, int n_methods,
                               const uint8_t *marker_raw, int marker_bits,
                               int per_chunk_raw,
                               uint8_t *out, int *stored_raw,
                               int *map_type, long *map_orig, long *map_comp, long map_cap, long *n_pkgs)
{
    int mb = (marker_bits + 7) / 8;
    /* _init_marker (:196-219): first marker_bits bits, left aligned, zero padded */
    uint8_t aligned[8] = { 0 };
    for (int b = 0; b < marker_bits; b++)
        if ((marker_raw[b >> 3] >> (7 - (b & 7))) & 1) aligned[b >> 3] |= (uint8_t)(0x80 >> (b & 7));
    uint8_t *hdr = out;
    long h = 0;
    memcpy(hdr, "AMBC", 4); h = 4;
    hdr[h++] = 2;
    h += 4; /* header size, patched below */
    hdr[h++] = (uint8_t)marker_bits;
    memcpy(hdr + h, marker_raw, (size_t)mb); h += mb; /* marker_bytes as found (:318) */
    hdr[h++] = 1;
    orc_md5(data, total, hdr + h); h += 16;
    for (int k = 0; k < 8; k++) hdr[h++] = (uint8_t)((uint64_t)total >> (8 * k));
    long comp_pos = h;
    h += 8;
    put_u32(hdr + 5, (uint32_t)h);
    long blen = orc_compress_body(data, total, cands, n_cands, methods, n_methods, aligned, mb,
                                  per_chunk_raw, out + h, map_type, map_orig, map_comp, map_cap, n_pkgs);
    if (h + blen > total) { /* :241-247 */
        memcpy(out, data, (size_t)total);
        *stored_raw = 1;
        return total;
    }
    for (int k = 0; k < 8; k++) hdr[comp_pos + k] = (uint8_t)((uint64_t)blen >> (8 * k));
    *stored_raw = 0;
    return h + blen;
}
