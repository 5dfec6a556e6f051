// dev tool (CPU): statistics of the lazy bucket parse on one chunk -- how many candidates a visited position
// enumerates, how many word compares, how many tokens.  Not product code.
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
static inline uint32_t ld32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static int g_mode = 0;
void lzsim_mode(int m) { g_mode = m; }
static inline uint32_t h3(const uint8_t *p, int hb) { return ((g_mode ? ld32(p) : (ld32(p) & 0xFFFFFFu)) * 2654435761u) >> (32 - hb); }
// out: [0] literals [1] matches [2] payload [3] candidates enumerated (visited match positions) [4] batches of G
// [5] word compares naive (full lcp each) [6] word compares with running best [7] has3 scan steps (bulk, all positions)
// [8] visited positions with has3 but final len<3 (hash collision only) [9] max bucket [10] batches with runbest early stop at cap
// [11] sum over visited of lcp words of best
void lzsim(const uint8_t *d0, int n, int hb, int G, long long *out)
{
    uint8_t *d = calloc(n + 64, 1); memcpy(d, d0, n);
    int nb = 1 << hb;
    int *cnt = calloc(nb + 1, sizeof(int)), *ord = malloc(n * sizeof(int)), *start = calloc(nb + 1, sizeof(int));
    int P = n - 2;
    for (int p = 0; p < P; p++) cnt[h3(d + p, hb)]++;
    int s = 0, mx = 0;
    for (int b = 0; b < nb; b++) { start[b] = s; s += cnt[b]; if (cnt[b] > mx) mx = cnt[b]; cnt[b] = 0; }
    start[nb] = s;
    for (int p = 0; p < P; p++) { uint32_t h = h3(d + p, hb); ord[start[h] + cnt[h]++] = p; }
    out[9] = mx;
    // bulk has3
    uint8_t *has3 = calloc(n, 1);
    for (int p = 0; p < P; p++) {
        uint32_t h = h3(d + p, hb); uint32_t w = ld32(d + p) & 0xFFFFFF;
        for (int i = start[h]; i < start[h + 1]; i++) {
            int q = ord[i]; out[7]++;
            if (q >= p) break;
            if ((ld32(d + q) & 0xFFFFFF) == w) { has3[p] = 1; break; }
        }
    }
    if (g_mode) { // exact has3 by brute force (a separate 3-gram index in the real thing)
        for (int p = 0; p < P; p++) { has3[p] = 0; for (int q = 0; q < p; q++) if ((ld32(d + q) & 0xFFFFFF) == (ld32(d + p) & 0xFFFFFF)) { has3[p] = 1; break; } }
    }
    int p = 0;
    while (p < n) {
        if (p >= P || !has3[p]) { out[0]++; out[2] += 2; p++; continue; }
        int cap = n - p < 32 ? n - p : 32;
        uint32_t h = h3(d + p, hb);
        int best = 0, bpos = 0, ncand = 0, stop = 0;
        for (int i = start[h]; i < start[h + 1] && !stop; i += G) {
            out[10]++;
            int bbest = best;
            for (int j = i; j < i + G && j < start[h + 1]; j++) {
                int q = ord[j];
                if (q >= p) { stop = 1; break; }
                ncand++;
                int l = 0; while (l < cap && d[q + l] == d[p + l]) l++;
                out[5] += l / 4 + 1;
                // running best (batch-level): check word containing byte `best` first
                if (best >= 3) {
                    int w0 = best / 4;
                    int ok = 1; for (int b = w0 * 4; b <= best && b < cap; b++) if (d[q + b] != d[p + b]) ok = 0;
                    if (!ok || best >= cap) out[6] += 1; else out[6] += 1 + l / 4 + 1;
                } else out[6] += l / 4 + 1;
                if (l > bbest) { bbest = l; bpos = q; }
            }
            best = bbest;
            if (best >= cap) stop = 1;
        }
        out[3] += ncand; out[4] += (ncand + G - 1) / G;
        (void)bpos;
        if (best >= 3) { out[1]++; out[2] += 4; out[11] += best / 4 + 1; p += best; }
        else { out[8]++; out[0]++; out[2] += 2; p++; }
    }
    free(d); free(cnt); free(ord); free(start); free(has3);
}
