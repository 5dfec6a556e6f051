/*
 * oracle/oracle_mt.c -- multi-threaded driver around the oracle for the CPU
 * baseline (bench.py cpu_baseline / --impl reference).  TEST INFRASTRUCTURE.
 *
 * The reference is single-threaded (adaptive_compressor.py:186-194 only sets a
 * flag), but in fixed-candidate mode chunks are independent until the tail-raw
 * rule fires, so the port can use every host core: each thread runs the
 * reference's per-chunk trial loop (orc_compress_body on one chunk-aligned
 * slice).  Returns total body bytes over all slices (each slice carries its
 * own END package; the figure is only used for timing, not for parity).
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>

#define ORC_API __attribute__((visibility("default")))

long orc_compress_body(const uint8_t *data, long total, const long *cands, int n_cands,
                       const int *methods, int n_methods, const uint8_t *marker, int marker_bytes,
                       int per_chunk_raw, uint8_t *out, int *map_type, long *map_orig, long *map_comp,
                       long map_cap, long *n_pkgs);
long orc_decompress_body(const uint8_t *body, long blen, long orig_size, const uint8_t *marker,
                         int marker_bytes, const int *known, int n_known, uint8_t *out);

typedef struct {
    const uint8_t *data; long total; long chunk; const int *methods; int n_methods;
    uint8_t *out; long out_len; long n_pkgs;
} slice_job;

static const uint8_t k_marker[4] = { 0xFF, 0xFF, 0x00, 0x00 };

static void *compress_worker(void *arg)
{
    slice_job *j = (slice_job *)arg;
    long cands[1] = { j->chunk };
    j->out_len = orc_compress_body(j->data, j->total, cands, 1, j->methods, j->n_methods, k_marker, 4, 0,
                                   j->out, NULL, NULL, NULL, 0, &j->n_pkgs);
    return NULL;
}

/* Compress `total` bytes as `threads` chunk-aligned slices in parallel.
 * outs[t] must hold slice_len + (slice_len/chunk + 2) * 18 + 16 bytes. */
ORC_API long orc_mt_compress(const uint8_t *data, long total, long chunk, const int *methods, int n_methods,
                             int threads, uint8_t **outs, long *out_lens)
{
    if (threads < 1) threads = 1;
    long n_chunks = (total + chunk - 1) / chunk;
    long per = (n_chunks + threads - 1) / threads;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    slice_job *jobs = (slice_job *)calloc((size_t)threads, sizeof(slice_job));
    int started = 0;
    for (int t = 0; t < threads; t++) {
        long c0 = t * per, c1 = c0 + per; if (c1 > n_chunks) c1 = n_chunks;
        if (c0 >= c1) break;
        long b0 = c0 * chunk, b1 = c1 * chunk; if (b1 > total) b1 = total;
        jobs[t].data = data + b0; jobs[t].total = b1 - b0; jobs[t].chunk = chunk;
        jobs[t].methods = methods; jobs[t].n_methods = n_methods; jobs[t].out = outs[t];
        pthread_create(&th[t], NULL, compress_worker, &jobs[t]);
        started++;
    }
    long sum = 0;
    for (int t = 0; t < started; t++) { pthread_join(th[t], NULL); out_lens[t] = jobs[t].out_len; sum += jobs[t].out_len; }
    for (int t = started; t < threads; t++) out_lens[t] = 0;
    free(th); free(jobs);
    return sum;
}

typedef struct { const uint8_t *body; long blen; long orig; uint8_t *out; long rc; } dslice_job;

static void *decompress_worker(void *arg)
{
    dslice_job *j = (dslice_job *)arg;
    int known[5] = { 1, 2, 3, 4, 255 };
    j->rc = orc_decompress_body(j->body, j->blen, j->orig, k_marker, 4, known, 5, j->out);
    return NULL;
}

/* Decode the slices produced by orc_mt_compress in parallel. */
ORC_API long orc_mt_decompress(uint8_t **bodies, const long *body_lens, const long *orig_lens, int threads,
                               uint8_t **outs)
{
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    dslice_job *jobs = (dslice_job *)calloc((size_t)threads, sizeof(dslice_job));
    int started = 0;
    for (int t = 0; t < threads; t++) {
        if (orig_lens[t] <= 0) break;
        jobs[t].body = bodies[t]; jobs[t].blen = body_lens[t]; jobs[t].orig = orig_lens[t]; jobs[t].out = outs[t];
        pthread_create(&th[t], NULL, decompress_worker, &jobs[t]);
        started++;
    }
    long sum = 0;
    for (int t = 0; t < started; t++) { pthread_join(th[t], NULL); sum += jobs[t].rc; }
    free(th); free(jobs);
    return sum;
}
