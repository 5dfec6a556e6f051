// api.cu -- error reporting, bookkeeping and the host-buffer entry points of libambc.
#include "ambc_internal.h"
#include <atomic>
#include <cstdarg>
#include <mutex>
#include <vector>

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

int ambc_fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}
void ambc_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

static AmbcTiming g_timing;
AmbcTiming &ambc_timing() { return g_timing; }
void ambc_timing_mark(int idx, cudaStream_t s)
{
    if (!g_timing.on) return;
    if (!g_timing.ev[idx]) cudaEventCreate(&g_timing.ev[idx]);
    cudaEventRecord(g_timing.ev[idx], s);
}
extern "C" void ambc_enable_timing(int on) { g_timing.on = on != 0; }
extern "C" int ambc_last_timing(float *ms4)
{
    AmbcTiming &t = g_timing;
    if (!ms4) return AMBC_E_ARG;
    if (t.pending_c && t.ev[0] && t.ev[3]) {
        cudaEventSynchronize(t.ev[3]);
        cudaEventElapsedTime(&t.ms[0], t.ev[0], t.ev[1]);
        cudaEventElapsedTime(&t.ms[1], t.ev[1], t.ev[2]);
        cudaEventElapsedTime(&t.ms[2], t.ev[2], t.ev[3]);
        t.pending_c = false;
    }
    if (t.pending_d && t.ev[4] && t.ev[5]) {
        cudaEventSynchronize(t.ev[5]);
        cudaEventElapsedTime(&t.ms[3], t.ev[4], t.ev[5]);
        t.pending_d = false;
    }
    for (int i = 0; i < 4; i++) ms4[i] = t.ms[i];
    return AMBC_OK;
}

int ambc_lz_levels_compress(const int *levels, int n);
int ambc_lz_levels_codec(const int *levels, int n);
// experiment knob (not part of include/ambc.h): n-gram levels of the Dictionary match search
extern "C" int ambc_set_lz_levels(const int *levels, int n)
{
    if (ambc_lz_levels_compress(levels, n) || ambc_lz_levels_codec(levels, n))
        return ambc_fail(AMBC_E_ARG, "ambc_set_lz_levels: need ascending levels starting at 3, each <= 16, at most 8");
    return AMBC_OK;
}

int ambc_lz_coop_compress(int t);
int ambc_lz_coop_codec(int t);
extern "C" int ambc_set_lz_coop_threshold(int t) { return (ambc_lz_coop_compress(t) || ambc_lz_coop_codec(t)) ? AMBC_E_CUDA : AMBC_OK; }

int ambc_lz_force_buckets_compress(int on);
int ambc_lz_force_buckets_codec(int on);
// test knob (not part of include/ambc.h): force the window-aware bucket search for every chunk size
extern "C" int ambc_set_lz_force_buckets(int on)
{
    return (ambc_lz_force_buckets_compress(on) || ambc_lz_force_buckets_codec(on)) ? AMBC_E_CUDA : AMBC_OK;
}

extern "C" const char *ambc_last_error(void) { return g_err; }
extern "C" int ambc_version(void) { return 100; }
extern "C" uint64_t ambc_launch_count(void) { return g_launches.load(); }
extern "C" int ambc_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return ambc_fail(AMBC_E_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    return n;
}
extern "C" void *ambc_host_alloc(uint64_t bytes)
{
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { ambc_fail(AMBC_E_CUDA, "cudaMallocHost failed"); return nullptr; }
    return p;
}
extern "C" void ambc_host_free(void *p) { if (p) cudaFreeHost(p); }

// ---- cached device buffers for the *_host calls (one set per device) ----------------------------
struct DevBuf {
    void *p = nullptr;
    uint64_t cap = 0;
    int ensure(uint64_t bytes)
    {
        if (bytes <= cap) return AMBC_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        uint64_t want = bytes + bytes / 8 + 4096;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { p = nullptr; return ambc_fail(AMBC_E_CUDA, "cudaMalloc(%llu): %s", (unsigned long long)want, cudaGetErrorString(e)); }
        cap = want;
        return AMBC_OK;
    }
};
struct HostCtx {
    DevBuf in, out, work, table, status;
    cudaStream_t stream = nullptr;
    std::vector<ambc_pkg> host_table;
};
int ambc_index_vector(const uint8_t *body, uint64_t body_len, const uint8_t *marker, uint32_t mb, uint64_t orig_size,
                      uint32_t known_mask, std::vector<ambc_pkg> &v, uint64_t *n_entries, uint64_t *out_bytes);
static std::mutex g_mu;
static HostCtx g_ctx[16];

static int host_ctx(HostCtx **out)
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return ambc_fail(AMBC_E_CUDA, "cudaGetDevice: %s (no CUDA device? libambc has no CPU fallback)", cudaGetErrorString(e));
    if (dev < 0 || dev >= 16) return ambc_fail(AMBC_E_ARG, "device index out of range");
    HostCtx *c = &g_ctx[dev];
    if (!c->stream) {
        e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) return ambc_fail(AMBC_E_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
    }
    *out = c;
    return AMBC_OK;
}

extern "C" int ambc_compress_host(const void *in_host, uint64_t n, uint32_t chunk, uint32_t method_mask,
                                  uint32_t flags, const uint8_t *marker, uint32_t marker_bytes, void *out_host,
                                  uint64_t out_cap, uint8_t *map_type, uint32_t *map_comp, ambc_compress_result *res)
{
    std::lock_guard<std::mutex> lk(g_mu);
    HostCtx *c;
    int rc = host_ctx(&c);
    if (rc) return rc;
    if (!res || chunk == 0) return ambc_fail(AMBC_E_ARG, "ambc_compress_host: bad argument");
    uint64_t bound = ambc_compress_bound(n, chunk, marker_bytes);
    uint64_t wbytes = ambc_compress_workspace_bytes(n, chunk);
    if ((rc = c->in.ensure(n + 64))) return rc;
    if ((rc = c->out.ensure(bound))) return rc;
    if ((rc = c->work.ensure(wbytes))) return rc;
    if (n) CUDA_TRY(cudaMemcpyAsync(c->in.p, in_host, n, cudaMemcpyHostToDevice, c->stream));
    rc = ambc_compress_dev(c->in.p, n, chunk, method_mask, flags, marker, marker_bytes, c->out.p, bound, c->work.p,
                           c->work.cap, res, c->stream);
    if (rc) return rc;
    if (res->body_len > out_cap) return ambc_fail(AMBC_E_CAPACITY, "ambc_compress_host: out_cap %llu < body %llu",
                                                  (unsigned long long)out_cap, (unsigned long long)res->body_len);
    CUDA_TRY(cudaMemcpyAsync(out_host, c->out.p, res->body_len, cudaMemcpyDeviceToHost, c->stream));
    if (map_type && res->n_chunks)
        CUDA_TRY(cudaMemcpyAsync(map_type, (uint8_t *)c->work.p + res->map_type_off, res->n_chunks,
                                 cudaMemcpyDeviceToHost, c->stream));
    if (map_comp && res->n_chunks)
        CUDA_TRY(cudaMemcpyAsync(map_comp, (uint8_t *)c->work.p + res->map_comp_off, res->n_chunks * 4,
                                 cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return AMBC_OK;
}

extern "C" int ambc_decompress_host(const void *body_host, uint64_t body_len, const uint8_t *marker,
                                    uint32_t marker_bytes, uint32_t known_mask, void *out_host, uint64_t orig_size,
                                    uint32_t *status)
{
    std::lock_guard<std::mutex> lk(g_mu);
    HostCtx *c;
    int rc = host_ctx(&c);
    if (rc) return rc;
    uint64_t ne = 0, covered = 0;
    // start the body upload first; the index walk on the host overlaps with it
    if ((rc = c->in.ensure(body_len + 64))) return rc;
    if (body_len) CUDA_TRY(cudaMemcpyAsync(c->in.p, body_host, body_len, cudaMemcpyHostToDevice, c->stream));
    rc = ambc_index_vector((const uint8_t *)body_host, body_len, marker, marker_bytes, orig_size, known_mask,
                           c->host_table, &ne, &covered);
    if (rc) { cudaStreamSynchronize(c->stream); return rc; }
    if ((rc = c->out.ensure(orig_size + 64))) return rc;
    if ((rc = c->table.ensure(ne * sizeof(ambc_pkg) + 64))) return rc;
    if ((rc = c->status.ensure(64))) return rc;
    CUDA_TRY(cudaMemsetAsync(c->status.p, 0, 8, c->stream));
    if (ne) CUDA_TRY(cudaMemcpyAsync(c->table.p, c->host_table.data(), ne * sizeof(ambc_pkg), cudaMemcpyHostToDevice, c->stream));
    rc = ambc_decompress_dev(c->in.p, body_len, (const ambc_pkg *)c->table.p, ne, c->out.p, orig_size,
                             (uint32_t *)c->status.p, c->stream);
    if (rc) return rc;
    if (orig_size) CUDA_TRY(cudaMemcpyAsync(out_host, c->out.p, orig_size, cudaMemcpyDeviceToHost, c->stream));
    uint32_t st[2] = {0, 0};
    CUDA_TRY(cudaMemcpyAsync(st, c->status.p, 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (status) { status[0] = st[0]; status[1] = st[1]; }
    return AMBC_OK;
}
