// index.cu -- GPU-side package index: the table ambc_index_host builds by walking the package chain
// on the host (adaptive_compressor.py:396-430, :444-445), built on the device instead.
//
// The chain is pointer chasing (pos += header + comp_len), but every package starts with the
// marker, so:
//   1. k_idx_scan   finds every occurrence of the marker bytes in the body (coalesced loads, matches
//                   into a shared-memory bitmap, ordered compaction) -> sorted candidate positions.
//                   The marker is NOT guaranteed absent from payloads, so candidates
//                   are a superset of the package starts;
//   2. k_idx_link   parses the header at every candidate and finds its successor (the candidate at
//                   pos + header + comp_len, by binary search), or why the walk would end there;
//   3. k_idx_jump   builds jump tables succ^(2^k); k_idx_end walks them from candidate 0 to the
//                   chain length, k_idx_chain lists the chain in order (binary lifting) -- only
//                   candidates reachable from position 0 are packages, exactly as in the serial walk;
//   4. scans of the nominal output sizes place every package (the walk stops once the output is
//                   complete, :444-445), raw packages are split into 64 KiB entries, and
//                   k_idx_emit writes the table.
// A marker mismatch the serial walk would hit is reported as AMBC_E_MARKER.
#include "ambc_internal.h"
#include "common.cuh"

#define IDX_TILE 16384   // body bytes per CTA in the marker scan
#define IDX_THREADS 256
#define IDX_RAW_PIECE 65536u
#define SC_TILE 2048     // elements per CTA in the device-wide scans

// ---- device-wide exclusive scan of uint64 (three small kernels) ----------------------------------
__global__ void __launch_bounds__(256) k_sc_tiles(const unsigned long long *__restrict__ v, uint64_t n, unsigned long long *tile)
{
    __shared__ unsigned long long red[8];
    const uint64_t base = (uint64_t)blockIdx.x * SC_TILE;
    unsigned long long s = 0;
    for (int k = threadIdx.x; k < SC_TILE; k += 256)
        if (base + k < n) s += v[base + k];
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(FULL_MASK, s, d);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int i = 0; i < 8; i++) t += red[i];
        tile[blockIdx.x] = t;
    }
}
__global__ void __launch_bounds__(1024) k_sc_top(unsigned long long *tile, uint64_t n_tiles, unsigned long long *total)
{
    __shared__ unsigned long long wsum[32];
    __shared__ unsigned long long carry_s;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (uint64_t base = 0; base < n_tiles; base += 1024) {
        const uint64_t i = base + tid;
        unsigned long long v = i < n_tiles ? tile[i] : 0, inc = v;
        for (int d = 1; d < 32; d <<= 1) {
            unsigned long long t = __shfl_up_sync(FULL_MASK, inc, d);
            if (lane >= d) inc += t;
        }
        if (lane == 31) wsum[w] = inc;
        __syncthreads();
        unsigned long long wbase = 0;
        for (int k = 0; k < w; k++) wbase += wsum[k];
        const unsigned long long carry = carry_s;
        if (i < n_tiles) tile[i] = carry + wbase + inc - v;
        __syncthreads();
        if (tid == 1023) carry_s = carry + wbase + inc;
        __syncthreads();
    }
    if (tid == 0) *total = carry_s;
}
__global__ void __launch_bounds__(256) k_sc_apply(const unsigned long long *__restrict__ v, uint64_t n,
                                                  const unsigned long long *__restrict__ tile, unsigned long long *out)
{
    __shared__ unsigned long long wsum[8];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const uint64_t base = (uint64_t)blockIdx.x * SC_TILE + (uint64_t)tid * 8;
    unsigned long long x[8], s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) { x[k] = base + k < n ? v[base + k] : 0; s += x[k]; }
    unsigned long long inc = s;
    for (int d = 1; d < 32; d <<= 1) {
        unsigned long long t = __shfl_up_sync(FULL_MASK, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    unsigned long long run = tile[blockIdx.x] + inc - s;
    for (int k = 0; k < w; k++) run += wsum[k];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        if (base + k < n) out[base + k] = run;
        run += x[k];
    }
}
// out[i] = sum of v[0..i); *total = sum of all.  tile: >= ceil(n / SC_TILE) + 1 words of scratch.
static int dev_excl_scan(const unsigned long long *v, uint64_t n, unsigned long long *out, unsigned long long *tile,
                         unsigned long long *total, cudaStream_t s)
{
    const uint64_t nt = (n + SC_TILE - 1) / SC_TILE;
    if (n == 0) { CUDA_TRY(cudaMemsetAsync(total, 0, 8, s)); return AMBC_OK; }
    k_sc_tiles<<<(unsigned)nt, 256, 0, s>>>(v, n, tile);
    k_sc_top<<<1, 1024, 0, s>>>(tile, nt, total);
    k_sc_apply<<<(unsigned)nt, 256, 0, s>>>(v, n, tile, out);
    for (int i = 0; i < 3; i++) ambc_count_launch();
    CUDA_TRY(cudaGetLastError());
    return AMBC_OK;
}

// ---- 1. marker scan -------------------------------------------------------------------------------
// WRITE == false: tile_cnt[tile] = matches in the tile; WRITE == true: positions written in order at
// tile_base[tile].  A match at p needs p + mb <= len.
template <bool WRITE>
__global__ void __launch_bounds__(IDX_THREADS)
k_idx_scan(const uint8_t *__restrict__ body, uint64_t len, uint32_t marker_word, uint32_t mb,
           unsigned long long *tile_cnt, const unsigned long long *__restrict__ tile_base, unsigned long long *cand)
{
    __shared__ uint32_t bits[IDX_TILE / 32];
    __shared__ int wsum[IDX_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const uint64_t t0 = (uint64_t)blockIdx.x * IDX_TILE;
    for (int i = tid; i < IDX_TILE / 32; i += IDX_THREADS) bits[i] = 0;
    __syncthreads();
    const uint32_t kmask = mb >= 4 ? 0xFFFFFFFFu : ((1u << (8 * mb)) - 1u);
    if (((uintptr_t)body & 15) == 0 && t0 + IDX_TILE + 16 <= len) {
        // full tile, aligned: 16 bytes (+ one halo word) per thread and step, 16 windows each
        const uint4 *b4 = (const uint4 *)(body + t0);
        for (int j = 0; j < IDX_TILE / 16 / IDX_THREADS; j++) {
            const int v = j * IDX_THREADS + tid;
            const uint4 x = __ldg(b4 + v);
            const uint32_t halo = __ldg((const uint32_t *)(b4 + v + 1));
            const uint32_t w[5] = {x.x, x.y, x.z, x.w, halo};
            uint32_t m16 = 0;
#pragma unroll
            for (int b = 0; b < 16; b++) {
                const uint32_t win = __funnelshift_r(w[b >> 2], w[(b >> 2) + 1], (b & 3) * 8);
                if ((win & kmask) == marker_word) m16 |= 1u << b;
            }
            if (m16) atomicOr(&bits[v >> 1], m16 << ((v & 1) * 16));
        }
    } else {
        for (int j = 0; j < IDX_TILE / IDX_THREADS; j++) {
            const uint32_t q = (uint32_t)j * IDX_THREADS + tid; // coalesced: consecutive lanes, consecutive bytes
            const uint64_t p = t0 + q;
            if (p + mb <= len) {
                uint32_t v = 0;
                for (uint32_t k = 0; k < mb; k++) v |= (uint32_t)__ldg(body + p + k) << (8 * k);
                if ((v & kmask) == marker_word) atomicOr(&bits[q >> 5], 1u << (q & 31));
            }
        }
    }
    __syncthreads();
    // thread t owns words 2t, 2t+1 (64 consecutive positions)
    const uint32_t b0 = bits[2 * tid], b1 = bits[2 * tid + 1];
    const int mine = __popc(b0) + __popc(b1);
    int inc = mine;
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(FULL_MASK, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    int base = inc - mine, tot = 0;
    for (int k = 0; k < IDX_THREADS / 32; k++) { if (k < w) base += wsum[k]; tot += wsum[k]; }
    if (!WRITE) {
        if (tid == 0) tile_cnt[blockIdx.x] = (unsigned long long)tot;
    } else {
        unsigned long long *dst = cand + tile_base[blockIdx.x] + base;
        uint32_t x = b0;
        while (x) { const int b = __ffs(x) - 1; x &= x - 1; *dst++ = t0 + 64ull * tid + b; }
        x = b1;
        while (x) { const int b = __ffs(x) - 1; x &= x - 1; *dst++ = t0 + 64ull * tid + 32 + b; }
    }
}

// ---- 2. headers and successors ----------------------------------------------------------------------
struct IdxRec {
    uint32_t orig, comp;
    uint8_t type, kind; // kind: 0 package, 1 END, 2 header cut off, 3 payload cut off
    uint16_t pad;
};
#define IDX_STOP(M) ((uint32_t)(M))      // the walk ends here without error
#define IDX_ERR(M) ((uint32_t)(M) + 1u)  // the walk raises "Marker mismatch" at the next position

__global__ void __launch_bounds__(256)
k_idx_link(const uint8_t *__restrict__ body, uint64_t len, uint32_t mb, const unsigned long long *__restrict__ cand,
           uint32_t M, IdxRec *rec, uint32_t *succ)
{
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= M + 2) return;
    if (i >= M) { succ[i] = i; return; } // absorbing
    const uint64_t hdr = mb + 14, pos = cand[i];
    IdxRec r;
    r.orig = r.comp = 0; r.type = 0; r.kind = 0; r.pad = 0;
    uint32_t sx = IDX_STOP(M);
    if (pos + hdr > len) r.kind = 2;                                    // :400-403
    else {
        r.type = body[pos + mb];
        uint32_t o = 0, c = 0;
        for (int k = 0; k < 4; k++) { o |= (uint32_t)body[pos + mb + 6 + k] << (8 * k); c |= (uint32_t)body[pos + mb + 10 + k] << (8 * k); }
        r.orig = o; r.comp = c;
        const uint64_t nxt = pos + hdr + c;
        if (r.type == 0) r.kind = 1;                                    // :422-424
        else if (nxt > len) r.kind = 3;                                 // :425-427
        else if (nxt + hdr > len) sx = IDX_STOP(M);                     // the next iteration stops at :400-403 (or pos == len)
        else {
            uint32_t lo = i + 1, hi = M; // candidates are sorted; nxt > pos
            while (lo < hi) { const uint32_t mid = lo + (hi - lo) / 2; if (cand[mid] < nxt) lo = mid + 1; else hi = mid; }
            sx = (lo < M && cand[lo] == nxt) ? lo : IDX_ERR(M);         // :405-407
        }
    }
    rec[i] = r;
    succ[i] = sx;
}

// ---- 3. jump tables, chain ----------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_idx_jump(const uint32_t *__restrict__ prev, uint32_t *next, uint32_t n)
{
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i < n) next[i] = prev[prev[i]];
}
// info[0] = chain length (nodes visited from candidate 0), info[1] = 1 when the walk ends in a marker mismatch
__global__ void k_idx_end(const uint32_t *__restrict__ J, uint32_t n, int K, uint32_t M, unsigned long long *info)
{
    if (threadIdx.x || blockIdx.x) return;
    uint32_t node = 0;
    unsigned long long steps = 0;
    for (int k = K - 1; k >= 0; k--) {
        const uint32_t nx = J[(uint64_t)k * n + node];
        if (nx < M) { node = nx; steps += 1ull << k; }
    }
    info[0] = steps + 1;
    info[1] = J[node] == IDX_ERR(M) ? 1 : 0;
}
__global__ void __launch_bounds__(256)
k_idx_chain(const uint32_t *__restrict__ J, uint32_t n, int K, uint64_t chain_len, uint32_t *chain)
{
    const uint64_t r = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (r >= chain_len) return;
    uint32_t node = 0;
    for (int k = 0; k < K; k++)
        if ((r >> k) & 1) node = J[(uint64_t)k * n + node];
    chain[r] = node;
}

// ---- 4. placement and table ------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t idx_nominal(uint32_t type, bool known, uint32_t comp, uint32_t orig)
{
    if (!known) return comp;                       // copied through (:432-435)
    switch (type) {
    case 255: return orig;                         // pad / truncate (compression_methods.py:703-713)
    case 4: return comp == 0 ? 0 : (comp < orig ? comp : orig); // :621-638
    default: return comp == 0 ? 0 : orig;          // `if not data: return b''`
    }
}
__device__ __forceinline__ bool idx_known(uint32_t type, uint32_t known_mask)
{
    return type == 255 || (type < 32 && ((known_mask >> type) & 1u));
}
__global__ void __launch_bounds__(256)
k_idx_nominal(const IdxRec *__restrict__ rec, const uint32_t *__restrict__ chain, uint64_t chain_len, uint32_t known_mask,
              unsigned long long *nominal)
{
    const uint64_t r = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (r >= chain_len) return;
    const IdxRec x = rec[chain[r]];
    nominal[r] = x.kind == 0 ? idx_nominal(x.type, idx_known(x.type, known_mask), x.comp, x.orig) : 0;
}
// cut = first chain rank after which the walk stops because the output is complete (:444-445)
__global__ void __launch_bounds__(256)
k_idx_cut(const unsigned long long *__restrict__ nominal, const unsigned long long *__restrict__ obefore, uint64_t chain_len,
          uint64_t orig_size, unsigned long long *cut)
{
    const uint64_t r = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (r >= chain_len) return;
    if (obefore[r] + nominal[r] >= orig_size) atomicMin(cut, (unsigned long long)r);
}
__global__ void __launch_bounds__(256)
k_idx_count(const IdxRec *__restrict__ rec, const uint32_t *__restrict__ chain, const unsigned long long *__restrict__ nominal,
            const unsigned long long *__restrict__ obefore, uint64_t chain_len, uint64_t orig_size, uint32_t known_mask,
            const unsigned long long *__restrict__ cut, unsigned long long *ecount)
{
    const uint64_t r = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (r >= chain_len) return;
    unsigned long long e = 0;
    const IdxRec x = rec[chain[r]];
    if (r <= *cut && x.kind == 0) {
        const uint64_t o = obefore[r], room = orig_size > o ? orig_size - o : 0;
        const uint64_t emit = nominal[r] < room ? nominal[r] : room;
        if (emit) e = (!idx_known(x.type, known_mask) || x.type == 255) ? (emit + IDX_RAW_PIECE - 1) / IDX_RAW_PIECE : 1;
    }
    ecount[r] = e;
}
__global__ void __launch_bounds__(256)
k_idx_emit(const IdxRec *__restrict__ rec, const uint32_t *__restrict__ chain, const unsigned long long *__restrict__ cand,
           const unsigned long long *__restrict__ nominal, const unsigned long long *__restrict__ obefore,
           const unsigned long long *__restrict__ ecount, const unsigned long long *__restrict__ ebase, uint64_t chain_len,
           uint64_t orig_size, uint32_t known_mask, uint32_t mb, ambc_pkg *table, uint64_t table_cap)
{
    const uint64_t r = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (r >= chain_len || ecount[r] == 0) return;
    const IdxRec x = rec[chain[r]];
    const uint64_t pos = cand[chain[r]] + mb + 14; // payload offset
    const uint64_t o = obefore[r], room = orig_size - o;
    const uint64_t emit = nominal[r] < room ? nominal[r] : room;
    uint64_t at = ebase[r];
    if (!idx_known(x.type, known_mask) || x.type == 255) {
        for (uint64_t done = 0; done < emit; done += IDX_RAW_PIECE, at++) {
            if (at >= table_cap) return;
            const uint64_t piece = emit - done < IDX_RAW_PIECE ? emit - done : IDX_RAW_PIECE;
            const uint64_t have = x.comp > done ? x.comp - done : 0;
            ambc_pkg e;
            e.src_off = pos + done; e.dst_off = o + done;
            e.comp_len = (uint32_t)(have < piece ? have : piece);
            e.orig_len = (uint32_t)piece; e.type = 255; e.out_len = (uint32_t)piece;
            table[at] = e;
        }
    } else if (at < table_cap) {
        ambc_pkg e;
        e.src_off = pos; e.dst_off = o; e.comp_len = x.comp; e.orig_len = x.orig; e.type = x.type; e.out_len = (uint32_t)emit;
        table[at] = e;
    }
}

// ---- host side ------------------------------------------------------------------------------------------------
struct IdxBuf {
    void *p = nullptr;
    uint64_t cap = 0;
    int ensure(uint64_t bytes)
    {
        if (bytes <= cap) return AMBC_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        const uint64_t want = bytes + bytes / 4 + 4096;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { p = nullptr; return ambc_fail(AMBC_E_CUDA, "cudaMalloc(%llu): %s", (unsigned long long)want, cudaGetErrorString(e)); }
        cap = want;
        return AMBC_OK;
    }
};
static IdxBuf g_idx_a[16], g_idx_b[16];

extern "C" int ambc_index_dev(const void *body_dev, uint64_t body_len, const uint8_t *marker, uint32_t mb, uint64_t orig_size,
                              uint32_t known_mask, ambc_pkg *table_dev, uint64_t table_cap, uint64_t *n_entries,
                              uint64_t *out_bytes, void *stream_)
{
    cudaStream_t s = (cudaStream_t)stream_;
    if (!marker || mb < 1 || mb > 4 || (body_len && !body_dev)) return ambc_fail(AMBC_E_ARG, "ambc_index_dev: bad argument");
    if (n_entries) *n_entries = 0;
    if (out_bytes) *out_bytes = 0;
    const uint64_t hdr = mb + 14;
    if (body_len < hdr) return AMBC_OK; // the walk stops before the first header (:400-403)
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 16) return ambc_fail(AMBC_E_ARG, "device index out of range");
    uint32_t marker_word = 0;
    for (uint32_t k = 0; k < mb; k++) marker_word |= (uint32_t)marker[k] << (8 * k);
    const uint8_t *body = (const uint8_t *)body_dev;

    // 1. candidates
    const uint64_t n_tiles = (body_len + IDX_TILE - 1) / IDX_TILE;
    IdxBuf &A = g_idx_a[dev], &B = g_idx_b[dev];
    int rc = A.ensure((n_tiles + 2) * 8 * 2 + ((n_tiles + SC_TILE) / SC_TILE + 2) * 8 + 64);
    if (rc) return rc;
    unsigned long long *tile_cnt = (unsigned long long *)A.p, *tile_base = tile_cnt + n_tiles + 1;
    unsigned long long *sc_tile = tile_base + n_tiles + 1, *d_total = sc_tile + (n_tiles + SC_TILE) / SC_TILE + 1;
    k_idx_scan<false><<<(unsigned)n_tiles, IDX_THREADS, 0, s>>>(body, body_len, marker_word, mb, tile_cnt, nullptr, nullptr);
    ambc_count_launch();
    if ((rc = dev_excl_scan(tile_cnt, n_tiles, tile_base, sc_tile, d_total, s))) return rc;
    unsigned long long h_total = 0;
    CUDA_TRY(cudaMemcpyAsync(&h_total, d_total, 8, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    const uint64_t M64 = h_total;
    if (M64 == 0) return ambc_fail(AMBC_E_MARKER, "Marker mismatch in chunk header."); // position 0 is not a marker
    if (M64 > 0x7FFFFFF0ull) return ambc_fail(AMBC_E_CAPACITY, "ambc_index_dev: too many marker occurrences");
    const uint32_t M = (uint32_t)M64, n = M + 2;
    int K = 1;
    while ((1ull << K) <= M) K++;
    // layout of B: cand | rec | J[K][n] | chain | nominal | obefore | ecount | ebase | scan tiles | info
    const uint64_t sc_tiles = (M64 + SC_TILE) / SC_TILE + 2;
    uint64_t off = 0;
    auto take = [&](uint64_t bytes) { uint64_t r = off; off += (bytes + 255) & ~255ull; return r; };
    const uint64_t o_cand = take(M64 * 8), o_rec = take(M64 * sizeof(IdxRec)), o_J = take((uint64_t)K * n * 4), o_chain = take(M64 * 4),
                   o_nom = take(M64 * 8), o_ob = take(M64 * 8), o_ec = take(M64 * 8), o_eb = take(M64 * 8), o_sct = take(sc_tiles * 8),
                   o_info = take(64);
    if ((rc = B.ensure(off))) return rc;
    uint8_t *W = (uint8_t *)B.p;
    unsigned long long *cand = (unsigned long long *)(W + o_cand);
    IdxRec *rec = (IdxRec *)(W + o_rec);
    uint32_t *J = (uint32_t *)(W + o_J), *chain = (uint32_t *)(W + o_chain);
    unsigned long long *nominal = (unsigned long long *)(W + o_nom), *obefore = (unsigned long long *)(W + o_ob);
    unsigned long long *ecount = (unsigned long long *)(W + o_ec), *ebase = (unsigned long long *)(W + o_eb);
    unsigned long long *sct = (unsigned long long *)(W + o_sct), *info = (unsigned long long *)(W + o_info);
    k_idx_scan<true><<<(unsigned)n_tiles, IDX_THREADS, 0, s>>>(body, body_len, marker_word, mb, nullptr, tile_base, cand);
    ambc_count_launch();
    unsigned long long first = 1;
    CUDA_TRY(cudaMemcpyAsync(&first, cand, 8, cudaMemcpyDeviceToHost, s));

    // 2. + 3. links, jump tables, chain
    k_idx_link<<<(n + 255) / 256, 256, 0, s>>>(body, body_len, mb, cand, M, rec, J);
    ambc_count_launch();
    for (int k = 1; k < K; k++) {
        k_idx_jump<<<(n + 255) / 256, 256, 0, s>>>(J + (uint64_t)(k - 1) * n, J + (uint64_t)k * n, n);
        ambc_count_launch();
    }
    k_idx_end<<<1, 1, 0, s>>>(J, n, K, M, info);
    ambc_count_launch();
    unsigned long long h_info[4] = {0, 0, 0, 0};
    CUDA_TRY(cudaMemcpyAsync(h_info, info, 16, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    if (first != 0) return ambc_fail(AMBC_E_MARKER, "Marker mismatch in chunk header."); // position 0 is not a marker
    const uint64_t chain_len = h_info[0];
    const bool ends_in_mismatch = h_info[1] != 0;
    const unsigned cgrid = (unsigned)((chain_len + 255) / 256);
    k_idx_chain<<<cgrid, 256, 0, s>>>(J, n, K, chain_len, chain);
    ambc_count_launch();

    // 4. placement
    k_idx_nominal<<<cgrid, 256, 0, s>>>(rec, chain, chain_len, known_mask, nominal);
    ambc_count_launch();
    if ((rc = dev_excl_scan(nominal, chain_len, obefore, sct, info + 2, s))) return rc;
    const unsigned long long none = ~0ull;
    CUDA_TRY(cudaMemcpyAsync(info + 3, &none, 8, cudaMemcpyHostToDevice, s));
    k_idx_cut<<<cgrid, 256, 0, s>>>(nominal, obefore, chain_len, orig_size, info + 3);
    k_idx_count<<<cgrid, 256, 0, s>>>(rec, chain, nominal, obefore, chain_len, orig_size, known_mask, info + 3, ecount);
    ambc_count_launch(); ambc_count_launch();
    if ((rc = dev_excl_scan(ecount, chain_len, ebase, sct, info + 4, s))) return rc;
    unsigned long long h_res[3] = {0, 0, 0}; // total nominal, cut, entries
    CUDA_TRY(cudaMemcpyAsync(h_res, info + 2, 24, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    const uint64_t cut = h_res[1], ne = h_res[2];
    if (ends_in_mismatch && cut == ~0ull) return ambc_fail(AMBC_E_MARKER, "Marker mismatch in chunk header.");
    // bytes the walk produced: everything up to and including the cut package
    uint64_t produced = h_res[0];
    if (cut != ~0ull) {
        unsigned long long ob = 0, nm = 0;
        CUDA_TRY(cudaMemcpyAsync(&ob, obefore + cut, 8, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaMemcpyAsync(&nm, nominal + cut, 8, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        produced = ob + nm;
    }
    if (n_entries) *n_entries = ne;
    if (out_bytes) *out_bytes = produced < orig_size ? produced : orig_size;
    if (table_dev) {
        if (ne > table_cap) return ambc_fail(AMBC_E_CAPACITY, "ambc_index_dev: table too small (%llu entries)", (unsigned long long)ne);
        k_idx_emit<<<cgrid, 256, 0, s>>>(rec, chain, cand, nominal, obefore, ecount, ebase, chain_len, orig_size, known_mask, mb,
                                        table_dev, table_cap);
        ambc_count_launch();
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaStreamSynchronize(s));
    }
    return AMBC_OK;
}
