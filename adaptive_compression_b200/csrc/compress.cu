// compress.cu -- chunk trial / select / pack kernels and the ambc_compress_* entry points.
//
// Replaces AdaptiveCompressor._adaptive_compress and its helpers
// (adaptive_compressor.py:363-394, 537-590, 595-700) in fixed-candidate mode.
//
//   k_select   one CTA per grid chunk: stage 4 KiB in shared memory, evaluate the gates,
//              compute every enabled candidate's exact size, pick the argmin in list order
//              with the benefit test len+overhead < n, encode the winner into its slot.
//   k_sizes / k_scan_blocks / k_offsets
//              exclusive scan of package sizes with the "rest of file raw" rule: everything
//              from the first chunk without a winner onward is ONE raw package.
//   k_pack     one CTA per chunk: 18-byte package header + payload to the final offset.
#define AMBC_BLOCK 512 // encoder CTAs: 16 warps per chunk, 2 CTAs per SM (100 KB of shared memory each)
#include "ambc_internal.h"
#include <vector>
#include "chunk_codec.cuh"
#include "select_fast.cuh"
#include <cstdlib>


// per-chunk decision ------------------------------------------------------------------------
struct SelectOut { int type; int len; };

// size-range eligibility (adaptive_compressor.py:114-127, tested on the clamped size :565-567)
__device__ __forceinline__ bool eligible(int id, int n)
{
    switch (id) {
    case 1: return n >= 32 && n <= 4096;
    case 2: return n >= 128 && n <= 8192;
    case 3: return n >= 32 && n <= 8192;
    case 4: return n >= 32 && n <= 4096;
    }
    return false;
}

// Evaluate one chunk already staged in c; leaves the winner's payload in c.pay.
__device__ SelectOut select_chunk(ChunkCtx &c, uint32_t mask, int ovh, unsigned int *trial = nullptr)
{
    const int n = c.n;
    ChunkFeatures f;
    PHASE_DECL
    chunk_features(c, f);
    PHASE(1);
    HuffScratch hs = huff_scratch(c);

    // gates (compression_methods.py:154-180, 315-343, 551-574)
    // (the fp64 quotients rep / (s - 1) > 0.3 and distinct / s < 0.8 as exact integer compares: the
    // operands are below 1000, so a quotient that is not exactly 3/10 or 8/10 is at least 1e-4 away
    // from it, and the exact ones round to the very doubles the literals denote)
    bool rle_ok = (mask & 2u) && eligible(1, n) && n >= 4 && 10 * f.rep > 3 * (min(1000, n) - 1);
    bool lz_ok = (mask & 4u) && eligible(2, n) && n >= 100 && 10 * f.distinct3 < 8 * min(1000, n);
    bool hf_ok = (mask & 8u) && eligible(3, n) && n >= 100 && f.K >= 2 && f.K <= 255;
    if (hf_ok) {
        if (fabsf(f.H - 7.0f) < 0.02f) { // near the threshold: fp64, in the reference's summation order (:566-574)
            chunk_first_order(c, hs.firstpos, hs.order);
            hf_ok = chunk_entropy_ordered(c, f.K, hs.order) < 7.0;
        } else hf_ok = f.H < 7.0f;
    }
    // Delta (id 4) always produces n bytes, so (n + overhead)/n > 1 never wins (:574-577).

    // The reference tries the methods in id order and keeps the first strictly smaller payload (:575).
    // The order of evaluation below differs, the outcome does not: a method's trial is skipped or cut
    // short only when its payload provably cannot end up as the winner.
    int best_type = 255, best_len = 0x7fffffff;
    if (rle_ok) {
        int len = 2 * f.rle_pairs;
        if (len + ovh < n) { best_type = 1; best_len = len; }
    }
    // Huffman lower bound from the entropy: bits >= max(n, n*H)
    int hf_lb = 0x7fffffff;
    // (f.H is within 3e-4 of the entropy: 0.002 and the 0.01 keep the bound below n * H whatever fp32 rounds to)
    if (hf_ok) hf_lb = 1 + 5 * f.K + 4 + (int)ceilf(fmaxf((float)n, (float)n * (f.H - 0.002f) - 0.01f) * 0.125f);
    // Huffman first when the Dictionary method looks weak (many distinct trigrams): its size then cuts
    // the match search short (lz2_match_all).  Its code table survives the search in c.hcode / c.hlen.
    // Chunks above LZ2_NMAX bytes always go Huffman first: their Dictionary trial is the window-aware
    // bucket search, twenty times the cost of a prefix trial with the names search (below).
    const bool big = n > LZ2_NMAX;
    const bool hf_first = hf_ok && lz_ok && (big || 100 * f.distinct3 >= 34 * min(1000, n));
    const bool lz_weak = !big && 100 * f.distinct3 >= 43 * min(1000, n);
    int hf_len = 0x7fffffff, hf_bits = 0; // hf_len: built, and a candidate against RLE
    if (hf_first && hf_lb < best_len && hf_lb + ovh < n) {
        hf_bits = chunk_huff_build(c, hs, f.K);
        const int len = 1 + 5 * f.K + 4 + ((hf_bits + 7) >> 3);
        if (len < best_len && len + ovh < n) hf_len = len;
    }
    if (lz_ok) {
        // a Dictionary payload can never be shorter than lz_lower_bound(n): skip the trial when it cannot
        // win.  It ends up as the winner only if it is < the RLE payload and <= the Huffman payload.
        const int cutoff = hf_len == 0x7fffffff ? best_len : min(best_len, hf_len + 1);
        int lb = lz_lower_bound(n);
        if (lb < cutoff && lb + ovh < n) {
            PHASE(10);
            // Prefix trial (a gamble on speed, never on the outcome): with a Huffman payload in hand and a
            // fair share of distinct trigrams the Dictionary method usually loses clearly.  The exact parse
            // of the first 5/8 of the chunk (positions below n1 see the same 32-byte look-ahead as in the
            // whole chunk, so their tokens are the real ones) then already costs more than the cutoff:
            // tokens starting below n1 cost at least lenp - 62 (the prefix parse may end with <= 31 bytes
            // of tokens starting at or after n1, <= 2 payload bytes per byte), and the n - np bytes behind
            // the prefix need at least 4 bytes per 32.  When the bound does not reach the cutoff the whole
            // chunk is parsed as usual; `trial` counts attempts / aborts per launch and switches the gamble
            // off where it keeps failing.
            bool aborted = false;
            if (hf_first && !lz_weak && hf_len != 0x7fffffff && n >= 2048 && (trial || big) && c.T) {
                // one thread reads the counters (other CTAs update them all the time): the whole block must
                // take the same branch, the calls below are collective.  Big chunks always try: the prefix
                // (the first LZ2_NMAX - 1 bytes, all inside the window of every position) costs a twentieth
                // of their bucket search.
                if (threadIdx.x == 0) {
                    int go = 1;
                    if (!big) {
                        const unsigned int att = ((volatile unsigned int *)trial)[0], hit = ((volatile unsigned int *)trial)[1];
                        go = (att < 64u || 4u * hit >= 3u * att) ? 1 : 0;
                    }
                    c.red[25] = go;
                }
                __syncthreads();
                const bool go = c.red[25] != 0;
                if (go) {
                    const int np = big ? LZ2_NMAX - 1 : ((5 * n / 8) & ~31) + 31;
                    c.n = np;
                    const int lenp = chunk_lz_encode(c);
                    c.n = n;
                    const int r = n - np;
                    aborted = lenp - 62 + 4 * (r >> 5) + min(4, 2 * (r & 31)) >= cutoff;
                    if (threadIdx.x == 0 && !big) {
                        atomicAdd(&trial[0], 1u);
                        if (aborted) atomicAdd(&trial[1], 1u);
                    }
                }
            }
            if (!aborted) {
                int len = chunk_lz_encode(c, (hf_first && lz_weak) ? cutoff : LZ_ABORTED);
                PHASE(11);
                if (len < best_len && len + ovh < n) { best_type = 2; best_len = len; }
            }
        }
    }
    if (hf_first) {
        if (hf_len < best_len) {
            best_type = 3; best_len = hf_len;
            chunk_huff_emit(c, hs, f.K, hf_bits);
        }
    } else if (hf_ok) {
        if (hf_lb < best_len && hf_lb + ovh < n) {
            int bits = chunk_huff_build(c, hs, f.K);
            int len = 1 + 5 * f.K + 4 + ((bits + 7) >> 3);
            if (len < best_len && len + ovh < n) {
                best_type = 3; best_len = len;
                chunk_huff_emit(c, hs, f.K, bits);
            }
        }
    }
    PHASE(12);
    if (best_type == 1) chunk_rle_encode(c);
    PHASE(13);
    SelectOut o;
    o.type = best_type;
    o.len = best_type == 255 ? n : best_len;
    return o;
}

#ifndef KSEL_MINB
#define KSEL_MINB 2
#endif
__global__ void __launch_bounds__(AMBC_BLOCK, KSEL_MINB)
k_select(const uint8_t *__restrict__ in, uint64_t total, uint32_t N, uint32_t mask, uint32_t ovh,
         uint8_t *__restrict__ slots, uint64_t slot_stride, uint8_t *__restrict__ type,
         uint32_t *__restrict__ comp, unsigned long long *first_raw, uint64_t chunk_begin, uint64_t n_chunks,
         unsigned int *trial)
{
    extern __shared__ uint4 smem4[];
    ChunkCtx c;
    if (N <= LZ2_NMAX) chunkctx_carve_fast(c, (uint8_t *)smem4, (int)N);
    else chunkctx_carve(c, (uint8_t *)smem4, (int)N, (int)N);
    for (uint64_t i = chunk_begin + blockIdx.x; i < n_chunks; i += gridDim.x) {
        uint64_t off = i * (uint64_t)N;
        int n = (int)min((uint64_t)N, total - off);
        chunk_load(c, in + off, n);
        SelectOut o = select_chunk(c, mask, (int)ovh, trial);
        __syncthreads();
        if (o.type != 255) {
            uint8_t *dst = slots + i * slot_stride; // 16-byte aligned
            int nv = (o.len + 15) >> 4;
            for (int k = threadIdx.x; k < nv; k += AMBC_BLOCK) ((uint4 *)dst)[k] = ((const uint4 *)c.pay)[k];
        }
        if (threadIdx.x == 0) {
            type[i] = (uint8_t)o.type;
            comp[i] = (uint32_t)o.len;
            if (o.type == 255) atomicMin(first_raw, (unsigned long long)i);
        }
        __syncthreads();
    }
}

// round-2 kernel (select_fast.cuh): 128 threads and ~38 KB of shared memory per chunk, five CTAs per SM
template <int NMAX>
__global__ void __launch_bounds__(SF_T, NMAX <= 4096 ? 6 : 3)
k_select_fast(const uint8_t *__restrict__ in, uint64_t total, uint32_t N, uint32_t mask, uint32_t ovh,
              uint8_t *__restrict__ slots, uint64_t slot_stride, uint8_t *__restrict__ type,
              uint32_t *__restrict__ comp, unsigned long long *first_raw, uint64_t chunk_begin, uint64_t n_chunks,
              uint32_t strict)
{
    extern __shared__ uint4 smem4[];
    SfCtx<NMAX> c;
    sf_carve<NMAX>(c, (uint8_t *)smem4);
    for (uint64_t i = chunk_begin + blockIdx.x; i < n_chunks; i += gridDim.x) {
        const uint64_t off = i * (uint64_t)N;
        const int n = (int)min((uint64_t)N, total - off);
        // strict mode: a chunk behind a chunk without a winner lies inside the one raw package that ends the file
        // (adaptive_compressor.py:586-590), whatever its own trial would say.  first_raw only ever decreases, so a
        // value below i seen now is final enough: the scan and pack kernels ignore the map from first_raw on.
        // (one thread reads for the CTA: first_raw changes under the kernel's feet, and threads that read
        // different values would part ways in front of the barriers of the trial)
        if (strict) {
            if (threadIdx.x == 0) c.red[30] = *(volatile unsigned long long *)first_raw < i;
            __syncthreads();
        }
        if (strict && c.red[30]) {
            if (threadIdx.x == 0) { type[i] = 255; comp[i] = (uint32_t)n; }
            continue;
        }
        sf_load<NMAX>(c, in + off, n);
        const SfOut o = sf_select<NMAX>(c, mask, (int)ovh);
        __syncthreads();
        if (o.type != 255) {
            uint8_t *dst = slots + i * slot_stride; // 16-byte aligned
            const int nv = (o.len + 15) >> 4;
            for (int k = threadIdx.x; k < nv; k += SF_T) ((uint4 *)dst)[k] = ((const uint4 *)c.pay)[k];
        }
        if (threadIdx.x == 0) {
            type[i] = (uint8_t)o.type;
            comp[i] = (uint32_t)o.len;
            if (o.type == 255) atomicMin(first_raw, (unsigned long long)i);
        }
        __syncthreads();
    }
}
// dev knob: AMBC_SELECT=old keeps the round-1 kernel (A/B timing)
static bool use_fast_select(uint32_t chunk)
{
    static int mode = -1;
    if (mode < 0) { const char *e = getenv("AMBC_SELECT"); mode = (e && !strcmp(e, "old")) ? 0 : 1; }
    return mode == 1 && chunk <= AMBC_NMAX;
}

// span mode (multi-candidate path, see k_select_span below) with the round-2 chunk code
struct ChunkSpan;
template <int NMAX>
__global__ void __launch_bounds__(SF_T, NMAX <= 4096 ? 6 : 3)
k_select_span_fast(const uint8_t *__restrict__ in, uint64_t total, uint32_t N, uint32_t mask, uint32_t ovh, uint32_t stride,
                   const unsigned long long *__restrict__ list, uint8_t *__restrict__ slots, uint8_t *__restrict__ type,
                   uint32_t *__restrict__ comp, uint64_t n_items)
{
    // list == nullptr: item i = [i * stride, i * stride + min(N, total - i * stride)), no payload kept;
    // else item i = (pos, size) = (list[2 i], low half of list[2 i + 1]) and its payload goes to slots + pos
    extern __shared__ uint4 smem4[];
    SfCtx<NMAX> c;
    sf_carve<NMAX>(c, (uint8_t *)smem4);
    for (uint64_t i = blockIdx.x; i < n_items; i += gridDim.x) {
        uint64_t off;
        int n;
        if (list) { off = list[2 * i]; n = (int)(uint32_t)list[2 * i + 1]; }
        else { off = i * (uint64_t)stride; n = (int)min((uint64_t)N, total - off); }
        sf_load<NMAX>(c, in + off, n);
        const SfOut o = sf_select<NMAX>(c, mask, (int)ovh);
        __syncthreads();
        if (slots && o.type != 255) {
            uint8_t *dst = slots + off;
            if ((((uintptr_t)dst) & 15) == 0) {
                const int nv = (o.len + 15) >> 4;
                for (int k = threadIdx.x; k < nv; k += SF_T) ((uint4 *)dst)[k] = ((const uint4 *)c.pay)[k];
            } else {
                for (int k = threadIdx.x; k < o.len; k += SF_T) dst[k] = c.pay[k];
            }
        }
        if (threadIdx.x == 0) {
            type[i] = (uint8_t)o.type;
            comp[i] = (uint32_t)o.len;
        }
        __syncthreads();
    }
}

// ---- package-size scan ---------------------------------------------------------------------
// In strict mode chunks >= first_raw contribute nothing (they live inside the single raw
// package); in per-chunk-raw mode every chunk is its own package.
#define SCAN_TILE 2048
struct ScanState {
    unsigned long long first_raw;  // min chunk index without a winner, ~0 if none
    unsigned long long body_len;
    unsigned long long n_packages;
    unsigned long long payload_bytes;
    unsigned long long usage[5];
    unsigned long long carry;      // body bytes before the chunks not yet scanned (piece-wise runs)
    unsigned int trial[2];         // prefix trials of k_select in this run: attempts, aborts (speed heuristic only)
};

__device__ __forceinline__ uint64_t pkg_size(uint64_t i, const uint8_t *type, const uint32_t *comp, uint32_t N,
                                             uint64_t total, uint32_t ovh, uint64_t first_raw, bool pcr)
{
    if (!pcr && i >= first_raw) return 0;
    uint64_t len = comp[i]; // raw chunks carry comp = n
    (void)N; (void)total; (void)type;
    return ovh + len;
}

__global__ void __launch_bounds__(256)
k_sizes(const uint8_t *__restrict__ type, const uint32_t *__restrict__ comp, uint64_t chunk_begin, uint64_t n_chunks,
        uint32_t N, uint64_t total, uint32_t ovh, uint32_t flags, const ScanState *st, unsigned long long *tile_sum)
{
    // chunks [chunk_begin, n_chunks); chunk_begin is a multiple of SCAN_TILE; tile_sum is indexed globally
    __shared__ unsigned long long red[8];
    const bool pcr = flags & 1u;
    const uint64_t fr = st->first_raw;
    uint64_t base = chunk_begin + (uint64_t)blockIdx.x * SCAN_TILE;
    unsigned long long s = 0;
    for (int k = threadIdx.x; k < SCAN_TILE; k += 256) {
        uint64_t i = base + k;
        if (i < n_chunks) s += pkg_size(i, type, comp, N, total, ovh, fr, pcr);
    }
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(FULL_MASK, s, d);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int i = 0; i < 8; i++) t += red[i];
        tile_sum[base / SCAN_TILE] = t;
    }
}

// single CTA: exclusive scan of the tile sums in place, body length, END package, stats
__global__ void __launch_bounds__(1024)
k_scan_tiles(unsigned long long *tile_sum, uint64_t tile_begin, uint64_t n_tiles, int is_last,
             const uint8_t *__restrict__ type, const uint32_t *__restrict__ comp, uint64_t n_chunks, uint32_t N,
             uint64_t total, uint32_t ovh, uint32_t flags, ScanState *st, uint8_t *out, uint64_t out_cap,
             uint32_t marker_word, uint32_t mb)
{
    // tiles [tile_begin, n_tiles), continuing from st->carry; the last call closes the body
    __shared__ unsigned long long wsum[32];
    __shared__ unsigned long long carry_s;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) carry_s = st->carry;
    __syncthreads();
    for (uint64_t base = tile_begin; base < n_tiles; base += 1024) {
        uint64_t i = base + tid;
        unsigned long long v = i < n_tiles ? tile_sum[i] : 0, inc = v;
        for (int d = 1; d < 32; d <<= 1) {
            unsigned long long t = __shfl_up_sync(FULL_MASK, inc, d);
            if (lane >= d) inc += t;
        }
        if (lane == 31) wsum[w] = inc;
        __syncthreads();
        unsigned long long wbase = 0;
        for (int k = 0; k < w; k++) wbase += wsum[k];
        unsigned long long carry = carry_s;
        if (i < n_tiles) tile_sum[i] = carry + wbase + inc - v;
        __syncthreads();
        if (tid == 1023) carry_s = carry + wbase + inc;
        __syncthreads();
    }
    if (tid == 0) st->carry = carry_s;
    if (tid == 0 && is_last) {
        const bool pcr = flags & 1u;
        unsigned long long body = carry_s;
        unsigned long long fr = st->first_raw;
        unsigned long long npk = n_chunks;
        if (!pcr && fr != ~0ull) {
            // one raw package for the rest of the file (adaptive_compressor.py:586-590)
            body += ovh + (total - fr * (unsigned long long)N);
            npk = fr + 1;
        }
        st->n_packages = npk;
        if (!(flags & 2u)) { // (bit 1: the caller appends a raw tail package and END itself -- span mode)
            body += mb + 12; // END package (:595-607)
            if (body <= out_cap) {
                uint8_t *e = out + body - (mb + 12);
                for (uint32_t k = 0; k < mb; k++) e[k] = (uint8_t)(marker_word >> (8 * k));
                for (uint32_t k = 0; k < 12; k++) e[mb + k] = 0;
            }
        }
        st->body_len = body;
    }
}

__global__ void __launch_bounds__(256)
k_offsets(const uint8_t *__restrict__ type, const uint32_t *__restrict__ comp, uint64_t chunk_begin, uint64_t n_chunks,
          uint32_t N, uint64_t total, uint32_t ovh, uint32_t flags, ScanState *st,
          const unsigned long long *tile_base, unsigned long long *offs)
{
    // one tile per CTA, 8 chunks per thread
    __shared__ unsigned long long wsum[8];
    __shared__ unsigned long long ustat[6];
    const bool pcr = flags & 1u;
    const uint64_t fr = st->first_raw;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid < 6) ustat[tid] = 0;
    uint64_t base = chunk_begin + (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)tid * 8;
    unsigned long long v[8], s = 0;
    unsigned long long usage[5] = {0, 0, 0, 0, 0}, pbytes = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        uint64_t i = base + k;
        v[k] = i < n_chunks ? pkg_size(i, type, comp, N, total, ovh, fr, pcr) : 0;
        s += v[k];
        if (i < n_chunks && (pcr || i < fr)) {
            int t = type[i];
            if (t >= 1 && t <= 4) { usage[t]++; pbytes += comp[i]; }
            else usage[0]++;
        }
    }
    unsigned long long inc = s;
    for (int d = 1; d < 32; d <<= 1) {
        unsigned long long t = __shfl_up_sync(FULL_MASK, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    unsigned long long wbase = 0;
    for (int k = 0; k < w; k++) wbase += wsum[k];
    unsigned long long run = tile_base[chunk_begin / SCAN_TILE + blockIdx.x] + wbase + inc - s;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        uint64_t i = base + k;
        if (i < n_chunks) offs[i] = run;
        run += v[k];
    }
    // statistics (adaptive_compressor.py:471-480)
#pragma unroll
    for (int k = 0; k < 5; k++) {
        unsigned long long u = usage[k];
        for (int d = 16; d > 0; d >>= 1) u += __shfl_xor_sync(FULL_MASK, u, d);
        if (lane == 0 && u) atomicAdd(&ustat[k], u);
    }
    for (int d = 16; d > 0; d >>= 1) pbytes += __shfl_xor_sync(FULL_MASK, pbytes, d);
    if (lane == 0 && pbytes) atomicAdd(&ustat[5], pbytes);
    __syncthreads();
    if (tid < 5 && ustat[tid]) atomicAdd(&st->usage[tid], ustat[tid]);
    if (tid == 5 && ustat[5]) atomicAdd(&st->payload_bytes, ustat[5]);
}

// ---- pack ----------------------------------------------------------------------------------
// package header (adaptive_compressor.py:609-621): marker, type, k=0, used u32, orig u32, comp u32
__device__ __forceinline__ void write_pkg_header(uint8_t *h, uint32_t marker_word, uint32_t mb, int type,
                                                 uint32_t orig, uint32_t comp)
{
    for (uint32_t k = 0; k < mb; k++) h[k] = (uint8_t)(marker_word >> (8 * k));
    h[mb] = (uint8_t)type;
    h[mb + 1] = 0;
    store_u32le(h + mb + 2, orig);
    store_u32le(h + mb + 6, orig);
    store_u32le(h + mb + 10, comp);
}

#define PACK_BLOCK 128
__global__ void __launch_bounds__(PACK_BLOCK)
k_pack(const uint8_t *__restrict__ in, uint64_t total, uint32_t N, const uint8_t *__restrict__ slots,
       uint64_t slot_stride, const uint8_t *__restrict__ type, const uint32_t *__restrict__ comp,
       const unsigned long long *__restrict__ offs, const ScanState *st, uint32_t flags, uint32_t marker_word,
       uint32_t mb, uint8_t *__restrict__ out, uint64_t out_cap, uint64_t chunk_begin, uint64_t n_chunks, int check_cap)
{
    extern __shared__ uint4 smem4[];
    uint8_t *buf = (uint8_t *)smem4; // [32 header area][payload]
    const bool pcr = flags & 1u;
    const uint64_t fr = st->first_raw;
    const uint32_t ovh = mb + 14;
    if (check_cap && st->body_len > out_cap) return;
    for (uint64_t i = chunk_begin + blockIdx.x; i < n_chunks; i += gridDim.x) {
        uint64_t coff = i * (uint64_t)N;
        int n = (int)min((uint64_t)N, total - coff);
        if (!pcr && i >= fr) {
            // inside the single raw package: header once, then plain bytes
            uint64_t pbase = offs[fr]; // == offset of the raw package (sizes beyond fr are 0)
            uint64_t rawlen = total - fr * (uint64_t)N;
            copy_g2s<PACK_BLOCK>(buf + 32, in + coff, n);
            uint8_t *src = buf + 32;
            int len = n;
            if (i == fr) {
                if (threadIdx.x == 0) write_pkg_header(buf + 32 - ovh, marker_word, mb, 255, (uint32_t)rawlen, (uint32_t)rawlen);
                src = buf + 32 - ovh;
                len = n + ovh;
            }
            __syncthreads();
            uint64_t dst = (i == fr) ? pbase : pbase + ovh + (coff - fr * (uint64_t)N);
            copy_s2g<PACK_BLOCK>(out + dst, src, len);
        } else {
            int t = type[i];
            int len = (int)comp[i];
            if (t == 255) copy_g2s<PACK_BLOCK>(buf + 32, in + coff, n);
            else copy_g2s<PACK_BLOCK>(buf + 32, slots + i * slot_stride, len);
            if (threadIdx.x == 0) write_pkg_header(buf + 32 - ovh, marker_word, mb, t, (uint32_t)n, (uint32_t)len);
            __syncthreads();
            copy_s2g<PACK_BLOCK>(out + offs[i], buf + 32 - ovh, len + ovh);
        }
        __syncthreads();
    }
}

// ---- span mode: chunks given by (position, size) instead of the fixed grid ---------------------
// Used by the multi-candidate ("dynamic chunk size") path, adaptive_compressor.py:537-590: first as a
// trial at every stride-aligned position for one candidate size (sizes only), then on the chosen
// chunk chain (payloads).  Separate kernels so that the fixed-grid hot path stays untouched.
struct ChunkSpan { unsigned long long pos; uint32_t size; uint32_t pad; };

template <int MINB> // 2: sizes up to LZ2_NMAX (two CTAs per SM like k_select), 1: larger chunks
__global__ void __launch_bounds__(AMBC_BLOCK, MINB)
k_select_span(const uint8_t *__restrict__ in, uint64_t total, uint32_t N, uint32_t mask, uint32_t ovh,
              uint32_t stride, const ChunkSpan *__restrict__ list, uint8_t *__restrict__ slots,
              uint8_t *__restrict__ type, uint32_t *__restrict__ comp, uint64_t n_items)
{
    // list == nullptr: item i = [i * stride, i * stride + min(N, total - i * stride)), no payload kept;
    // else item i = list[i] and its payload goes to slots + list[i].pos (payloads are shorter than
    // their chunks, so the chunk's own offset is a collision-free slot)
    extern __shared__ uint4 smem4[];
    ChunkCtx c;
    if (N <= LZ2_NMAX) chunkctx_carve_fast(c, (uint8_t *)smem4, (int)N);
    else chunkctx_carve(c, (uint8_t *)smem4, (int)N, (int)N);
    for (uint64_t i = blockIdx.x; i < n_items; i += gridDim.x) {
        uint64_t off;
        int n;
        if (list) { off = list[i].pos; n = (int)list[i].size; }
        else { off = i * (uint64_t)stride; n = (int)min((uint64_t)N, total - off); }
        chunk_load(c, in + off, n);
        SelectOut o = select_chunk(c, mask, (int)ovh);
        __syncthreads();
        if (slots && o.type != 255) {
            uint8_t *dst = slots + off;
            if ((((uintptr_t)dst) & 15) == 0) {
                int nv = (o.len + 15) >> 4;
                for (int k = threadIdx.x; k < nv; k += AMBC_BLOCK) ((uint4 *)dst)[k] = ((const uint4 *)c.pay)[k];
            } else {
                for (int k = threadIdx.x; k < o.len; k += AMBC_BLOCK) dst[k] = c.pay[k];
            }
        }
        if (threadIdx.x == 0) {
            type[i] = (uint8_t)o.type;
            comp[i] = (uint32_t)o.len;
        }
        __syncthreads();
    }
}

static int launch_select_span(unsigned grid, size_t smem, cudaStream_t stream, const uint8_t *in, uint64_t total, uint32_t N,
                              uint32_t mask, uint32_t ovh, uint32_t stride, const ChunkSpan *list, uint8_t *slots,
                              uint8_t *type, uint32_t *comp, uint64_t n_items)
{
    static_assert(sizeof(ChunkSpan) == 16, "k_select_span_fast reads spans as two 64-bit words");
    if (use_fast_select(N) && N <= 4096) {
        CUDA_TRY(cudaFuncSetAttribute(k_select_span_fast<4096>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SfCfg<4096>::SMEM));
        k_select_span_fast<4096><<<grid, SF_T, SfCfg<4096>::SMEM, stream>>>(in, total, N, mask, ovh, stride, (const unsigned long long *)list,
                                                                            slots, type, comp, n_items);
    } else if (use_fast_select(N)) {
        CUDA_TRY(cudaFuncSetAttribute(k_select_span_fast<8192>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SfCfg<8192>::SMEM));
        k_select_span_fast<8192><<<grid, SF_T, SfCfg<8192>::SMEM, stream>>>(in, total, N, mask, ovh, stride, (const unsigned long long *)list,
                                                                            slots, type, comp, n_items);
    } else if (N <= LZ2_NMAX) {
        CUDA_TRY(cudaFuncSetAttribute(k_select_span<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_select_span<2><<<grid, AMBC_BLOCK, smem, stream>>>(in, total, N, mask, ovh, stride, list, slots, type, comp, n_items);
    } else {
        CUDA_TRY(cudaFuncSetAttribute(k_select_span<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_select_span<1><<<grid, AMBC_BLOCK, smem, stream>>>(in, total, N, mask, ovh, stride, list, slots, type, comp, n_items);
    }
    ambc_count_launch();
    CUDA_TRY(cudaGetLastError());
    return AMBC_OK;
}

__global__ void __launch_bounds__(PACK_BLOCK)
k_pack_span(const uint8_t *__restrict__ in, const ChunkSpan *__restrict__ list, const uint8_t *__restrict__ slots,
            const uint8_t *__restrict__ type, const uint32_t *__restrict__ comp,
            const unsigned long long *__restrict__ offs, uint32_t marker_word, uint32_t mb,
            uint8_t *__restrict__ out, uint64_t n_items)
{
    extern __shared__ uint4 smem4[];
    uint8_t *buf = (uint8_t *)smem4; // [32 header area][payload]
    const uint32_t ovh = mb + 14;
    for (uint64_t i = blockIdx.x; i < n_items; i += gridDim.x) {
        const uint64_t coff = list[i].pos;
        const int n = (int)list[i].size;
        const int t = type[i];
        const int len = (int)comp[i];
        if (t == 255) {
            for (int k = threadIdx.x; k < n; k += PACK_BLOCK) buf[32 + k] = in[coff + k];
        } else {
            for (int k = threadIdx.x; k < len; k += PACK_BLOCK) buf[32 + k] = slots[coff + k];
        }
        if (threadIdx.x == 0) write_pkg_header(buf + 32 - ovh, marker_word, mb, t, (uint32_t)n, (uint32_t)len);
        __syncthreads();
        copy_s2g<PACK_BLOCK>(out + offs[i], buf + 32 - ovh, len + ovh);
        __syncthreads();
    }
}

// ---- host side -----------------------------------------------------------------------------
struct WorkLayout {
    uint64_t slots, type, comp, offs, tiles, state, total;
    uint64_t slot_stride, n_chunks, n_tiles;
};
static WorkLayout work_layout(uint64_t n, uint32_t chunk)
{
    WorkLayout L;
    L.n_chunks = chunk ? (n + chunk - 1) / chunk : 0;
    L.n_tiles = (L.n_chunks + SCAN_TILE - 1) / SCAN_TILE;
    L.slot_stride = chunk <= AMBC_NMAX ? ((uint64_t)chunk + 15) & ~15ull : 0;
    uint64_t o = 0;
    auto take = [&](uint64_t bytes) { uint64_t r = o; o += (bytes + 255) & ~255ull; return r; };
    L.slots = take(L.n_chunks * L.slot_stride + 16);
    L.type = take(L.n_chunks + 16);
    L.comp = take(L.n_chunks * 4 + 16);
    L.offs = take(L.n_chunks * 8 + 16);
    L.tiles = take(L.n_tiles * 8 + 16);
    L.state = take(sizeof(ScanState));
    L.total = o;
    return L;
}

extern "C" uint64_t ambc_compress_workspace_bytes(uint64_t n, uint32_t chunk)
{
    return work_layout(n, chunk).total;
}

extern "C" uint64_t ambc_compress_bound(uint64_t n, uint32_t chunk, uint32_t marker_bytes)
{
    uint64_t chunks = chunk ? (n + chunk - 1) / chunk : 0;
    return n + (chunks + 1) * (marker_bytes + 14) + marker_bytes + 12 + 64;
}

// piece_ready[k] (optional): event after which chunks [piece_start[k], piece_start[k + 1]) of the input
// are resident (ambc_compress_host uploads piece-wise so that H2D overlaps k_select); piece_start has
// n_pieces + 1 ascending entries, multiples of SCAN_TILE, the last one >= the chunk count
int ambc_compress_dev_impl(const void *in_dev, uint64_t n, uint32_t chunk, uint32_t method_mask, uint32_t flags,
                           const uint8_t *marker, uint32_t marker_bytes, void *out_dev, uint64_t out_cap,
                           void *work_dev, uint64_t work_bytes, ambc_compress_result *res, cudaStream_t stream,
                           const cudaEvent_t *piece_ready, const uint64_t *piece_start, uint32_t n_pieces,
                           const AmbcPieceOut *po);
extern "C" uint64_t ambc_scan_state_bytes(void) { return sizeof(ScanState); }

extern "C" int ambc_compress_dev(const void *in_dev, uint64_t n, uint32_t chunk, uint32_t method_mask,
                                 uint32_t flags, const uint8_t *marker, uint32_t marker_bytes, void *out_dev,
                                 uint64_t out_cap, void *work_dev, uint64_t work_bytes, ambc_compress_result *res,
                                 void *stream_)
{
    return ambc_compress_dev_impl(in_dev, n, chunk, method_mask, flags, marker, marker_bytes, out_dev, out_cap, work_dev,
                                  work_bytes, res, (cudaStream_t)stream_, nullptr, nullptr, 0, nullptr);
}

int ambc_compress_dev_impl(const void *in_dev, uint64_t n, uint32_t chunk, uint32_t method_mask, uint32_t flags,
                           const uint8_t *marker, uint32_t marker_bytes, void *out_dev, uint64_t out_cap,
                           void *work_dev, uint64_t work_bytes, ambc_compress_result *res, cudaStream_t stream,
                           const cudaEvent_t *piece_ready, const uint64_t *piece_start, uint32_t n_pieces,
                           const AmbcPieceOut *po)
{
    if (!res || !marker || marker_bytes < 1 || marker_bytes > 4 || chunk == 0)
        return ambc_fail(AMBC_E_ARG, "ambc_compress_dev: bad argument");
    if ((n && (!in_dev || !work_dev)) || !out_dev) return ambc_fail(AMBC_E_ARG, "ambc_compress_dev: null buffer");
    WorkLayout L = work_layout(n, chunk);
    if (work_bytes < L.total) return ambc_fail(AMBC_E_CAPACITY, "ambc_compress_dev: workspace too small");
    if ((uint64_t)chunk > 0xFFFFFFFFull - 32) return ambc_fail(AMBC_E_ARG, "chunk too large");
    uint32_t ovh = marker_bytes + 14;
    uint32_t marker_word = 0;
    for (uint32_t k = 0; k < marker_bytes; k++) marker_word |= (uint32_t)marker[k] << (8 * k);
    memset(res, 0, sizeof(*res));
    res->n_chunks = L.n_chunks;
    res->map_type_off = L.type;
    res->map_comp_off = L.comp;

    uint8_t *W = (uint8_t *)work_dev;
    ScanState *st = (ScanState *)(W + L.state);
    ScanState h_st;
    memset(&h_st, 0, sizeof h_st);
    h_st.first_raw = ~0ull;

    if (n == 0) { // only the END package
        uint8_t endp[16];
        memcpy(endp, marker, marker_bytes);
        memset(endp + marker_bytes, 0, 12);
        if (out_cap < marker_bytes + 12) return ambc_fail(AMBC_E_CAPACITY, "out too small");
        CUDA_TRY(cudaMemcpyAsync(out_dev, endp, marker_bytes + 12, cudaMemcpyHostToDevice, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
        if (po && po->out_host) {
            if (po->out_cap < marker_bytes + 12) return ambc_fail(AMBC_E_CAPACITY, "ambc_compress_host: out_cap too small");
            memcpy(po->out_host, endp, marker_bytes + 12);
        }
        res->body_len = marker_bytes + 12;
        res->first_raw = -1;
        return AMBC_OK;
    }
    // chunks of more than 4 GiB cannot be framed (u32 length fields, :617-619)
    bool native = chunk <= AMBC_NMAX && (method_mask & AMBC_NATIVE_MASK);
    if (!native) h_st.first_raw = 0; // no native method is eligible: everything is raw
    CUDA_TRY(cudaMemcpyAsync(st, &h_st, sizeof h_st, cudaMemcpyHostToDevice, stream));

    uint8_t *type = W + L.type;
    uint32_t *comp = (uint32_t *)(W + L.comp);
    unsigned long long *offs = (unsigned long long *)(W + L.offs);
    unsigned long long *tiles = (unsigned long long *)(W + L.tiles);
    unsigned grid_chunks = (unsigned)min<uint64_t>(L.n_chunks, 0x7fffffffull);

    if (native) {
        size_t smem = chunk <= LZ2_NMAX ? chunkctx_fast_smem_bytes((int)chunk) : chunkctx_smem_bytes((int)chunk, (int)chunk);
        size_t psmem = 32 + (((size_t)chunk + 15) & ~(size_t)15) + 32;
        CUDA_TRY(cudaFuncSetAttribute(k_select, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_TRY(cudaFuncSetAttribute(k_pack, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));
        CUDA_TRY(cudaFuncSetAttribute(k_select_fast<4096>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SfCfg<4096>::SMEM));
        CUDA_TRY(cudaFuncSetAttribute(k_select_fast<8192>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SfCfg<8192>::SMEM));
        // piece-wise run (host-buffer path): select / scan / pack of piece k are queued behind the
        // upload of piece k; the finished body bytes of a piece go home while later pieces compute
        bool pieces = piece_ready && n_pieces > 1 && piece_start && piece_start[0] == 0 && piece_start[n_pieces] >= L.n_chunks;
        for (uint32_t k = 0; pieces && k < n_pieces; k++)
            pieces = piece_start[k] % SCAN_TILE == 0 && piece_start[k] < piece_start[k + 1];
        if (!pieces && piece_ready)
            for (uint32_t k = 0; k < n_pieces; k++) CUDA_TRY(cudaStreamWaitEvent(stream, piece_ready[k], 0));
        const uint32_t np = pieces ? n_pieces : 1;
        const bool stream_out = pieces && po && po->out_host && po->states && po->done;
        cudaStream_t s2 = (stream_out && po->aux && po->sel) ? po->aux : stream;
        for (uint32_t k = 0; k < np; k++) {
            const uint64_t c0 = pieces ? piece_start[k] : 0;
            const uint64_t c1 = pieces ? min<uint64_t>(L.n_chunks, piece_start[k + 1]) : L.n_chunks;
            if (c0 >= c1) break;
            const bool last = c1 == L.n_chunks;
            const unsigned gch = (unsigned)min<uint64_t>(c1 - c0, 0x7fffffffull);
            const unsigned gt = (unsigned)((c1 - c0 + SCAN_TILE - 1) / SCAN_TILE);
            if (pieces) CUDA_TRY(cudaStreamWaitEvent(stream, piece_ready[k], 0));
            if (k == 0) ambc_timing_mark(0, stream);
            if (use_fast_select(chunk) && chunk <= 4096)
                k_select_fast<4096><<<gch, SF_T, SfCfg<4096>::SMEM, stream>>>((const uint8_t *)in_dev, n, chunk, method_mask, ovh,
                                                                             W + L.slots, L.slot_stride, type, comp,
                                                                             &st->first_raw, c0, c1, (flags & AMBC_F_PER_CHUNK_RAW) ? 0u : 1u);
            else if (use_fast_select(chunk))
                k_select_fast<8192><<<gch, SF_T, SfCfg<8192>::SMEM, stream>>>((const uint8_t *)in_dev, n, chunk, method_mask, ovh,
                                                                             W + L.slots, L.slot_stride, type, comp,
                                                                             &st->first_raw, c0, c1, (flags & AMBC_F_PER_CHUNK_RAW) ? 0u : 1u);
            else
                k_select<<<gch, AMBC_BLOCK, smem, stream>>>((const uint8_t *)in_dev, n, chunk, method_mask, ovh, W + L.slots,
                                                            L.slot_stride, type, comp, &st->first_raw, c0, c1, st->trial);
            ambc_count_launch();
            if (last) ambc_timing_mark(1, stream);
            // scan + pack of piece k run on their own stream so that k_select of piece k + 1 follows at once
            // (k_select only lowers st->first_raw to chunks of its own piece, which the scans of earlier
            // pieces never look at)
            if (s2 != stream) {
                CUDA_TRY(cudaEventRecord(po->sel[k], stream));
                CUDA_TRY(cudaStreamWaitEvent(s2, po->sel[k], 0));
            }
            k_sizes<<<gt, 256, 0, s2>>>(type, comp, c0, c1, chunk, n, ovh, flags, st, tiles);
            ambc_count_launch();
            k_scan_tiles<<<1, 1024, 0, s2>>>(tiles, c0 / SCAN_TILE, (c1 + SCAN_TILE - 1) / SCAN_TILE, last ? 1 : 0, type,
                                             comp, L.n_chunks, chunk, n, ovh, flags, st, (uint8_t *)out_dev, out_cap,
                                             marker_word, marker_bytes);
            ambc_count_launch();
            k_offsets<<<gt, 256, 0, s2>>>(type, comp, c0, c1, chunk, n, ovh, flags, st, tiles, offs);
            ambc_count_launch();
            if (last) ambc_timing_mark(2, s2);
            k_pack<<<gch, PACK_BLOCK, psmem, s2>>>((const uint8_t *)in_dev, n, chunk, W + L.slots, L.slot_stride, type,
                                                   comp, offs, st, flags, marker_word, marker_bytes, (uint8_t *)out_dev,
                                                   out_cap, c0, c1, pieces ? 0 : 1);
            ambc_count_launch();
            CUDA_TRY(cudaGetLastError());
            if (stream_out) {
                CUDA_TRY(cudaMemcpyAsync((uint8_t *)po->states + (size_t)k * sizeof(ScanState), st, sizeof(ScanState),
                                         cudaMemcpyDeviceToHost, s2));
                CUDA_TRY(cudaEventRecord(po->done[k], s2));
                if (last && s2 != stream) CUDA_TRY(cudaStreamWaitEvent(stream, po->done[k], 0));
            }
        }
        ambc_timing_mark(3, stream);
        ambc_timing().pending_c = ambc_timing().on;
        uint64_t sent = 0;
        if (stream_out) {
            // bytes [sent, carry_k) are final after piece k as long as no chunk lacked a winner
            for (uint32_t k = 0; k + 1 < np; k++) {
                CUDA_TRY(cudaEventSynchronize(po->done[k]));
                const ScanState *hs = (const ScanState *)((const uint8_t *)po->states + (size_t)k * sizeof(ScanState));
                if (hs->first_raw != ~0ull || hs->carry > po->out_cap) break;
                if (hs->carry > sent) {
                    CUDA_TRY(cudaMemcpyAsync((uint8_t *)po->out_host + sent, (const uint8_t *)out_dev + sent, hs->carry - sent,
                                             cudaMemcpyDeviceToHost, po->d2h));
                    sent = hs->carry;
                }
            }
        }
        CUDA_TRY(cudaMemcpyAsync(&h_st, st, sizeof h_st, cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
        if (po && po->out_host) {
            if (h_st.body_len > po->out_cap || h_st.body_len > out_cap) {
                cudaStreamSynchronize(po->d2h);
                return ambc_fail(AMBC_E_CAPACITY, "ambc_compress_host: out_cap %llu < body %llu",
                                 (unsigned long long)po->out_cap, (unsigned long long)h_st.body_len);
            }
            CUDA_TRY(cudaMemcpyAsync((uint8_t *)po->out_host + sent, (const uint8_t *)out_dev + sent, h_st.body_len - sent,
                                     cudaMemcpyDeviceToHost, stream));
            CUDA_TRY(cudaStreamSynchronize(stream));
            CUDA_TRY(cudaStreamSynchronize(po->d2h));
        }
    } else {
        // one raw package: header + memcpy + END, no kernel needed beyond the copy engine
        if (piece_ready)
            for (uint32_t k = 0; k < n_pieces; k++) CUDA_TRY(cudaStreamWaitEvent(stream, piece_ready[k], 0));
        uint64_t body = (uint64_t)ovh + n + marker_bytes + 12;
        if (n > 0xFFFFFFFFull) return ambc_fail(AMBC_E_ARG, "raw package over 4 GiB cannot be framed (u32 fields)");
        if (body > out_cap) return ambc_fail(AMBC_E_CAPACITY, "out too small");
        uint8_t hdr[18], endp[16];
        memcpy(hdr, marker, marker_bytes);
        hdr[marker_bytes] = 255; hdr[marker_bytes + 1] = 0;
        uint32_t n32 = (uint32_t)n;
        for (int r = 0; r < 3; r++) memcpy(hdr + marker_bytes + 2 + 4 * r, &n32, 4);
        memcpy(endp, marker, marker_bytes);
        memset(endp + marker_bytes, 0, 12);
        uint8_t *o = (uint8_t *)out_dev;
        CUDA_TRY(cudaMemcpyAsync(o, hdr, ovh, cudaMemcpyHostToDevice, stream));
        CUDA_TRY(cudaMemcpyAsync(o + ovh, in_dev, n, cudaMemcpyDeviceToDevice, stream));
        CUDA_TRY(cudaMemcpyAsync(o + ovh + n, endp, marker_bytes + 12, cudaMemcpyHostToDevice, stream));
        CUDA_TRY(cudaMemsetAsync(type, 255, L.n_chunks, stream));
        if (po && po->out_host) {
            if (body > po->out_cap) return ambc_fail(AMBC_E_CAPACITY, "ambc_compress_host: out_cap too small");
            CUDA_TRY(cudaMemcpyAsync(po->out_host, out_dev, body, cudaMemcpyDeviceToHost, stream));
        }
        CUDA_TRY(cudaStreamSynchronize(stream));
        h_st.body_len = body;
        h_st.n_packages = 1;
        h_st.usage[0] = 1;
    }
    if (h_st.body_len > out_cap) return ambc_fail(AMBC_E_CAPACITY, "ambc_compress_dev: out_cap too small");
    // the rest-of-file raw package carries its length in u32 fields (adaptive_compressor.py:617-619)
    if (native && !(flags & AMBC_F_PER_CHUNK_RAW) && h_st.first_raw != ~0ull &&
        n - h_st.first_raw * (uint64_t)chunk > 0xFFFFFFFFull)
        return ambc_fail(AMBC_E_ARG, "raw package over 4 GiB cannot be framed (u32 fields)");
    res->body_len = h_st.body_len;
    res->first_raw = h_st.first_raw == ~0ull ? -1 : (int64_t)h_st.first_raw;
    res->n_packages = h_st.n_packages;
    res->payload_bytes = h_st.payload_bytes;
    for (int k = 0; k < 5; k++) res->usage[k] = h_st.usage[k];
    if (native && !(flags & AMBC_F_PER_CHUNK_RAW) && res->first_raw >= 0) res->usage[0] = 1;
    return AMBC_OK;
}

#ifdef AMBC_PHASE_TIMING
extern "C" int ambc_sf_phase_read(unsigned long long *out48, int reset)
{
    cudaMemcpyFromSymbol(out48, g_sf, sizeof(unsigned long long) * 48);
    if (reset) { unsigned long long z[48] = {0}; cudaMemcpyToSymbol(g_sf, z, sizeof z); }
    return 0;
}
extern "C" int ambc_phase_read(unsigned long long *out32, int reset)
{
    cudaMemcpyFromSymbol(out32, g_phase, sizeof(unsigned long long) * 32);
    if (reset) { unsigned long long z[32] = {0}; cudaMemcpyToSymbol(g_phase, z, sizeof z); }
    return 0;
}
#endif
// ---- multi-candidate ("dynamic chunk size") mode ------------------------------------------------
// adaptive_compressor.py:537-590 with several CHUNK_SIZE_CANDIDATES: at every position each candidate
// size (clamped to the remaining bytes) is tried, the smallest ratio (len + overhead) / size wins,
// larger candidates win ties; no winner -> the rest of the file is one raw package.
// Every reachable position is a multiple of g = gcd(candidates), so each distinct size min(cand, 8192)
// is tried at every multiple of g on the GPU (sizes only), the chain is resolved on the host (it is a
// serial walk over <= n / g table entries), and the chosen chunks are encoded and framed in span mode.
static uint64_t gcd64(uint64_t a, uint64_t b) { while (b) { uint64_t t = a % b; a = b; b = t; } return a; }

// ---- the chain of the multi-candidate mode on the GPU ---------------------------------------------------------
// The reference walks the file once: at every position it tries every candidate size and moves on by the winner's
// size (adaptive_compressor.py:363-394, 537-590).  Only the positions on that chain matter, about one in six of
// the multiples of g -- round 1 tried every size at EVERY multiple of g and walked the tables on the host.  Here
// the file is cut into segments of DYN_SEGU * g bytes; one CTA walks the chain of a segment from its first byte,
// evaluating the candidates at the positions it actually visits (speculative: the true chain may enter the segment
// elsewhere).  Chains that enter a segment at different positions meet after a few packages (the winner at a
// position is a function of the position), so the host stitches the true chain from the speculative ones and
// re-enters (a second launch, same kernel) only the segments whose true entry differs, each until it meets the
// segment's speculative chain.
#define DYN_SEGU 64
struct WalkRec { unsigned long long pos; uint32_t size; uint32_t type_len; };   // type << 24 | payload length
struct WalkSeg { unsigned long long exit; unsigned long long stop; uint32_t count; uint32_t merged; uint32_t pad[2]; };
struct WalkCands { uint32_t n; uint32_t c[32]; };
__global__ void __launch_bounds__(SF_T, 3)
k_dynamic_walk(const uint8_t *__restrict__ in, uint64_t total, WalkCands cands, uint32_t mask, uint32_t ovh, uint32_t pcr,
               uint32_t g, const unsigned long long *__restrict__ list, uint64_t n_items,
               const WalkRec *__restrict__ spec_recs, const WalkSeg *__restrict__ spec_segs,
               WalkRec *__restrict__ recs, WalkSeg *__restrict__ segs)
{
    // list == nullptr: item i = segment i entered at its first byte; else item i = (segment, entry position) and the
    // walk stops where it meets the segment's speculative chain (spec_recs / spec_segs)
    extern __shared__ uint4 smem4[];
    SfCtx<8192> c;
    sf_carve<8192>(c, (uint8_t *)smem4);
    __shared__ uint32_t visited[DYN_SEGU / 32];
    const uint64_t seglen = (uint64_t)DYN_SEGU * g;
    for (uint64_t it = blockIdx.x; it < n_items; it += gridDim.x) {
        const uint64_t sg = list ? list[2 * it] : it;
        const uint64_t s0 = sg * seglen, s1 = min(total, s0 + seglen);
        uint64_t pos = list ? list[2 * it + 1] : s0;
        if (list) { // positions the speculative chain of this segment visits
            if (threadIdx.x < DYN_SEGU / 32) visited[threadIdx.x] = 0;
            __syncthreads();
            const uint32_t cnt = spec_segs[sg].count;
            for (uint32_t k = threadIdx.x; k < cnt; k += SF_T) {
                const uint32_t u = (uint32_t)((spec_recs[sg * DYN_SEGU + k].pos - s0) / g);
                atomicOr(&visited[u >> 5], 1u << (u & 31));
            }
            __syncthreads();
        }
        uint32_t count = 0, merged = 0;
        unsigned long long stop = ~0ull;
        while (pos < s1) {
            if (list) {
                const uint32_t u = (uint32_t)((pos - s0) / g);
                if ((visited[u >> 5] >> (u & 31)) & 1u) { merged = 1; break; }
            }
            const uint64_t remain = total - pos;
            // (adaptive_compressor.py:548-584) every candidate, clamped to the rest of the file, in list order; the
            // smallest (len + overhead) / size wins as a double, the first of equals stays
            double best_ratio = 1.0;
            uint32_t best_c = 0, best_t = 255, best_len = 0, last_c = 0, last_t = 255, last_len = 0;
            for (uint32_t ci = 0; ci < cands.n; ci++) {
                const uint32_t cs = (uint64_t)cands.c[ci] < remain ? cands.c[ci] : (uint32_t)remain;
                if (cs > AMBC_NMAX) continue; // no native method is eligible (:114-127)
                uint32_t t, len;
                if (cs == last_c) { t = last_t; len = last_len; } // (clamped candidates repeat near the end of the file)
                else {
                    sf_load<8192>(c, in + pos, (int)cs);
                    const SfOut o = sf_select<8192>(c, mask, (int)ovh);
                    __syncthreads();
                    t = (uint32_t)o.type; len = (uint32_t)o.len;
                    last_c = cs; last_t = t; last_len = len;
                }
                if (t == 255) continue;
                const double ratio = (double)((uint64_t)len + ovh) / (double)cs;
                if (ratio < best_ratio) { best_ratio = ratio; best_c = cs; best_t = t; best_len = len; }
            }
            if (best_t == 255) {
                if (!pcr) { stop = pos; break; } // rest of the file raw (:586-590)
                best_c = (uint64_t)cands.c[cands.n - 1] < remain ? cands.c[cands.n - 1] : (uint32_t)remain; // labelled extension
                best_len = best_c;
            }
            if (threadIdx.x == 0) {
                WalkRec r;
                r.pos = pos; r.size = best_c; r.type_len = (best_t << 24) | best_len;
                recs[sg * DYN_SEGU + count] = r;
            }
            count++;
            pos += best_c;
        }
        if (threadIdx.x == 0) {
            WalkSeg w;
            w.exit = pos; w.stop = stop; w.count = count; w.merged = merged; w.pad[0] = w.pad[1] = 0;
            segs[list ? it : sg] = w;
        }
        __syncthreads();
    }
}

struct DynLayout {
    uint64_t slots, spans, type, comp, offs, tiles, state, trial_type, trial_len, walk_recs, walk_segs, walk_list, total;
    uint64_t n_pos, n_tiles, n_pass, g, n_seg;
    uint32_t pass_size[16];
    int cand_pass[32];
};
static int dyn_layout(uint64_t n, const uint32_t *cands, uint32_t n_cands, DynLayout &L)
{
    if (!cands || n_cands < 1 || n_cands > 32) return ambc_fail(AMBC_E_ARG, "need 1..32 candidate sizes");
    uint64_t g = 0;
    for (uint32_t i = 0; i < n_cands; i++) {
        if (cands[i] == 0) return ambc_fail(AMBC_E_ARG, "candidate size 0");
        if (i && cands[i] >= cands[i - 1]) return ambc_fail(AMBC_E_ARG, "candidate sizes must be strictly descending");
        g = gcd64(g, cands[i]);
    }
    if (g < 256 || g % 16) return ambc_fail(AMBC_E_ARG, "candidate sizes need a common divisor that is a multiple of 16 and >= 256");
    L.g = g;
    L.n_pos = (n + g - 1) / g;
    L.n_pass = 0;
    for (uint32_t i = 0; i < n_cands; i++) {
        uint32_t sz = cands[i] < AMBC_NMAX ? cands[i] : AMBC_NMAX;
        int found = -1;
        for (uint64_t j = 0; j < L.n_pass; j++) if (L.pass_size[j] == sz) found = (int)j;
        if (found < 0) { if (L.n_pass >= 16) return ambc_fail(AMBC_E_ARG, "too many distinct candidate sizes"); found = (int)L.n_pass; L.pass_size[L.n_pass++] = sz; }
        L.cand_pass[i] = found;
    }
    L.n_tiles = (L.n_pos + SCAN_TILE - 1) / SCAN_TILE;
    uint64_t o = 0;
    auto take = [&](uint64_t bytes) { uint64_t r = o; o += (bytes + 255) & ~255ull; return r; };
    L.slots = take(n + 64);
    L.spans = take(L.n_pos * sizeof(ChunkSpan) + 16);
    L.type = take(L.n_pos + 16);
    L.comp = take(L.n_pos * 4 + 16);
    L.offs = take(L.n_pos * 8 + 16);
    L.tiles = take(L.n_tiles * 8 + 16);
    L.state = take(sizeof(ScanState));
    L.trial_type = take(L.n_pass * (L.n_pos + 16));
    L.trial_len = take(L.n_pass * (L.n_pos * 4 + 16));
    L.n_seg = (L.n_pos + DYN_SEGU - 1) / DYN_SEGU;
    L.walk_recs = take(2 * L.n_pos * sizeof(WalkRec) + 64);   // speculative chains | re-entered chains
    L.walk_segs = take(2 * L.n_seg * sizeof(WalkSeg) + 64);
    L.walk_list = take(L.n_seg * 16 + 64);
    L.total = o;
    return AMBC_OK;
}

extern "C" uint64_t ambc_compress_dynamic_workspace_bytes(uint64_t n, const uint32_t *cands, uint32_t n_cands)
{
    DynLayout L;
    if (dyn_layout(n, cands, n_cands, L)) return 0;
    return L.total;
}

extern "C" int ambc_compress_dynamic_dev(const void *in_dev, uint64_t n, const uint32_t *cands, uint32_t n_cands,
                                         uint32_t method_mask, uint32_t flags, const uint8_t *marker,
                                         uint32_t marker_bytes, void *out_dev, uint64_t out_cap, void *work_dev,
                                         uint64_t work_bytes, ambc_compress_result *res, ambc_chunk_info *map_out,
                                         uint64_t map_cap, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!res || !marker || marker_bytes < 1 || marker_bytes > 4) return ambc_fail(AMBC_E_ARG, "ambc_compress_dynamic_dev: bad argument");
    if ((n && (!in_dev || !work_dev)) || !out_dev) return ambc_fail(AMBC_E_ARG, "ambc_compress_dynamic_dev: null buffer");
    DynLayout L;
    int rc = dyn_layout(n, cands, n_cands, L);
    if (rc) return rc;
    if (work_bytes < L.total) return ambc_fail(AMBC_E_CAPACITY, "ambc_compress_dynamic_dev: workspace too small");
    const uint32_t ovh = marker_bytes + 14;
    uint32_t marker_word = 0;
    for (uint32_t k = 0; k < marker_bytes; k++) marker_word |= (uint32_t)marker[k] << (8 * k);
    memset(res, 0, sizeof(*res));
    res->first_raw = -1;
    uint8_t endp[16], hdr[18];
    memcpy(endp, marker, marker_bytes);
    memset(endp + marker_bytes, 0, 12);
    uint8_t *W = (uint8_t *)work_dev;
    const bool pcr = flags & AMBC_F_PER_CHUNK_RAW;

    const bool native = (method_mask & AMBC_NATIVE_MASK) != 0;
    std::vector<ChunkSpan> spans;
    std::vector<uint8_t> s_type;
    std::vector<uint32_t> s_len;
    uint64_t raw_from = n; // raw_from < n: one raw package [raw_from, n)
    static int dyn_mode = -1; // dev knob: AMBC_DYNAMIC=trial keeps the round-1 path (every size at every multiple of g)
    if (dyn_mode < 0) { const char *e = getenv("AMBC_DYNAMIC"); dyn_mode = (e && !strcmp(e, "trial")) ? 0 : 1; }
    if (n && native && dyn_mode == 1) {
        // ---- the chain, walked on the GPU segment by segment (k_dynamic_walk) ------------------------------------
        WalkCands wc;
        wc.n = n_cands;
        for (uint32_t i = 0; i < n_cands; i++) wc.c[i] = cands[i];
        const uint64_t seglen = (uint64_t)DYN_SEGU * L.g, n_seg = L.n_seg;
        WalkRec *d_spec = (WalkRec *)(W + L.walk_recs), *d_fix = d_spec + L.n_pos;
        WalkSeg *d_sseg = (WalkSeg *)(W + L.walk_segs), *d_fseg = d_sseg + n_seg;
        unsigned long long *d_list = (unsigned long long *)(W + L.walk_list);
        const size_t wsmem = SfCfg<8192>::SMEM;
        CUDA_TRY(cudaFuncSetAttribute(k_dynamic_walk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem));
        k_dynamic_walk<<<(unsigned)min<uint64_t>(n_seg, 0x7fffffffull), SF_T, wsmem, stream>>>(
            (const uint8_t *)in_dev, n, wc, method_mask, ovh, pcr ? 1u : 0u, (uint32_t)L.g, nullptr, n_seg, nullptr, nullptr, d_spec, d_sseg);
        ambc_count_launch();
        CUDA_TRY(cudaGetLastError());
        std::vector<WalkRec> h_spec(L.n_pos), h_fix(L.n_pos);
        std::vector<WalkSeg> h_sseg(n_seg), h_fseg(n_seg), fixseg(n_seg);
        std::vector<uint64_t> fix_entry(n_seg, ~0ull); // entry position the stored re-entered chain of a segment belongs to
        CUDA_TRY(cudaMemcpyAsync(h_spec.data(), d_spec, L.n_pos * sizeof(WalkRec), cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaMemcpyAsync(h_sseg.data(), d_sseg, n_seg * sizeof(WalkSeg), cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
        std::vector<unsigned long long> list;
        for (int round = 0;; round++) {
            // follow the true chain through the segments with what is known; collect the (segment, entry) pairs that
            // still have to be walked (unknown ones are assumed to meet their speculative chain)
            list.clear();
            uint64_t e = 0;
            for (uint64_t sg = 0; sg < n_seg && e < n; sg++) {
                const uint64_t s0 = sg * seglen, s1 = min<uint64_t>(n, s0 + seglen);
                if (e >= s1) continue;
                if (e == s0) { if (h_sseg[sg].stop != ~0ull) break; e = h_sseg[sg].exit; continue; }
                if (fix_entry[sg] == e) {
                    const WalkSeg &f = fixseg[sg];
                    if (f.merged) { if (h_sseg[sg].stop != ~0ull) break; e = h_sseg[sg].exit; }
                    else { if (f.stop != ~0ull) break; e = f.exit; }
                    continue;
                }
                list.push_back(sg); list.push_back(e);
                if (h_sseg[sg].stop != ~0ull) break;
                e = h_sseg[sg].exit;
            }
            if (list.empty()) break;
            if (round > (int)n_seg + 2) return ambc_fail(AMBC_E_CUDA, "internal: the chain stitch does not settle");
            const uint64_t ni = list.size() / 2;
            CUDA_TRY(cudaMemcpyAsync(d_list, list.data(), list.size() * 8, cudaMemcpyHostToDevice, stream));
            k_dynamic_walk<<<(unsigned)min<uint64_t>(ni, 0x7fffffffull), SF_T, wsmem, stream>>>(
                (const uint8_t *)in_dev, n, wc, method_mask, ovh, pcr ? 1u : 0u, (uint32_t)L.g, d_list, ni, d_spec, d_sseg, d_fix, d_fseg);
            ambc_count_launch();
            CUDA_TRY(cudaGetLastError());
            CUDA_TRY(cudaMemcpyAsync(h_fseg.data(), d_fseg, ni * sizeof(WalkSeg), cudaMemcpyDeviceToHost, stream));
            CUDA_TRY(cudaStreamSynchronize(stream));
            // (one copy of the whole record buffer: thousands of short copies cost more than the 16 bytes per g bytes)
            const uint64_t lo = list[0] * DYN_SEGU, hi = min<uint64_t>(L.n_pos, (list[2 * (ni - 1)] + 1) * DYN_SEGU);
            CUDA_TRY(cudaMemcpyAsync(h_fix.data() + lo, d_fix + lo, (hi - lo) * sizeof(WalkRec), cudaMemcpyDeviceToHost, stream));
            CUDA_TRY(cudaStreamSynchronize(stream));
            for (uint64_t k = 0; k < ni; k++) {
                const uint64_t sg = list[2 * k];
                fixseg[sg] = h_fseg[k];
                fix_entry[sg] = list[2 * k + 1];
            }
        }
        // the chain itself
        auto push = [&](const WalkRec &r) {
            ChunkSpan sp; sp.pos = r.pos; sp.size = r.size; sp.pad = 0;
            spans.push_back(sp); s_type.push_back((uint8_t)(r.type_len >> 24)); s_len.push_back(r.type_len & 0xFFFFFFu);
        };
        uint64_t e = 0;
        for (uint64_t sg = 0; sg < n_seg && e < n; sg++) {
            const uint64_t s0 = sg * seglen, s1 = min<uint64_t>(n, s0 + seglen);
            if (e >= s1) continue;
            uint32_t from = 0; // first speculative record that belongs to the true chain
            bool use_spec = true;
            if (e != s0) {
                const WalkSeg &f = fixseg[sg];
                for (uint32_t k = 0; k < f.count; k++) push(h_fix[sg * DYN_SEGU + k]);
                if (f.merged) {
                    while (from < h_sseg[sg].count && h_spec[sg * DYN_SEGU + from].pos != f.exit) from++;
                    if (from == h_sseg[sg].count) return ambc_fail(AMBC_E_CUDA, "internal: meeting point not on the speculative chain");
                } else {
                    use_spec = false;
                    if (f.stop != ~0ull) { raw_from = f.stop; break; }
                    e = f.exit;
                }
            }
            if (use_spec) {
                for (uint32_t k = from; k < h_sseg[sg].count; k++) push(h_spec[sg * DYN_SEGU + k]);
                if (h_sseg[sg].stop != ~0ull) { raw_from = h_sseg[sg].stop; break; }
                e = h_sseg[sg].exit;
            }
        }
    } else {
        // ---- trial passes: sizes only, every multiple of g, one pass per distinct size ----------------
        std::vector<uint8_t> h_type(L.n_pass * L.n_pos);
        std::vector<uint32_t> h_len(L.n_pass * L.n_pos);
        if (n && native) {
            for (uint64_t j = 0; j < L.n_pass; j++) {
                const uint32_t S = L.pass_size[j];
                size_t smem = S <= LZ2_NMAX ? chunkctx_fast_smem_bytes((int)S) : chunkctx_smem_bytes((int)S, (int)S);
                uint8_t *tt = W + L.trial_type + j * (L.n_pos + 16);
                uint32_t *tl = (uint32_t *)(W + L.trial_len + j * (L.n_pos * 4 + 16));
                unsigned grid = (unsigned)min<uint64_t>(L.n_pos, 0x7fffffffull);
                if ((rc = launch_select_span(grid, smem, stream, (const uint8_t *)in_dev, n, S, method_mask, ovh, (uint32_t)L.g,
                                             nullptr, nullptr, tt, tl, L.n_pos))) return rc;
                CUDA_TRY(cudaMemcpyAsync(h_type.data() + j * L.n_pos, tt, L.n_pos, cudaMemcpyDeviceToHost, stream));
                CUDA_TRY(cudaMemcpyAsync(h_len.data() + j * L.n_pos, tl, L.n_pos * 4, cudaMemcpyDeviceToHost, stream));
            }
            CUDA_TRY(cudaStreamSynchronize(stream));
        }

        // ---- the chain (adaptive_compressor.py:363-394 + 537-590) -------------------------------------
        uint64_t pos = 0;
        while (pos < n) {
            const uint64_t remain = n - pos, i = pos / L.g;
            double best_ratio = 1.0;
            uint64_t best_c = remain;
            int best_t = 255;
            uint32_t best_len = 0;
            for (uint32_t ci = 0; ci < n_cands && native; ci++) {
                const uint64_t c = cands[ci] < remain ? cands[ci] : remain;
                if (c > AMBC_NMAX) continue; // no native method is eligible (:114-127)
                const uint64_t j = (uint64_t)L.cand_pass[ci];
                const int t = h_type[j * L.n_pos + i];
                if (t == 255) continue;
                const uint32_t len = h_len[j * L.n_pos + i];
                const double ratio = (double)((uint64_t)len + ovh) / (double)c; // (:573-574)
                if (ratio < best_ratio) { best_ratio = ratio; best_c = c; best_t = t; best_len = len; }
            }
            if (best_t == 255) {
                if (!pcr) { raw_from = pos; break; }                       // rest of the file raw (:586-590)
                best_c = cands[n_cands - 1] < remain ? cands[n_cands - 1] : remain; // labelled extension
                best_len = (uint32_t)best_c;
            }
            ChunkSpan sp; sp.pos = pos; sp.size = (uint32_t)best_c; sp.pad = 0;
            spans.push_back(sp); s_type.push_back((uint8_t)best_t); s_len.push_back(best_len);
            pos += best_c;
        }
    }
    const uint64_t n_list = spans.size();
    if (n_list > L.n_pos) return ambc_fail(AMBC_E_ARG, "internal: chain longer than the position table");
    uint64_t body = 0;
    for (uint64_t k = 0; k < n_list; k++) body += ovh + s_len[k];
    const uint64_t raw_len = raw_from < n ? n - raw_from : 0;
    if (raw_len > 0xFFFFFFFFull) return ambc_fail(AMBC_E_ARG, "raw package over 4 GiB cannot be framed (u32 fields)");
    const uint64_t total_body = body + (raw_len ? ovh + raw_len : 0) + marker_bytes + 12;
    if (total_body > out_cap) return ambc_fail(AMBC_E_CAPACITY, "ambc_compress_dynamic_dev: out_cap too small");

    // ---- encode + frame the chosen chunks ---------------------------------------------------------
    if (n_list) {
        ChunkSpan *d_spans = (ChunkSpan *)(W + L.spans);
        uint8_t *type = W + L.type;
        uint32_t *comp = (uint32_t *)(W + L.comp);
        unsigned long long *offs = (unsigned long long *)(W + L.offs);
        unsigned long long *tiles = (unsigned long long *)(W + L.tiles);
        ScanState *st = (ScanState *)(W + L.state);
        ScanState h_st;
        memset(&h_st, 0, sizeof h_st);
        h_st.first_raw = ~0ull;
        CUDA_TRY(cudaMemcpyAsync(st, &h_st, sizeof h_st, cudaMemcpyHostToDevice, stream));
        CUDA_TRY(cudaMemcpyAsync(d_spans, spans.data(), n_list * sizeof(ChunkSpan), cudaMemcpyHostToDevice, stream));
        uint32_t maxsz = 0;
        for (uint64_t k = 0; k < n_list; k++) maxsz = max(maxsz, spans[k].size);
        if (maxsz > AMBC_NMAX) { // only per-chunk-raw pieces can be this large
            return ambc_fail(AMBC_E_ARG, "per-chunk-raw pieces larger than 8192 bytes are not supported");
        }
        size_t smem = maxsz <= LZ2_NMAX ? chunkctx_fast_smem_bytes((int)maxsz) : chunkctx_smem_bytes((int)maxsz, (int)maxsz);
        unsigned grid = (unsigned)min<uint64_t>(n_list, 0x7fffffffull);
        if ((rc = launch_select_span(grid, smem, stream, (const uint8_t *)in_dev, n, maxsz, method_mask, ovh, 0, d_spans,
                                     W + L.slots, type, comp, n_list))) return rc;
        const unsigned gt = (unsigned)((n_list + SCAN_TILE - 1) / SCAN_TILE);
        const uint32_t sflags = 1u | 2u; // every entry is a package; no END (appended below)
        k_sizes<<<gt, 256, 0, stream>>>(type, comp, 0, n_list, maxsz, n, ovh, sflags, st, tiles);
        ambc_count_launch();
        k_scan_tiles<<<1, 1024, 0, stream>>>(tiles, 0, gt, 1, type, comp, n_list, maxsz, n, ovh, sflags, st, (uint8_t *)out_dev,
                                             out_cap, marker_word, marker_bytes);
        ambc_count_launch();
        k_offsets<<<gt, 256, 0, stream>>>(type, comp, 0, n_list, maxsz, n, ovh, sflags, st, tiles, offs);
        ambc_count_launch();
        size_t psmem = 32 + (((size_t)maxsz + 15) & ~(size_t)15) + 32;
        CUDA_TRY(cudaFuncSetAttribute(k_pack_span, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));
        k_pack_span<<<grid, PACK_BLOCK, psmem, stream>>>((const uint8_t *)in_dev, d_spans, W + L.slots, type, comp, offs,
                                                        marker_word, marker_bytes, (uint8_t *)out_dev, n_list);
        ambc_count_launch();
        CUDA_TRY(cudaGetLastError());
        // the second evaluation must agree with the trial the chain was built from
        std::vector<uint8_t> chk_t(n_list);
        std::vector<uint32_t> chk_l(n_list);
        CUDA_TRY(cudaMemcpyAsync(chk_t.data(), type, n_list, cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaMemcpyAsync(chk_l.data(), comp, n_list * 4, cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaMemcpyAsync(&h_st, st, sizeof h_st, cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
        for (uint64_t k = 0; k < n_list; k++)
            if (chk_t[k] != s_type[k] || chk_l[k] != s_len[k])
                return ambc_fail(AMBC_E_CUDA, "internal: span encode disagrees with its trial at chunk %llu", (unsigned long long)k);
        if (h_st.body_len != body) return ambc_fail(AMBC_E_CUDA, "internal: span body length mismatch");
    }
    // ---- raw tail package (if any) and END ---------------------------------------------------------
    uint8_t *o = (uint8_t *)out_dev + body;
    if (raw_len) {
        memcpy(hdr, marker, marker_bytes);
        hdr[marker_bytes] = 255; hdr[marker_bytes + 1] = 0;
        const uint32_t r32 = (uint32_t)raw_len;
        for (int r = 0; r < 3; r++) memcpy(hdr + marker_bytes + 2 + 4 * r, &r32, 4);
        CUDA_TRY(cudaMemcpyAsync(o, hdr, ovh, cudaMemcpyHostToDevice, stream));
        CUDA_TRY(cudaMemcpyAsync(o + ovh, (const uint8_t *)in_dev + raw_from, raw_len, cudaMemcpyDeviceToDevice, stream));
        o += ovh + raw_len;
    }
    CUDA_TRY(cudaMemcpyAsync(o, endp, marker_bytes + 12, cudaMemcpyHostToDevice, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));

    res->body_len = total_body;
    res->n_chunks = n_list + (raw_len ? 1 : 0);
    res->n_packages = res->n_chunks;
    res->first_raw = raw_len ? (int64_t)n_list : -1;
    for (uint64_t k = 0; k < n_list; k++) {
        const int t = s_type[k];
        if (t >= 1 && t <= 4) { res->usage[t]++; res->payload_bytes += s_len[k]; }
        else res->usage[0]++;
        if (map_out && k < map_cap) { map_out[k].pos = spans[k].pos; map_out[k].orig_len = spans[k].size; map_out[k].comp_len = s_len[k]; map_out[k].type = (uint32_t)t; }
    }
    if (raw_len) {
        res->usage[0]++;
        if (map_out && n_list < map_cap) { map_out[n_list].pos = raw_from; map_out[n_list].orig_len = (uint32_t)raw_len; map_out[n_list].comp_len = (uint32_t)raw_len; map_out[n_list].type = 255; }
    }
    return AMBC_OK;
}

int ambc_lz_levels_compress(const int *levels, int n) { return lz_levels_upload(levels, n); }
int ambc_lz_coop_compress(int t) { return lz_coop_upload(t); }
int ambc_lz_force_buckets_compress(int on) { return lz_force_buckets_upload(on); }
