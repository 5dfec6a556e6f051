// decode_codec.cuh -- per-package decoders.  One CTA (AMBC_BLOCK threads) per package,
// block-collective functions.  Fast path: payload and output staged in shared memory.
// Slow path (foreign files with oversized packages): one thread, directly on global memory.
//
// Reference behaviour restated (file:line relative to the reference repo):
//   RLE        compression_methods.py:116-152   (pairs, odd tail ignored, truncate / zero pad)
//   Dictionary compression_methods.py:236-281   (token walk incl. its quirks, no padding)
//   Huffman    compression_methods.py:407-470   (tree rebuilt from the table, bit walk)
//   Delta      compression_methods.py:610-638   (running sum, min(len, orig) bytes)
//   raw        compression_methods.py:691-713   (truncate / zero pad)
#pragma once
#include "common.cuh"

#define DEC_OUT_CAP 8192      // largest orig_len decoded in shared memory
#define DEC_OUT_SLACK 512     // LZ may overrun orig_len by one match (<= 255 bytes) before truncation
#define DEC_LUT_BITS 10
#ifndef DEC_SHORT_CODE_BITS
#define DEC_SHORT_CODE_BITS 6   // alphabets whose codes all fit: near-fixed-length codes, decoded for all entry offsets
#endif

struct DecCtx {
    uint8_t *in;   // staged payload, 16-byte aligned, in_cap + 16 bytes
    int in_cap;
    uint8_t *out;  // DEC_OUT_CAP + DEC_OUT_SLACK
    uint8_t *X;    // scratch: max(2*(in_cap/2+1), 12 KiB)
    int *red;      // 16 ints
};

__host__ __device__ inline size_t dec_r16(size_t x) { return (x + 15) & ~(size_t)15; }
__host__ __device__ inline size_t dec_x_bytes(int in_cap)
{
    size_t rle = 2 * ((size_t)in_cap / 2 + 8);
    return dec_r16(rle < 12288 ? 12288 : rle);
}
__host__ __device__ inline size_t decctx_smem_bytes(int in_cap)
{
    return dec_r16((size_t)in_cap) + 16 + DEC_OUT_CAP + DEC_OUT_SLACK + dec_x_bytes(in_cap) + 128;
}
__device__ inline void decctx_carve(DecCtx &d, uint8_t *base, int in_cap)
{
    uint8_t *p = base;
    d.in = p; p += dec_r16((size_t)in_cap) + 16;
    d.in_cap = in_cap;
    d.out = p; p += DEC_OUT_CAP + DEC_OUT_SLACK;
    d.X = p; p += dec_x_bytes(in_cap);
    d.red = (int *)p;
}

// ---- RLE ------------------------------------------------------------------------------------
// payload in d.in[0..len), writes d.out[0..orig).  Returns bytes produced (0 for an empty
// payload, :127-128, else orig).  Collective.
__device__ inline int dec_rle(DecCtx &d, int len, int orig)
{
    if (len <= 0) return 0;
    const int tid = threadIdx.x;
    const int P = len >> 1; // complete pairs (:132-133)
    uint16_t *pos = (uint16_t *)d.X; // start offset of pair k, clamped to orig; pos[P] = end
    const int ppt = (P + AMBC_BLOCK - 1) / AMBC_BLOCK;
    int sum = 0;
    for (int k = tid * ppt; k < min(P, (tid + 1) * ppt); k++) sum += d.in[2 * k + 1];
    int total;
    int run = block_excl_scan(sum, d.red, &total);
    for (int k = tid * ppt; k < min(P, (tid + 1) * ppt); k++) {
        pos[k] = (uint16_t)min(run, orig);
        run += d.in[2 * k + 1];
    }
    if (tid == 0) pos[P] = (uint16_t)min(total, orig);
    __syncthreads();
    const int filled = min(total, orig);
    for (int g = tid; 4 * g < orig; g += AMBC_BLOCK) {
        int j0 = 4 * g;
        uint32_t word = 0;
        if (j0 < filled) {
            int lo = 0, hi = P; // largest k in [0,P) with pos[k] <= j0
            while (hi - lo > 1) {
                int mid = (lo + hi) >> 1;
                if (pos[mid] <= j0) lo = mid; else hi = mid;
            }
            int k = lo;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                int j = j0 + b;
                if (j < filled) {
                    while (pos[k + 1] <= j) k++;
                    word |= (uint32_t)d.in[2 * k] << (8 * b);
                }
            }
        }
        *(uint32_t *)(d.out + j0) = word; // zero pad beyond `filled` (:145-150)
    }
    __syncthreads();
    return orig;
}

// ---- Delta ----------------------------------------------------------------------------------
__device__ inline int dec_delta(DecCtx &d, int len, int orig)
{
    if (len <= 0) return 0;
    const int tid = threadIdx.x;
    const int m = min(len, orig);
    const int bpt = ((m + AMBC_BLOCK - 1) / AMBC_BLOCK + 3) & ~3; // bytes per thread
    const int beg = min(m, tid * bpt), end = min(m, beg + bpt);
    int sum = 0;
    for (int i = beg; i < end; i++) sum += d.in[i];
    int total;
    int run = block_excl_scan(sum, d.red, &total);
    for (int i = beg; i < end; i++) {
        run += d.in[i];
        d.out[i] = (uint8_t)run;
    }
    __syncthreads();
    return m;
}

// ---- Dictionary -----------------------------------------------------------------------------
// Serial token walk with the reference's exact quirks (see oracle/ambc_oracle.c:orc_lz_decompress).
// Works on any address space (shared for the fast path, global for the slow path).
__device__ inline int lz_walk(const uint8_t *in, long len, long orig, uint8_t *out, long out_cap)
{
    if (len <= 0) return 0;
    long pos = 0, o = 0;
    while (pos < len && o < orig) {
        uint8_t flag = in[pos++];
        if (flag == 0) {
            if (pos < len) { uint8_t v = in[pos++]; if (o < out_cap) out[o] = v; o++; }
        } else if (pos + 2 < len) {
            long dist = in[pos] | (in[pos + 1] << 8);
            pos += 2;
            long length = in[pos++];
            long start = o - dist;
            for (long i = 0; i < length; i++) {
                long idx;
                if (start + i < o) {
                    idx = start + i;
                    if (idx < 0) idx += o; // Python negative index
                    if (idx < 0) return -1;
                } else {
                    if (o == 0) return -1;
                    idx = o - 1;
                }
                if (o < out_cap) out[o] = (idx < out_cap) ? out[idx] : 0;
                o++;
            }
        }
    }
    return (int)(o < orig ? o : orig);
}

#define DLZ_WARPS 8
#define DLZ_OUT (4096 + 512)        // per-warp output buffer, packages of at most 4096 bytes
#define DLZ_OUT_BIG (8192 + 512)    // packages of 4097 .. 8192 bytes
#define DLZ_MAX_COMP 8192

// ---- warp-per-package Dictionary decode ----------------------------------------------------
// Tokens of a well-formed payload start at even offsets ("slots"): [0, byte] or
// [flag != 0, dist lo, dist hi, len] (compression_methods.py:236-281).  Slot i+1 is the second
// half of a match token iff slot i starts one, so the token-start mask of 32 slots follows from
// the non-zero-flag mask with one add-carry (the recurrence skipped[i+1] = flag[i] & !skipped[i]).
// One warp: a prefix sum of token output lengths places every token; literals are stored at
// once, matches are copied in order, 32 bytes per step.  Anything irregular (distance 0 or
// reaching before the start of the output) replays the package with the serial walk, which
// restates the reference's quirks exactly.

__device__ __forceinline__ uint32_t dlz_skipped(uint32_t M, uint32_t &carry)
{
    const uint32_t b = M & ~carry;
    const uint32_t follows = (b << 1) | carry;
    const uint32_t even = 0x55555555u;
    const uint32_t odd_starts = b & ~even & ~follows;
    const uint32_t seq = odd_starts + b;
    carry = seq < b ? 1u : 0u; // add overflow
    return (even ^ (seq << 1)) & follows;
}

__device__ __forceinline__ uint32_t dlz_lds8(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void dlz_sts8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// returns bytes produced (before truncation to `cap` by the caller) or -1; out = this warp's buffer (shared memory)
__device__ int dec_lz_warp(const uint8_t *__restrict__ in, int len, int orig, uint8_t *out, int out_cap = DLZ_OUT)
{
    if (len <= 0) return 0;
    const int lane = threadIdx.x & 31;
    const uint32_t so = (uint32_t)__cvta_generic_to_shared(out);
    const int nfull = len >= 4 ? (len - 2) >> 1 : 0; // slots whose token is complete whatever its kind
    uint32_t carry = 0, tail_skipped = 0;
    int o = 0;
    bool stop = false, irregular = false;
    // software pipeline: token bytes of the next 32 slots are loaded while this window is decoded
    uint32_t nf = 0, nb1 = 0, nb2 = 0, nb3 = 0;
    {
        const int slot = lane;
        if (slot < nfull) { const uint8_t *q = in + 2 * slot; nf = q[0]; nb1 = q[1]; nb2 = q[2]; nb3 = q[3]; }
    }
    for (int base = 0; base < nfull && !stop; base += 32) {
        const uint32_t f = nf, b1 = nb1, b2 = nb2, b3 = nb3;
        {
            const int slot = base + 32 + lane;
            nf = nb1 = nb2 = nb3 = 0;
            if (slot < nfull) { const uint8_t *q = in + 2 * slot; nf = q[0]; nb1 = q[1]; nb2 = q[2]; nb3 = q[3]; }
        }
        const int slot = base + lane;
        const uint32_t valid = __ballot_sync(FULL_MASK, slot < nfull);
        const uint32_t M = __ballot_sync(FULL_MASK, f != 0) & valid;
        const uint32_t skipped = dlz_skipped(M, carry);
        // is the first slot after the full region the second half of a match token?
        tail_skipped = nfull - base < 32 ? (skipped >> (nfull - base)) & 1u : carry;
        const uint32_t S = ~skipped & valid;
        const bool is_start = (S >> lane) & 1u;
        const bool is_match = is_start && f != 0;
        const int mlen = (int)b3, dist = (int)(b1 | (b2 << 8));
        const int v = is_start ? (is_match ? mlen : 1) : 0;
        int inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int t = __shfl_up_sync(FULL_MASK, inc, d);
            if (lane >= d) inc += t;
        }
        const int off = o + inc - v;
        const bool exec = is_start && off < orig; // the walk stops once orig_len bytes exist (:249)
        if (__ballot_sync(FULL_MASK, is_start && !exec)) stop = true;
        if (exec && !is_match) out[off] = (uint8_t)b1;
        const bool bad = exec && is_match && (dist == 0 || dist > off);
        if (__ballot_sync(FULL_MASK, bad)) { irregular = true; break; }
        uint32_t mm = __ballot_sync(FULL_MASK, exec && is_match);
        __syncwarp();
        // matches in order, one warp-wide copy each.  32-bit shared addresses and explicit ld / st.shared: with a
        // generic `out` the copy loop alone cost 17 warp instructions per match (ncu, 35 with its bookkeeping).
        // (destination address | length << 20 and the source address are formed once per window by every lane; a
        // match then costs two shuffles: shared-window addresses are below 2^18)
        const uint32_t recD = (so + (uint32_t)off) | ((uint32_t)mlen << 20);
        const uint32_t recS = so + (uint32_t)off - (uint32_t)dist;
        while (mm) {
            const int i = __ffs(mm) - 1;
            mm &= mm - 1;
            const uint32_t a = __shfl_sync(FULL_MASK, recD, i);
            const uint32_t srca = __shfl_sync(FULL_MASK, recS, i);
            const uint32_t ml = a >> 20, dsta = a & 0xFFFFFu, mdist = dsta - srca;
            if (ml <= min(mdist, 32u)) { // the usual match: at most 32 bytes, source in front of the destination
                if ((uint32_t)lane < ml) dlz_sts8(dsta + lane, dlz_lds8(srca + lane));
            } else if (mdist >= ml) {
                for (uint32_t t = lane; t < ml; t += 32) dlz_sts8(dsta + t, dlz_lds8(srca + t));
            } else if (mdist == 1u) { // a run of one byte
                const uint32_t v = dlz_lds8(srca);
                for (uint32_t t = lane; t < ml; t += 32) dlz_sts8(dsta + t, v);
            } else { // overlapping copy = periodic extension of the last `dist` bytes; t mod dist without a division
                const float inv = 1.0f / (float)mdist;
                for (uint32_t t = lane; t < ml; t += 32) {
                    const uint32_t q = (uint32_t)(((float)t + 0.5f) * inv);
                    dlz_sts8(dsta + t, dlz_lds8(srca + t - q * mdist));
                }
            }
            __syncwarp();
        }
        const int tot = __shfl_sync(FULL_MASK, inc, 31);
        if (stop) { // o = end of the last executed token
            const int endv = exec ? off + v : 0;
            int m = endv;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) m = max(m, __shfl_xor_sync(FULL_MASK, m, d));
            o = max(m, o);
        } else o += tot;
    }
    int result;
    if (irregular) {
        int r = 0;
        if (lane == 0) r = lz_walk(in, len, orig, out, out_cap);
        result = __shfl_sync(FULL_MASK, r, 0);
    } else {
        if (!stop) { // the last <= 3 bytes: incomplete tokens, serial rules
            int r = o;
            if (lane == 0) {
                long pos = 2L * nfull + (tail_skipped ? 2 : 0);
                long oo = o;
                while (pos < len && oo < orig) {
                    uint8_t flag = in[pos++];
                    if (flag == 0) {
                        if (pos < len) { uint8_t vb = in[pos++]; if (oo < out_cap) out[oo] = vb; oo++; }
                    } else if (pos + 2 < len) {
                        r = -2; break; // cannot happen: a complete match token lies in the full region
                    }
                }
                if (r != -2) r = (int)oo;
            }
            o = __shfl_sync(FULL_MASK, r, 0);
            if (o == -2) {
                int r2 = 0;
                if (lane == 0) r2 = lz_walk(in, len, orig, out, out_cap);
                o = __shfl_sync(FULL_MASK, r2, 0);
                __syncwarp();
                return o;
            }
        }
        result = o < orig ? o : orig;
    }
    __syncwarp();
    return result;
}

__device__ inline int dec_lz(DecCtx &d, int len, int orig)
{
    volatile int *res = d.red;
    if (orig <= DEC_OUT_CAP) { // warp 0 decodes, same code as k_decode_lz
        if (threadIdx.x < 32) {
            int r = dec_lz_warp(d.in, len, orig, d.out, DEC_OUT_CAP + DEC_OUT_SLACK);
            if (threadIdx.x == 0) res[24] = r;
        }
    } else if (threadIdx.x == 0) res[24] = lz_walk(d.in, len, orig, d.out, DEC_OUT_CAP + DEC_OUT_SLACK);
    __syncthreads();
    int r = res[24];
    __syncthreads();
    return r;
}

// ---- Huffman --------------------------------------------------------------------------------
struct HuffDec {
    unsigned long long *nodeW; // [512] node weights (leaves sorted by (weight, symbol), then merges)
    unsigned long long *key;   // [256] weight << 8 | symbol; dead after the rank sort
    uint16_t *lead;            // [512] leader symbol of the node (leaf: its symbol)
    uint16_t *child0, *child1; // [512]
    uint16_t *lut;             // [1 << DEC_LUT_BITS]: leaf -> 0x8000 | len << 8 | sym ; else node id
    uint32_t *firstidx;        // [256] first table entry of the symbol
    uint32_t *lastidx;         // [256] last table entry (overlays lut, dead before it is built)
    uint32_t *start;           // [AMBC_BLOCK + 4] subsequence start bits (overlays key)
    uint32_t *cnt;             // [AMBC_BLOCK] symbols per subsequence (overlays firstidx)
    int K, root;
    int short_codes; // no code is longer than DEC_SHORT_CODE_BITS
};
__device__ inline HuffDec huffdec_scratch(uint8_t *X) // needs 12288 bytes
{
    HuffDec h;
    h.nodeW = (unsigned long long *)X;            // 0     .. 4096
    h.key = (unsigned long long *)(X + 4096);     // 4096  .. 6144
    h.lead = (uint16_t *)(X + 6144);              // 6144  .. 7168
    h.child0 = (uint16_t *)(X + 7168);            // 7168  .. 8192
    h.child1 = (uint16_t *)(X + 8192);            // 8192  .. 9216
    h.lut = (uint16_t *)(X + 9216);               // 9216  .. 11264
    h.firstidx = (uint32_t *)(X + 11264);         // 11264 .. 12288
    h.lastidx = (uint32_t *)(X + 9216);
    h.start = (uint32_t *)(X + 4096);
    h.cnt = (uint32_t *)(X + 11264); // firstidx is dead once the tree is built
    static_assert(4 * (AMBC_BLOCK + 4) <= 2048 && 4 * AMBC_BLOCK <= 1024, "Huffman decode scratch layout");
    h.K = 0; h.root = 0; h.short_codes = 0;
    return h;
}

// Decode one code starting at bit `pos` of the MSB-first stream `bits` (readable 8 bytes past
// the last stream byte).  Returns the symbol and its length via *len (0 = no complete code
// before nbits).
__device__ __forceinline__ int huff_next(const HuffDec &h, const uint8_t *bits, uint32_t pos, uint32_t nbits, int *len)
{
    const uint8_t *p = bits + (pos >> 3);
    uint32_t w = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3];
    uint32_t top = (w << (pos & 7)) >> (32 - DEC_LUT_BITS);
    uint32_t e = h.lut[top];
    if (e & 0x8000u) {
        int l = (e >> 8) & 0x7F;
        if (pos + l > nbits) { *len = 0; return 0; }
        *len = l;
        return e & 0xFF;
    }
    int node = (int)e;
    uint32_t q = pos + DEC_LUT_BITS;
    for (;;) {
        if (q >= nbits) { *len = 0; return 0; }
        uint32_t bit = (bits[q >> 3] >> (7 - (q & 7))) & 1;
        node = bit ? h.child1[node] : h.child0[node];
        q++;
        if (node < h.K) { *len = (int)(q - pos); return h.lead[node]; }
    }
}

// Parse the table, rebuild the reference's tree (same rule as the encoder) and the lookup
// table.  in[0..len).  Returns 0, or -1 where the reference raises (IndexError).  On success
// *bits_off = offset of the bit stream, *nbits = number of stream bits to decode.  Collective.
__device__ inline int huffdec_build(DecCtx &d, HuffDec &h, const uint8_t *in, int len, int *bits_off, uint32_t *nbits)
{
    const int tid = threadIdx.x;
    const int ne = in[0];
    // IndexError when a table entry's symbol byte lies past the payload (:430)
    if (ne > 0 && 1 + 5 * (ne - 1) >= len) return -1;
    uint32_t *firstidx = h.firstidx, *lastidx = h.lastidx;
    for (int b = tid; b < 256; b += AMBC_BLOCK) { firstidx[b] = 0xFFFFFFFFu; lastidx[b] = 0; }
    __syncthreads();
    for (int e = tid; e < ne; e += AMBC_BLOCK) {
        int s = in[1 + 5 * e];
        atomicMin(&firstidx[s], (uint32_t)e);
        atomicMax(&lastidx[s], (uint32_t)e);
    }
    __syncthreads();
    // key = weight << 8 | sym for present symbols (dict semantics: last value wins, :436)
    int present = 0;
    for (int b = tid; b < 256; b += AMBC_BLOCK) {
        unsigned long long k = ~0ull;
        if (firstidx[b] != 0xFFFFFFFFu) {
            int o = 2 + 5 * (int)lastidx[b];
            uint32_t wgt = 0;
            for (int t = 0; t < 4; t++) if (o + t < len) wgt |= (uint32_t)in[o + t] << (8 * t);
            k = ((unsigned long long)wgt << 8) | (unsigned long long)b;
            present++;
        }
        h.key[b] = k;
    }
    const int K = block_sum(present, d.red);
    h.K = K;
    if (K <= 1) return -1; // heappop on an empty heap / code[-1] of an empty code (:497, :527)
    {   // rank sort over the K present keys only (compacted into nodeW[256..512), dead until the merge)
        unsigned long long *ck = h.nodeW + 256;
        volatile int *cw = d.red + 16;
        const int lane = tid & 31, w = tid >> 5;
        unsigned long long key = ~0ull;
        if (tid < 256) key = h.key[tid];
        const uint32_t m = __ballot_sync(FULL_MASK, key != ~0ull);
        if (tid < 256 && lane == 0) cw[w] = __popc(m);
        __syncthreads();
        int base = 0;
        for (int i = 0; i < w && i < 8; i++) base += cw[i];
        if (key != ~0ull) ck[base + __popc(m & ((1u << lane) - 1))] = key;
        __syncthreads();
        unsigned long long mine = ~0ull;
        int r = 0;
        if (tid < K) {
            mine = ck[tid];
            for (int j = 0; j < K; j++) r += (ck[j] < mine);
        }
        __syncthreads(); // ck aliases nodeW[256..): all ranks are computed before anything is written
        if (tid < K) {
            h.nodeW[r] = mine >> 8;
            h.lead[r] = (uint16_t)(mine & 0xFF);
        }
    }
    __syncthreads();
    if (tid == 0) {
        int li = 0, mi = K, t = K;
        for (int it = 0; it < K - 1; it++) {
            int pick[2];
#pragma unroll
            for (int z = 0; z < 2; z++) {
                bool hasL = li < K, hasM = mi < t;
                bool takeL;
                if (hasL && hasM) {
                    unsigned long long wl = h.nodeW[li], wm = h.nodeW[mi];
                    takeL = (wl < wm) || (wl == wm && h.lead[li] < h.lead[mi]);
                } else takeL = hasL;
                pick[z] = takeL ? li++ : mi++;
            }
            h.nodeW[t] = h.nodeW[pick[0]] + h.nodeW[pick[1]];
            h.lead[t] = h.lead[pick[0]];
            h.child0[t] = (uint16_t)pick[0];
            h.child1[t] = (uint16_t)pick[1];
            t++;
        }
    }
    __syncthreads();
    h.root = 2 * K - 2;
    int longc = 0;
    for (int idx = tid; idx < (1 << DEC_LUT_BITS); idx += AMBC_BLOCK) {
        int node = h.root, l = 0;
        while (node >= K && l < DEC_LUT_BITS) {
            int bit = (idx >> (DEC_LUT_BITS - 1 - l)) & 1;
            node = bit ? h.child1[node] : h.child0[node];
            l++;
        }
        h.lut[idx] = node < K ? (uint16_t)(0x8000u | (l << 8) | h.lead[node]) : (uint16_t)node;
        if (node >= K || l > DEC_SHORT_CODE_BITS) longc = 1;
    }
    h.short_codes = !__syncthreads_or(longc);
    int off = 1 + 5 * ne;
    uint32_t nb = 0;
    for (int t = 0; t < 4; t++) if (off + t < len) nb |= (uint32_t)in[off + t] << (8 * t);
    off += 4;
    uint32_t avail = off < len ? (uint32_t)(len - off) * 8u : 0u;
    *bits_off = off;
    *nbits = nb < avail ? nb : avail;
    __syncthreads();
    return 0;
}

// Codes of at most E <= 8 bits (small alphabets): near-fixed-length codes never re-synchronise, so
// instead of iterating, every thread decodes its bit segment for EVERY entry offset 0..E-1 (entries
// whose first code stays below E chain into an entry already computed), the true entry offsets are
// composed across segments, and a last pass decodes from them.  Returns symbols produced.  Collective.
__device__ inline int dec_huff_short_codes(DecCtx &d, HuffDec &h, const uint8_t *bits, uint32_t nbits, int orig, int E)
{
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    uint8_t *xo = (uint8_t *)h.nodeW;           // [AMBC_BLOCK][8] exit offset into the next segment, 0xFF = stream ended
    uint8_t *cn = xo + AMBC_BLOCK * 8;          // [AMBC_BLOCK][8] symbols decoded
    uint8_t *ent = cn + AMBC_BLOCK * 8;         // [AMBC_BLOCK] true entry offset of each segment (0xFF: none)
    uint8_t *gx = ent + AMBC_BLOCK;             // [AMBC_WARPS][8] exits of groups of 32 segments
    uint8_t *ge = gx + AMBC_WARPS * 8;          // [AMBC_WARPS] entry offset of each group
    static_assert(AMBC_BLOCK * 17 + AMBC_WARPS * 9 <= 4096 + 2048, "short-code tables overlay nodeW + key");
    const uint32_t S = max(32u, (nbits + AMBC_BLOCK - 1) / AMBC_BLOCK);
    const uint32_t seg0 = (uint32_t)tid * S;
    const uint32_t lim = min(nbits, seg0 + S);
    for (int e = E - 1; e >= 0; e--) {
        uint32_t pos = seg0 + (uint32_t)e;
        uint32_t x = 0xFFu, c = 0;
        if (pos < lim) {
            int l;
            huff_next(h, bits, pos, nbits, &l);
            if (l == 0) x = 0xFFu;                                   // incomplete tail code: nothing more decodes
            else if (e + l < E && pos + (uint32_t)l < lim) { x = xo[tid * 8 + e + l]; c = (uint32_t)cn[tid * 8 + e + l] + 1u; }
            else {
                pos += (uint32_t)l; c = 1;
                while (pos < lim) {
                    huff_next(h, bits, pos, nbits, &l);
                    if (l == 0) { pos = nbits; break; }
                    pos += (uint32_t)l; c++;
                }
                x = pos >= nbits ? 0xFFu : pos - lim;
            }
        } else if (seg0 + (uint32_t)e < nbits) x = 0xFFu; // (cannot happen: S >= E)
        xo[tid * 8 + e] = (uint8_t)x;
        cn[tid * 8 + e] = (uint8_t)c;
    }
    __syncthreads();
    // true entry offsets: groups of 32 segments composed for all entry offsets, groups chained, groups walked
    if (lane < E) {
        uint32_t x = (uint32_t)lane;
        for (int sgm = 32 * wid; sgm < 32 * wid + 32; sgm++)
            if (x != 0xFFu) x = xo[sgm * 8 + x];
        gx[wid * 8 + lane] = (uint8_t)x;
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t x = 0;
        for (int g = 0; g < AMBC_WARPS; g++) { ge[g] = (uint8_t)x; if (x != 0xFFu) x = gx[g * 8 + x]; }
    }
    __syncthreads();
    if (tid < AMBC_WARPS) {
        uint32_t x = ge[tid];
        for (int sgm = 32 * tid; sgm < 32 * tid + 32; sgm++) { ent[sgm] = (uint8_t)x; if (x != 0xFFu) x = xo[sgm * 8 + x]; }
    }
    __syncthreads();
    const uint32_t en = ent[tid];
    const int mine = en == 0xFFu ? 0 : (int)cn[tid * 8 + en];
    int total;
    int o = block_excl_scan(mine, d.red, &total);
    const int limit = max(orig, 1); // stops after the append that reaches orig_len (:464-468)
    if (en != 0xFFu) {
        uint32_t pos = seg0 + en;
        while (pos < lim && o < limit) {
            int l;
            const int sym = huff_next(h, bits, pos, nbits, &l);
            if (l == 0) break;
            if (o < DEC_OUT_CAP + DEC_OUT_SLACK) d.out[o] = (uint8_t)sym;
            pos += (uint32_t)l; o++;
        }
    }
    __syncthreads();
    return min(total, limit);
}

// Fast path: payload in d.in[0..len) (zero padded 16 bytes), output to d.out.  Self-synchronising
// parallel decode: every thread decodes one bit range; ranges re-align to the previous
// thread's true end until nothing moves.  Returns symbols produced or -1.  Collective.
__device__ inline int dec_huff(DecCtx &d, int len, int orig)
{
    if (len <= 0) return 0;
    const int tid = threadIdx.x;
    HuffDec h = huffdec_scratch(d.X);
    int boff;
    uint32_t nbits;
    if (huffdec_build(d, h, d.in, len, &boff, &nbits) < 0) return -1;
    const uint8_t *bits = d.in + boff;
    uint32_t *start = h.start, *cnt = h.cnt;
    // every code fits DEC_SHORT_CODE_BITS (small alphabet): all-entry-offsets decode
    // (segments of at most 255 bits keep the per-entry symbol counts within a byte)
    if (h.short_codes && nbits > 0 && nbits <= 255u * AMBC_BLOCK) return dec_huff_short_codes(d, h, bits, nbits, orig, DEC_SHORT_CODE_BITS);
    const uint32_t S = max(32u, (nbits + AMBC_BLOCK - 1) / AMBC_BLOCK);
    if (tid == 0) start[0] = 0;
    start[tid + 1] = min(nbits, (uint32_t)(tid + 1) * S); // provisional
    __syncthreads();
    const uint32_t lim = min(nbits, (uint32_t)(tid + 1) * S);
    uint32_t my_start = 0xFFFFFFFFu, my_end = lim, my_cnt = 0; // (my_end: the provisional start of the next range)
    for (int iter = 0; iter <= AMBC_BLOCK; iter++) {
        const uint32_t st = start[tid];
        bool changed = false;
        if (st != my_start) { // a range whose start did not move decodes to the same end: nothing to redo
            uint32_t pos = st, c = 0;
            while (pos < lim) {
                int l;
                huff_next(h, bits, pos, nbits, &l);
                if (l == 0) { pos = nbits; break; } // incomplete tail code: nothing more decodes
                pos += l; c++;
            }
            if (pos > nbits) pos = nbits;
            changed = my_end != pos;
            my_start = st; my_end = pos; my_cnt = c;
        }
        __syncthreads();
        if (changed) start[tid + 1] = my_end;
        int any = __syncthreads_or(changed);
        if (!any) break;
    }
    cnt[tid] = my_cnt;
    __syncthreads();
    int total;
    int o = block_excl_scan((int)cnt[tid], d.red, &total);
    const int limit = max(orig, 1); // stops after the append that reaches orig_len (:464-468)
    {
        uint32_t pos = start[tid];
        while (pos < lim && o < limit) {
            int l;
            int sym = huff_next(h, bits, pos, nbits, &l);
            if (l == 0) break;
            if (o < DEC_OUT_CAP + DEC_OUT_SLACK) d.out[o] = (uint8_t)sym;
            pos += l; o++;
        }
    }
    __syncthreads();
    return min(total, limit);
}

// ---- slow paths: one thread, global memory, any size -------------------------------------------
__device__ inline long slow_rle(const uint8_t *in, long len, long orig, uint8_t *out, long cap)
{
    if (len <= 0) return 0;
    long o = 0;
    for (long i = 0; i + 1 < len && o < orig; i += 2) {
        uint8_t v = in[i];
        long c = in[i + 1];
        for (long k = 0; k < c && o < orig; k++, o++) if (o < cap) out[o] = v;
    }
    for (; o < orig; o++) if (o < cap) out[o] = 0;
    return orig;
}
__device__ inline long slow_delta(const uint8_t *in, long len, long orig, uint8_t *out, long cap)
{
    if (len <= 0) return 0;
    long m = len < orig ? len : orig;
    uint8_t prev = 0;
    for (long i = 0; i < m; i++) { prev = (uint8_t)(prev + in[i]); if (i < cap) out[i] = prev; }
    return m;
}
