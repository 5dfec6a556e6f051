// decode_warp.cuh -- two warps (64 lanes) per package for the two kinds that carry most of a decode: Huffman
// (type 3) and RLE (type 1) packages whose payload and output are at most CAP bytes (round 2; the round-1 decoders
// ran one 256-thread CTA per package and re-decoded every Huffman bit range about four times, 18 k warp
// instructions per 4 KiB package, ncu).
//
// Reference behaviour restated (file:line relative to the reference repo):
//   Huffman  compression_methods.py:407-470   (tree rebuilt from the table exactly as the encoder builds it,
//                                              :472-500; bit walk; stops after the append that reaches orig_len)
//   RLE      compression_methods.py:116-152   (pairs, odd tail ignored, truncate / zero pad)
//
// Huffman: the 64 lanes decode 64 bit ranges of the stream.  Only lane 0 knows where its first code starts, so
//   pass 1     every lane decodes its range from the range start and remembers, at NCP - 1 checkpoints (bit
//              boundaries inside the range), the first code start at or behind the boundary and how many symbols came
//              before;
//   re-sync    a lane whose predecessor ended somewhere else restarts there and decodes only until it stands on a
//              checkpoint of its previous chain -- from there on the two chains are the same chain (the next code
//              start is a function of the position), so the old end and the old counts are adopted.  Huffman codes
//              re-synchronise within a few codes, so this costs a fraction of a pass; fixed-length codes never do,
//              which is why ranges are made a multiple of the code length when the tree is flat;
//   output     one more pass from the true starts writes the symbols at scanned offsets.
// A step is one look-up in a 2^LB-entry table built per package that holds up to two codes; longer codes (rare:
// a symbol seen once or twice in the chunk) continue bit by bit through the tree.
// One package occupies 13.6 KB of shared memory; with one warp per package an SM held 16 warps and half of the
// issue slots stayed empty behind the dependent look-ups of a chain (ncu: IPC 1.98), hence 64 lanes per package.
#pragma once
#include "common.cuh"

#define DW_T 64       // lanes per package = threads per CTA
#define DW_LB 10
#ifndef DW_NCP
#define DW_NCP 2   // range pieces between checkpoints; measured on the 1 GiB mixed body: 2 -> 3.50, 3 -> 3.54, 4 -> 3.58 ms
#endif
#ifndef DW_CASCADE
#define DW_CASCADE 2  // re-synchronisation rounds before a code is treated as one that does not re-synchronise
#endif

template <int CAP> struct DwCfg {
    static constexpr int W_BYTES = CAP + 32;            // staged bit words (big-endian), or the RLE tables
    static constexpr int OUT_BYTES = CAP + 32;          // decoded bytes before they go to global memory
    static constexpr int LUT_BYTES = 4 << DW_LB;        // two-symbol table
    static constexpr int TREE_BYTES = 512 + 512 + 256;  // child0[256], child1[256] (u16, internal nodes), leafsym[256]
    static constexpr int PER_PKG = W_BYTES + OUT_BYTES + LUT_BYTES + TREE_BYTES;
    static_assert(W_BYTES + OUT_BYTES >= 513 * 8 + 1024 + 1024 + (2 << DW_LB), "tree-build scratch overlays W + out");
    static_assert(OUT_BYTES >= 4 * (DW_T * (DW_NCP + 1) + 2 + DW_LB * DW_T + DW_T + 1), "checkpoints, range ends and the entry table overlay out");
};

struct DwCtx {
    uint32_t *W;
    uint8_t *out;
    uint32_t *lut;               // [1 << LB] sym1 | sym2 << 8 | n << 16 | l1 << 24 | l << 28; n == 0: node id of an escape
    uint16_t *child0, *child1;
    uint8_t *leafsym;
    int K;
};
template <int CAP> __device__ __forceinline__ void dw_carve(DwCtx &d, uint8_t *base)
{
    d.W = (uint32_t *)base;
    d.out = base + DwCfg<CAP>::W_BYTES;
    d.lut = (uint32_t *)(d.out + DwCfg<CAP>::OUT_BYTES);
    d.child0 = (uint16_t *)((uint8_t *)d.lut + DwCfg<CAP>::LUT_BYTES);
    d.child1 = d.child0 + 256;
    d.leafsym = (uint8_t *)(d.child1 + 256);
    d.K = 0;
}

__device__ __forceinline__ uint32_t dw_lane() { return threadIdx.x; } // 0 .. DW_T - 1

// byte i of a payload of `len` bytes, 0 behind its end (the guards of the round-1 parser)
__device__ __forceinline__ uint32_t dw_byte(const uint8_t *__restrict__ in, int i, int len) { return i < len ? (uint32_t)__ldg(in + i) : 0u; }

// out[0 .. n) (shared, 16-byte aligned) -> dst (global, any alignment); DW_T lanes
__device__ __forceinline__ void dw_store(uint8_t *__restrict__ dst, const uint8_t *out, uint32_t n)
{
    const uint32_t lane = dw_lane();
    uint32_t head = (uint32_t)((16 - ((uintptr_t)dst & 15)) & 15);
    if (head > n) head = n;
    for (uint32_t i = lane; i < head; i += DW_T) dst[i] = out[i];
    const uint32_t body = (n - head) >> 4;
    uint4 *d4 = (uint4 *)(dst + head);
    if (head == 0) {
        const uint4 *s4 = (const uint4 *)out;
        for (uint32_t i = lane; i < body; i += DW_T) d4[i] = s4[i];
    } else {
        const uint8_t *s = out + head;
        for (uint32_t i = lane; i < body; i += DW_T) {
            const uint8_t *p = s + (i << 4);
            uint4 v;
            v.x = lds_u32u(p); v.y = lds_u32u(p + 4); v.z = lds_u32u(p + 8); v.w = lds_u32u(p + 12);
            d4[i] = v;
        }
    }
    for (uint32_t i = head + (body << 4) + lane; i < n; i += DW_T) dst[i] = out[i];
}

// ---- Huffman: table -> tree -> look-up table ----------------------------------------------------------------
// The reference's merges (compression_methods.py:482-494): the two smallest nodes under (weight, leader), lo -> bit
// 0, hi -> bit 1, the merged node is led by lo's leader.  Leaves sorted by (weight, symbol) in keyL (a sentinel
// behind them), merged nodes come out in key order into keyM (filled with sentinels), so two queue fronts decide.
// T = uint32_t when every key fits (weights below 2^15), else 64 bits.  One lane.
template <class T>
__device__ __forceinline__ void dw_merge(const T *keyL, T *keyM, int K, uint16_t *parent, uint16_t *child0, uint16_t *child1)
{
    int li = 0, mi = 0;
    T kl = keyL[0], km = keyM[0];
    for (int t = 0; t < K - 1; t++) {
        const bool aL = kl < km;
        const T ka = aL ? kl : km;
        const int a = aL ? li : K + mi;
        if (aL) { li++; kl = keyL[li]; } else { mi++; km = keyM[mi]; }
        const bool bL = kl < km;
        const T kb = bL ? kl : km;
        const int b = bL ? li : K + mi;
        if (bL) { li++; kl = keyL[li]; } else { mi++; km = keyM[mi]; }
        const T nk = (((ka >> 8) + (kb >> 8)) << 8) | (ka & (T)0xFF);
        keyM[t] = nk;
        if (mi == t) km = nk; // the new node is the front of its queue
        parent[a] = (uint16_t)(K + t);
        parent[b] = (uint16_t)((K + t) | 0x8000);
        child0[t] = (uint16_t)a;
        child1[t] = (uint16_t)b;
    }
}

// Returns 0 or -1 where the reference raises; *boff = offset of the bit stream, *nbits = stream bits to decode,
// *flat = the common code length when every code has the same length (else 0), *maxlen = the longest code.  Collective for the DW_T lanes of the package.
template <int CAP>
__device__ inline int dw_huff_build(DwCtx &d, const uint8_t *__restrict__ in, int len, int *boff, uint32_t *nbits, int *flat, int *maxlen)
{
    const uint32_t tid = dw_lane(), lane = tid & 31, wid = tid >> 5;
    const int ne = (int)dw_byte(in, 0, len);
    if (ne > 0 && 1 + 5 * (ne - 1) >= len) return -1; // IndexError: a table entry's symbol byte lies past the payload (:430)
    // scratch (dead before the stream is staged)
    unsigned long long *keyL = (unsigned long long *)d.W;          // [257] leaves by (weight, symbol), then a sentinel
    unsigned long long *keyM = keyL + 257;                          // [256] merged nodes; first the compacted unsorted keys
    uint16_t *parent = (uint16_t *)(keyM + 256);                    // [512] parent | bit << 15
    uint32_t *lastidx = (uint32_t *)(parent + 512);                 // [256] 1 + last table entry of the symbol
    uint16_t *lut1 = (uint16_t *)(lastidx + 256);                   // [1 << LB] one-symbol table: 0x8000 | len << 8 | sym, else node
    int *sh = (int *)d.child0;                                      // K, wmax, dmin[2], dmax[2] (child0 is written by the merge, later)
    for (int b = tid; b < 256; b += DW_T) lastidx[b] = 0;
    __syncthreads();
    // dict semantics: the last entry of a symbol wins (:436)
    for (int e = tid; e < ne; e += DW_T) atomicMax(&lastidx[dw_byte(in, 1 + 5 * e, len)], (uint32_t)(e + 1));
    __syncthreads();
    if (wid == 0) { // compact the present symbols: key = weight << 8 | symbol
        int K = 0;
        uint32_t wmax = 0;
        for (int b0 = 0; b0 < 256; b0 += 32) {
            const int b = b0 + (int)lane;
            const uint32_t li = lastidx[b];
            const uint32_t m = __ballot_sync(FULL_MASK, li != 0);
            if (li) {
                const int o = 2 + 5 * ((int)li - 1);
                const uint32_t wgt = dw_byte(in, o, len) | (dw_byte(in, o + 1, len) << 8) | (dw_byte(in, o + 2, len) << 16) | (dw_byte(in, o + 3, len) << 24);
                keyM[K + __popc(m & ((1u << lane) - 1))] = ((unsigned long long)wgt << 8) | (unsigned long long)b;
                wmax = max(wmax, wgt);
            }
            K += __popc(m);
        }
        wmax = __reduce_max_sync(FULL_MASK, wmax);
        if (lane == 0) { sh[0] = K; sh[1] = (int)(wmax < (1u << 15)); }
    }
    __syncthreads();
    const int K = sh[0];
    const bool small = sh[1] != 0; // sums stay below 2^23: 32-bit keys
    d.K = K;
    if (K <= 1) return -1; // heappop on an empty heap / code[-1] of an empty code (:497, :527)
    {   // rank sort (keys are distinct); up to 4 keys per lane
        unsigned long long mine[4];
        int rk[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int j = (int)tid + DW_T * q;
            mine[q] = ~0ull; rk[q] = 0;
            if (j < K) {
                mine[q] = keyM[j];
                int r = 0;
                for (int x = 0; x < K; x++) r += (keyM[x] < mine[q]);
                rk[q] = r;
            }
        }
        __syncthreads(); // every rank is known before keyM is given to the merge
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int j = (int)tid + DW_T * q;
            if (j < K) {
                if (small) ((uint32_t *)keyL)[rk[q]] = (uint32_t)mine[q]; else keyL[rk[q]] = mine[q];
                d.leafsym[rk[q]] = (uint8_t)(mine[q] & 0xFFull);
            }
        }
        if (small) {
            for (int j = tid; j < 256; j += DW_T) ((uint32_t *)keyM)[j] = 0xFFFFFFFFu;
            if (tid == 0) ((uint32_t *)keyL)[K] = 0xFFFFFFFFu;
        } else {
            for (int j = tid; j < 256; j += DW_T) keyM[j] = ~0ull;
            if (tid == 0) keyL[K] = ~0ull;
        }
    }
    for (int i = tid; i < (1 << DW_LB) / 2; i += DW_T) ((uint32_t *)lut1)[i] = 0;
    __syncthreads();
    if (tid == 0) {
        if (small) dw_merge<uint32_t>((const uint32_t *)keyL, (uint32_t *)keyM, K, parent, d.child0, d.child1);
        else dw_merge<unsigned long long>(keyL, keyM, K, parent, d.child0, d.child1);
    }
    __syncthreads();
    // every node walks up to the root: depth and the root-side bits of its code.  Leaves of depth <= LB and the
    // internal nodes of depth LB mark the first table entry they own.
    const int root = 2 * K - 2;
    int dmin = 1 << 20, dmax = 0;
    for (int n = tid; n < root; n += DW_T) {
        uint32_t code = 0;
        int dep = 0, x = n;
        while (x != root) {
            const uint32_t p = parent[x];
            code = (code >> 1) | ((p >> 15) << 31);
            x = (int)(p & 0x7FFFu);
            dep++;
        }
        if (n < K) {
            dmin = min(dmin, dep); dmax = max(dmax, dep);
            if (dep <= DW_LB) lut1[code >> (32 - DW_LB)] = (uint16_t)(0x8000u | ((uint32_t)dep << 8) | d.leafsym[n]);
        } else if (dep == DW_LB) lut1[code >> (32 - DW_LB)] = (uint16_t)n;
    }
    dmin = __reduce_min_sync(FULL_MASK, dmin);
    dmax = __reduce_max_sync(FULL_MASK, dmax);
    int *shd = (int *)lastidx; // (dead)
    if (lane == 0) { shd[wid] = dmin; shd[2 + wid] = dmax; }
    __syncthreads();
    dmin = min(shd[0], shd[1]);
    dmax = max(shd[2], shd[3]);
    *flat = dmin == dmax ? dmin : 0;
    *maxlen = dmax;
    // an entry without a mark belongs to the nearest mark before it.  Each warp fills its half of the table: entry
    // 2^(LB-1) (the code "1" followed by zeros) starts the right subtree of the root and always carries a mark.
    uint32_t carry = 0;
    for (int r = (int)wid << (DW_LB - 1); r < (int)(wid + 1) << (DW_LB - 1); r += 32) {
        const uint32_t e = lut1[r + lane];
        const uint32_t m = __ballot_sync(FULL_MASK, e != 0);
        const uint32_t below = m & (0xFFFFFFFFu >> (31 - lane));
        const uint32_t v = __shfl_sync(FULL_MASK, e, below ? 31 - __clz(below) : 0);
        const uint32_t mine = below ? v : carry;
        if (!e) lut1[r + lane] = (uint16_t)mine;
        carry = __shfl_sync(FULL_MASK, mine, 31);
    }
    __syncthreads();
    // two-symbol table: the second code counts when it lies completely inside the LB bits
    for (int i = tid; i < (1 << DW_LB); i += DW_T) {
        const uint32_t e1 = lut1[i];
        uint32_t e;
        if (e1 & 0x8000u) {
            const uint32_t l1 = (e1 >> 8) & 0x7Fu;
            const uint32_t e2 = lut1[(i << l1) & ((1 << DW_LB) - 1)];
            const uint32_t l2 = (e2 >> 8) & 0x7Fu;
            if ((e2 & 0x8000u) && l1 + l2 <= DW_LB) e = (e1 & 0xFFu) | ((e2 & 0xFFu) << 8) | (2u << 16) | (l1 << 24) | ((l1 + l2) << 28);
            else e = (e1 & 0xFFu) | (1u << 16) | (l1 << 24) | (l1 << 28);
        } else e = e1; // n = 0: the node where the walk goes on
        d.lut[i] = e;
    }
    int off = 1 + 5 * ne;
    const uint32_t nb = dw_byte(in, off, len) | (dw_byte(in, off + 1, len) << 8) | (dw_byte(in, off + 2, len) << 16) | (dw_byte(in, off + 3, len) << 24);
    off += 4;
    const uint32_t avail = off < len ? (uint32_t)(len - off) * 8u : 0u;
    *boff = off;
    *nbits = nb < avail ? nb : avail;
    __syncthreads();
    return 0;
}

// stream bytes in[boff ..) -> W as big-endian 32-bit words (two words of slack behind the stream)
__device__ inline void dw_stage_bits(DwCtx &d, const uint8_t *__restrict__ in, int len, int boff, uint32_t nbits)
{
    const uint32_t tid = dw_lane();
    const uint32_t nw = (nbits + 31) >> 5;
    const uintptr_t a = (uintptr_t)(in + boff);
    const uint32_t *g = (const uint32_t *)(a & ~(uintptr_t)3);
    const uint32_t sh = (uint32_t)(a & 3) * 8;
    // aligned words that hold at least one payload byte: [0, gmax)
    const uint32_t gmax = len > boff ? (uint32_t)(((a & 3) + (uint32_t)(len - boff) + 3) >> 2) : 0u;
    for (uint32_t i = tid; i < nw + 2; i += DW_T) {
        const uint32_t g0 = i < gmax ? __ldg(g + i) : 0u;
        const uint32_t g1 = i + 1 < gmax ? __ldg(g + i + 1) : 0u;
        d.W[i] = __byte_perm(__funnelshift_r(g0, g1, sh), 0, 0x0123);
    }
    __syncthreads();
}

__device__ __forceinline__ uint32_t dw_top(const DwCtx &d, uint32_t pos)
{
    const uint32_t wi = pos >> 5;
    return __funnelshift_l(d.W[wi + 1], d.W[wi], pos);
}

// One code at bit `pos` with every check: returns length << 8 | symbol (length 0: no complete code before nbits).
// Not inlined and handed plain pointers: a context struct that escapes to a call would live in local memory and
// turn every shared-memory access of the fast loops into a generic load (ncu: long-scoreboard stalls).
__device__ __noinline__ uint32_t dw_code_slow(const uint32_t *lut, const uint32_t *W, const uint16_t *child0, const uint16_t *child1,
                                              const uint8_t *leafsym, int K, uint32_t pos, uint32_t nbits)
{
    const uint32_t wi = pos >> 5;
    const uint32_t e = lut[__funnelshift_l(W[wi + 1], W[wi], pos) >> (32 - DW_LB)];
    uint32_t l, sym;
    if ((e >> 16) & 3u) {
        l = (e >> 24) & 15u;
        sym = e & 0xFFu;
    } else { // longer than LB bits: on through the tree
        uint32_t node = e & 0xFFFFu, q = pos + DW_LB;
        for (;;) {
            if (q >= nbits) return 0;
            const uint32_t bit = (W[q >> 5] >> (31 - (q & 31))) & 1u;
            node = bit ? child1[node - K] : child0[node - K];
            q++;
            if ((int)node < K) break;
        }
        sym = leafsym[node];
        l = q - pos;
    }
    return pos + l > nbits ? 0u : (l << 8) | sym;
}

// One lane's chain from pos to the first code start at or behind `bound`.  Far from the bound a step is one table
// look-up of one or two codes; within LB bits of it single codes, so that the chain ends on the FIRST code start
// behind the bound whatever the pairing was (checkpoints and range ends must not depend on it: two chains that
// have met would otherwise keep visiting alternate code starts for ever).  cnt counts the codes; EMIT stores the
// symbols at out[o ..] (the caller guarantees room).  DEEP: some codes are longer than LB bits and leave the fast
// loops.  A chain that meets an incomplete code ends at nbits.
template <bool EMIT, bool DEEP>
__device__ __forceinline__ void dw_walk(const DwCtx &d, uint32_t &pos, uint32_t bound, uint32_t nbits, uint32_t &cnt, uint32_t &o)
{
    const uint32_t fa = bound > DW_LB ? bound - DW_LB : 0u;             // pairs cannot overshoot the bound before fa
    const uint32_t fb = min(bound, nbits > DW_LB ? nbits - DW_LB : 0u); // a look-up cannot run past nbits before fb
    for (;;) {
        bool esc = false;
        while (pos < fa) {
            uint32_t top = dw_top(d, pos);
#pragma unroll
            for (int u = 0; u < 3; u++) {
                const uint32_t e = d.lut[top >> (32 - DW_LB)];
                const uint32_t l = e >> 28;
                if (DEEP) { if (l == 0) { esc = true; break; } }
                if (EMIT) {
                    d.out[o] = (uint8_t)e;
                    if (e & (2u << 16)) d.out[o + 1] = (uint8_t)(e >> 8);
                }
                const uint32_t n = __byte_perm(e, 0, 0x4442);
                o += n; cnt += n;
                pos += l;
                top <<= l;
                if (u < 2 && pos >= fa) break;
            }
            if (DEEP) { if (esc) break; }
        }
        if (!esc) {
            while (pos < fb) {
                const uint32_t e = d.lut[dw_top(d, pos) >> (32 - DW_LB)];
                if (DEEP) { if ((e >> 28) == 0) break; }
                if (EMIT) d.out[o] = (uint8_t)e;
                o++; cnt++;
                pos += (e >> 24) & 15u;
            }
        }
        if (pos >= bound) return;
        // an escape, or the last LB bits of the stream: one code with every check
        const uint32_t c = dw_code_slow(d.lut, d.W, d.child0, d.child1, d.leafsym, d.K, pos, nbits);
        if (c == 0) { pos = nbits; return; }
        if (EMIT) d.out[o] = (uint8_t)c;
        o++; cnt++;
        pos += c >> 8;
    }
}

// the same chain, stopping after the append that reaches `limit` symbols (lanes of a stream that holds more symbols
// than orig_len; rare)
__device__ __noinline__ void dw_walk_limited(const uint32_t *lut, const uint32_t *W, const uint16_t *child0, const uint16_t *child1,
                                             const uint8_t *leafsym, int K, uint8_t *out, uint32_t pos, uint32_t bound, uint32_t nbits,
                                             uint32_t o, uint32_t limit)
{
    while (pos < bound && o < limit) {
        const uint32_t c = dw_code_slow(lut, W, child0, child1, leafsym, K, pos, nbits);
        if (c == 0) return;
        out[o++] = (uint8_t)c;
        pos += c >> 8;
    }
}

template <bool DEEP>
__device__ __forceinline__ uint32_t dw_huff_ranges(DwCtx &d, uint32_t nbits, int flat, int dmax, uint32_t limit)
{
    const uint32_t tid = dw_lane();
    // ranges: S bits per lane, S / 32 odd (the lanes' word loads fall into different banks); a flat code of F bits
    // never re-synchronises, so S is made a multiple of F as well
    uint32_t wps = (((nbits + 31) >> 5) + DW_T - 1) / DW_T;
    if (wps == 0) wps = 1;
    uint32_t op = 1;
    if (flat) { op = (uint32_t)flat; while (!(op & 1)) op >>= 1; }
    wps = (wps + op - 1) / op * op;
    if (!(wps & 1)) wps += op;
    const uint32_t S = 32 * wps, G = S / DW_NCP;
    const uint32_t base = min(nbits, tid * S), lim = min(nbits, (tid + 1) * S);
    uint32_t *cp = (uint32_t *)d.out;   // [NCP][DW_T] (pos - base) << 16 | symbols before pos; dead before the output pass
    uint32_t *ends = cp + DW_NCP * DW_T; // [DW_T] where the chain of each lane ends
    uint32_t start = base, pos = base, idx = 0, dummy = 0;
    for (int j = 1; j <= DW_NCP; j++) {
        const uint32_t bound = j < DW_NCP ? min(lim, base + (uint32_t)j * G) : lim;
        dw_walk<false, DEEP>(d, pos, bound, nbits, idx, dummy);
        if (j < DW_NCP) cp[j * DW_T + tid] = ((pos - base) << 16) | idx;
    }
    uint32_t end = pos, cnt = idx;
    for (int iter = 0; iter <= DW_T; iter++) {
        ends[tid] = end;
        __syncthreads();
        const uint32_t ns = tid ? ends[tid - 1] : 0u;
        const bool need = ns != start;
        if (!__syncthreads_or(need)) break; // (also orders the reads of `ends` before the next round's writes)
        if (!DEEP && iter >= DW_CASCADE) {
            // Chains of this code do not meet inside a range (near-fixed-length codes: few symbols of almost equal
            // frequency), so every round fixes one more lane.  Instead: a range is entered less than dmax bits behind
            // its first bit, so every lane walks its range from each of those entries, and one lane follows the
            // true entries through the table.
            uint32_t *tab = ends + DW_T + 2;          // [dmax][DW_T] (end - base) << 16 | codes
            uint32_t *tstart = tab + DW_LB * DW_T;    // [DW_T + 1] true entry of each lane, then the end of the last
            for (int k = 0; k < dmax; k++) {
                uint32_t p = base + (uint32_t)k, c = 0;
                if (p < lim) dw_walk<false, DEEP>(d, p, lim, nbits, c, dummy);
                tab[k * DW_T + tid] = ((p - base) << 16) | c;
            }
            __syncthreads();
            if (tid == 0) {
                uint32_t sp = 0;
                for (uint32_t t = 0; t < DW_T; t++) {
                    const uint32_t bt = min(nbits, t * S), lt = min(nbits, (t + 1) * S);
                    tstart[t] = sp;
#ifdef DW_DEBUG
                    if (sp < lt && (sp < bt || sp - bt >= (uint32_t)dmax))
                        printf("DWBUG entry t=%u sp=%u bt=%u lt=%u dmax=%d nbits=%u S=%u flat=%d K=%d\n", t, sp, bt, lt, dmax, nbits, S, flat, d.K);
#endif
                    if (sp < lt) sp = bt + (tab[(sp - bt) * DW_T + t] >> 16);
                }
                tstart[DW_T] = sp;
            }
            __syncthreads();
            start = tstart[tid];
            end = tstart[tid + 1];
            cnt = start < lim ? tab[(start - base) * DW_T + tid] & 0xFFFFu : 0u;
            __syncthreads();
            break;
        }
        bool merged = false;
        int mj = DW_NCP;
        uint32_t delta = 0;
        if (need) {
            start = ns; pos = ns; idx = 0;
            for (int j = 1; j <= DW_NCP && !merged; j++) {
                const uint32_t bound = j < DW_NCP ? min(lim, base + (uint32_t)j * G) : lim;
                dw_walk<false, DEEP>(d, pos, bound, nbits, idx, dummy);
                if (j < DW_NCP) {
                    const uint32_t old = cp[j * DW_T + tid];
                    if ((old >> 16) == pos - base) { merged = true; mj = j; delta = idx - (old & 0xFFFFu); }
                    else cp[j * DW_T + tid] = ((pos - base) << 16) | idx;
                } else { end = pos; cnt = idx; }
            }
            if (merged) { // the rest of the previous chain stands; its symbol counts shift by delta
                cnt += delta;
                for (int j = mj; j < DW_NCP; j++) {
                    const uint32_t old = cp[j * DW_T + tid];
                    cp[j * DW_T + tid] = (old & 0xFFFF0000u) | ((old + delta) & 0xFFFFu);
                }
            }
        }
    }
    // output pass: symbol offsets = scan of the counts over the DW_T lanes
    const uint32_t inc = (uint32_t)warp_incl_scan((int)cnt);
    if ((tid & 31) == 31) ends[DW_T + (tid >> 5)] = inc; // (ends[DW_T ..]: still inside the checkpoint area's slack)
    __syncthreads();
    const uint32_t w0 = ends[DW_T], total = w0 + ends[DW_T + 1];
    uint32_t o = inc - cnt + (tid >= 32 ? w0 : 0u);
    __syncthreads(); // cp / ends are dead now: out is written
    pos = start; idx = 0;
    if (o + cnt <= limit) dw_walk<true, DEEP>(d, pos, lim, nbits, idx, o);
    else dw_walk_limited(d.lut, d.W, d.child0, d.child1, d.leafsym, d.K, d.out, pos, end, nbits, o, limit);
    __syncthreads();
    return total;
}

// payload in[0 .. len) in global memory -> d.out.  Returns the bytes produced or -1.  Collective (DW_T lanes).
template <int CAP>
__device__ inline int dw_huff(DwCtx &d, const uint8_t *__restrict__ in, int len, int orig)
{
    if (len <= 0) return 0;
    int boff, flat, dmax;
    uint32_t nbits;
    if (dw_huff_build<CAP>(d, in, len, &boff, &nbits, &flat, &dmax) < 0) return -1;
    dw_stage_bits(d, in, len, boff, nbits);
    const uint32_t limit = (uint32_t)max(orig, 1); // stops after the append that reaches orig_len (:464-468)
    const uint32_t total = dmax > DW_LB ? dw_huff_ranges<true>(d, nbits, flat, dmax, limit) : dw_huff_ranges<false>(d, nbits, flat, dmax, limit);
    return (int)min(total, limit);
}

// ---- RLE --------------------------------------------------------------------------------------------------
// Warp 0 scans the pairs 32 at a time: a pair with a non-zero count that starts before orig sets the bit of its
// first output byte and appends its value to a list; output byte x is list[popc(bits 0 .. x) - 1].  A run of
// zeros behind the last pair pads to orig (:145-150).  All lanes expand, 16 output bytes per lane and step.
template <int CAP>
__device__ inline int dw_rle(DwCtx &d, const uint8_t *__restrict__ in, int len, int orig)
{
    if (len <= 0) return 0;
    const uint32_t tid = dw_lane(), lane = tid & 31;
    constexpr int NW = CAP / 32;                 // bitmap words
    uint32_t *bm = d.W;                          // [NW + 1]
    uint16_t *pre = (uint16_t *)(bm + NW + 1);   // [NW + 1] set bits before word w
    uint8_t *vals = (uint8_t *)(pre + NW + 2);   // [CAP / 2 + 1]
    static_assert(4 * (NW + 1) + 2 * (NW + 2) + CAP / 2 + 1 <= DwCfg<CAP>::W_BYTES, "RLE tables fit W");
    for (int w = tid; w <= NW; w += DW_T) bm[w] = 0;
    __syncthreads();
    if (tid < 32) {
        const int P = len >> 1; // complete pairs (:132-133)
        uint32_t run = 0, nv = 0;
        for (int k0 = 0; k0 < P && run < (uint32_t)orig; k0 += 32) {
            const int k = k0 + (int)lane;
            uint32_t v = 0, c = 0;
            if (k < P) { v = __ldg(in + 2 * k); c = __ldg(in + 2 * k + 1); }
            const uint32_t inc = (uint32_t)warp_incl_scan((int)c);
            const uint32_t st = run + inc - c;
            const bool live = c != 0 && st < (uint32_t)orig;
            const uint32_t m = __ballot_sync(FULL_MASK, live);
            if (live) {
                vals[nv + __popc(m & ((1u << lane) - 1))] = (uint8_t)v;
                atomicOr(&bm[st >> 5], 1u << (st & 31));
            }
            nv += __popc(m);
            run += __shfl_sync(FULL_MASK, inc, 31);
        }
        if (run < (uint32_t)orig && lane == 0) { vals[nv] = 0; atomicOr(&bm[run >> 5], 1u << (run & 31)); }
        __syncwarp();
        // set bits before each bitmap word
        constexpr int WPL = NW / 32;
        uint32_t pc[WPL], s = 0;
#pragma unroll
        for (int q = 0; q < WPL; q++) { pc[q] = __popc(bm[lane * WPL + q]); s += pc[q]; }
        uint32_t ex = (uint32_t)warp_incl_scan((int)s) - s;
#pragma unroll
        for (int q = 0; q < WPL; q++) { pre[lane * WPL + q] = (uint16_t)ex; ex += pc[q]; }
    }
    __syncthreads();
    for (int g = tid; 16 * g < orig; g += DW_T) {
        const int x = 16 * g;
        const uint32_t w = bm[x >> 5], sh = x & 31;
        uint32_t slice = (w >> sh) & 0xFFFFu;
        int r = (int)pre[x >> 5] + __popc(w & ((1u << sh) - 1u)) - 1; // the run that covers byte x - 1
        uint4 v4;
        if (slice == 0) {
            const uint32_t b = (uint32_t)vals[r] * 0x01010101u;
            v4 = make_uint4(b, b, b, b);
        } else {
            uint32_t o4[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                uint32_t acc = 0;
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    r += (int)(slice & 1u); slice >>= 1;
                    acc |= (uint32_t)vals[r] << (8 * b);
                }
                o4[q] = acc;
            }
            v4 = make_uint4(o4[0], o4[1], o4[2], o4[3]);
        }
        *(uint4 *)(d.out + x) = v4;
    }
    __syncthreads();
    return orig;
}

__device__ __forceinline__ bool dw_eligible(const ambc_pkg &e, uint32_t cap)
{
    return (e.type == 1 || e.type == 3) && e.comp_len <= cap && e.orig_len <= cap;
}

// CAP = 4096: packages of at most 4096 bytes; CAP = 8192: the rest up to 8192.  One CTA of DW_T lanes per package.
template <int CAP>
__global__ void __launch_bounds__(DW_T)
k_decode_warp(const uint8_t *__restrict__ body, const ambc_pkg *__restrict__ table, uint64_t n_entries,
              uint8_t *__restrict__ out, uint32_t *status)
{
    extern __shared__ uint4 smem4[];
    const uint32_t tid = dw_lane();
    DwCtx d;
    dw_carve<CAP>(d, (uint8_t *)smem4);
    __shared__ uint32_t s_want[DW_T / 32];
    // entry blockIdx.x + k * gridDim.x is this CTA's k-th; the lanes look at DW_T of them at once (one round trip
    // to the table instead of one per entry: many entries belong to another decoder)
    for (uint64_t k0 = 0; blockIdx.x + k0 * gridDim.x < n_entries; k0 += DW_T) {
        const uint64_t mine = blockIdx.x + (k0 + tid) * (uint64_t)gridDim.x;
        bool want = false;
        if (mine < n_entries) {
            const ambc_pkg e = table[mine];
            want = dw_eligible(e, 8192) && (CAP == 4096) == (e.comp_len <= 4096 && e.orig_len <= 4096);
        }
        const uint32_t bal = __ballot_sync(FULL_MASK, want);
        if ((tid & 31) == 0) s_want[tid >> 5] = bal;
        __syncthreads();
        for (int wv = 0; wv < DW_T / 32; wv++) {
            uint32_t todo = s_want[wv];
            while (todo) {
                const int b = __ffs(todo) - 1;
                todo &= todo - 1;
                const ambc_pkg e = table[blockIdx.x + (k0 + 32 * wv + b) * (uint64_t)gridDim.x];
                uint8_t *dst = out + e.dst_off;
                const uint8_t *src = body + e.src_off;
                const int produced = e.type == 3 ? dw_huff<CAP>(d, src, (int)e.comp_len, (int)e.orig_len)
                                                 : dw_rle<CAP>(d, src, (int)e.comp_len, (int)e.orig_len);
                const uint32_t nominal = e.comp_len == 0 ? 0 : e.orig_len; // what the index assumed (nominal_out)
                const uint32_t good = produced < 0 ? 0u : min((uint32_t)produced, e.out_len);
                dw_store(dst, d.out, good);
                if (produced < 0 || (uint32_t)produced != nominal) { // codec raised (:440-442) / malformed stream
                    for (uint32_t k = good + tid; k < e.out_len; k += DW_T) dst[k] = 0;
                    if (tid == 0 && status) atomicAdd(&status[produced < 0 ? 0 : 1], 1u);
                }
                __syncthreads();
            }
        }
        __syncthreads();
    }
}
