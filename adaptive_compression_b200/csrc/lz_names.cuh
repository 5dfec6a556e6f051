// lz_names.cuh -- exact earliest-longest match search of the Dictionary method for chunks of at
// most LZ2_NMAX bytes (the whole chunk lies inside window_size = 4096, compression_methods.py:187).
//
// The reference scans every earlier position i < p in ascending order and keeps the first
// strictly longer match, capped at min(32, n - p) (compression_methods.py:283-313).  So
//   mlen[p] = the largest L <= min(32, n - p) whose L-gram at p already occurred at some i < p, and
//   mpos[p] = the FIRST occurrence of that L-gram.
// Let name_L[p] = first position whose L-gram equals the one at p ("name"; name_L[p] == p: p is the
// head of its class).  Then mlen[p] = max { L : name_L[p] < p } and mpos[p] = name_mlen[p].
// Names are built without comparing strings:
//   name_4      hash of the 4 raw bytes (verified against the data),
//   name_3      only for heads of name_4 (a 4-byte match implies the 3-byte one),
//   name_2k     from the pair (name_k[p], name_k[p + k])                        (k = 4, 8, 16),
//   name_{k+j}  from the pair (name_k[p], name_k[p + j]), 0 < j < k, only for positions that
//               are heads at 2k and whose k-gram occurs elsewhere (everybody else either has a
//               match of >= 2k bytes or no k-byte match at all).
// Every "first occurrence of a key" is one open-addressing insert with atomicMin on the position;
// keys are verified through the name arrays, so the result is exact for any hash function.
// Block-collective; returns false when a table overflowed (caller falls back to the bucket search).
#pragma once

#define LZ2_EMPTY 0xFFFFFFFFu
#define LZ2_GOLD 2654435761u
#define LZ2_RSLOTS 4096          // table slots of a refinement pass (plist / islot live above them)
#define LZ2_PART_TARGET 1365     // expected entries per refinement pass (load factor 1/3)

__device__ __forceinline__ void lz2_clear(uint32_t *T, int slots)
{
    for (int i = threadIdx.x * 4; i < slots; i += AMBC_BLOCK * 4)
        *(uint4 *)(T + i) = make_uint4(LZ2_EMPTY, LZ2_EMPTY, LZ2_EMPTY, LZ2_EMPTY);
}

__device__ __forceinline__ bool lz2_ns(const uint16_t *S, const uint32_t *fol, int p)
{
    return S[p] != p || ((fol[p >> 5] >> (p & 31)) & 1u);
}

// first occurrence of the raw key `w` (the low `bytes` bytes at sd + p): insert p, return the slot
__device__ __forceinline__ uint32_t lz2_insert_raw(uint32_t *T, const uint8_t *sd, uint32_t w, uint32_t kmask, int p)
{
    volatile uint32_t *V = T;
    uint32_t s = (w * LZ2_GOLD) >> (32 - 13);
    for (;;) {
        uint32_t q = V[s];
        if (q == LZ2_EMPTY) {
            q = atomicCAS(&T[s], LZ2_EMPTY, (uint32_t)p);
            if (q == LZ2_EMPTY) return s;
        }
        if ((lds_u32u(sd + q) & kmask) == w) {
            if ((uint32_t)p < q) atomicMin(&T[s], (uint32_t)p);
            return s;
        }
        s = (s + 1) & (LZ2_TSLOTS - 1);
    }
}

// first occurrence of the key (S[p], S[p + j], tag) in a table of (mask + 1) slots.
// Returns the slot, or 0xFFFF and sets *overflow when the table is full.
__device__ __forceinline__ uint32_t lz2_insert_pair(uint32_t *T, uint32_t mask, uint32_t h, const uint16_t *S,
                                                    uint32_t a, uint32_t b, int p, int j, uint32_t tag, int *overflow)
{
    volatile uint32_t *V = T;
    uint32_t s = h & mask;
    const uint32_t val = (tag << 16) | (uint32_t)p;
    for (uint32_t probes = 0; probes <= mask; probes++) {
        uint32_t q = V[s];
        if (q == LZ2_EMPTY) {
            q = atomicCAS(&T[s], LZ2_EMPTY, val);
            if (q == LZ2_EMPTY) return s;
        }
        if ((q >> 16) == tag) {
            const uint32_t qp = q & 0xFFFFu;
            if (S[qp] == a && S[qp + j] == b) {
                if (val < q) atomicMin(&T[s], val);
                return s;
            }
        }
        s = (s + 1) & mask;
    }
    *overflow = 1;
    return 0xFFFFu;
}

// name_4 for every position (D), follower bits (folD), matches of length 4.  Returns "any match".
__device__ inline int lz2_level4(ChunkCtx &c, uint16_t *D, uint32_t *folD)
{
    const int n = c.n, tid = threadIdx.x;
    lz2_clear(c.T, LZ2_TSLOTS);
    for (int i = tid; i < LZ2_NMAX / 32; i += AMBC_BLOCK) folD[i] = 0;
    __syncthreads();
    const int P = n - 3;
    for (int p = tid; p < n; p += AMBC_BLOCK) {
        uint32_t slot = 0xFFFFu;
        if (p < P) slot = lz2_insert_raw(c.T, c.sd, lds_u32u(c.sd + p), 0xFFFFFFFFu, p);
        D[p] = (uint16_t)slot;
    }
    __syncthreads();
    int nonhead = 0;
    for (int p = tid; p < n; p += AMBC_BLOCK) {
        const uint32_t slot = D[p];
        uint32_t nm = (uint32_t)p;
        if (slot != 0xFFFFu) nm = c.T[slot];
        D[p] = (uint16_t)nm;
        if (nm < (uint32_t)p) {
            c.mlen[p] = 4; c.mpos[p] = (uint16_t)nm;
            atomicOr(&folD[nm >> 5], 1u << (nm & 31));
            nonhead = 1;
        }
    }
    return __syncthreads_or(nonhead);
}

// matches of length exactly 3: only heads of name_4 can have one.  tmp: scratch names buffer.
__device__ inline void lz2_level3(ChunkCtx &c, const uint16_t *N4, uint16_t *tmp)
{
    const int n = c.n, tid = threadIdx.x;
    lz2_clear(c.T, LZ2_TSLOTS);
    __syncthreads();
    for (int p = tid; p < n; p += AMBC_BLOCK) {
        uint32_t slot = 0xFFFFu;
        if (p + 3 <= n && N4[p] == p) // (positions past n - 4 are their own name)
            slot = lz2_insert_raw(c.T, c.sd, lds_u32u(c.sd + p) & 0xFFFFFFu, 0xFFFFFFu, p);
        tmp[p] = (uint16_t)slot;
    }
    __syncthreads();
    for (int p = tid; p < n; p += AMBC_BLOCK) {
        const uint32_t slot = tmp[p];
        if (slot != 0xFFFFu) {
            const uint32_t nm = c.T[slot];
            if (nm < (uint32_t)p) { c.mlen[p] = 3; c.mpos[p] = (uint16_t)nm; }
        }
    }
    __syncthreads();
}

// name_2k (D) from name_k (S).  Positions whose k-gram at p or at p + k occurs nowhere else are
// heads by construction and do not enter the table.  Returns "any match of 2k bytes".
__device__ inline int lz2_double(ChunkCtx &c, const uint16_t *S, uint16_t *D, const uint32_t *folS, uint32_t *folD, int k)
{
    const int n = c.n, tid = threadIdx.x;
    lz2_clear(c.T, LZ2_TSLOTS);
    for (int i = tid; i < LZ2_NMAX / 32; i += AMBC_BLOCK) folD[i] = 0;
    __syncthreads();
    const int P = n - 2 * k + 1;
    int dummy = 0;
    for (int p = tid; p < n; p += AMBC_BLOCK) {
        uint32_t slot = 0xFFFFu;
        if (p < P && lz2_ns(S, folS, p) && lz2_ns(S, folS, p + k)) {
            const uint32_t a = S[p], b = S[p + k];
            const uint32_t h = ((a | (b << 12)) * LZ2_GOLD) >> (32 - 13);
            slot = lz2_insert_pair(c.T, LZ2_TSLOTS - 1, h, S, a, b, p, k, 0u, &dummy);
        }
        D[p] = (uint16_t)slot;
    }
    __syncthreads();
    int nonhead = 0;
    for (int p = tid; p < n; p += AMBC_BLOCK) {
        const uint32_t slot = D[p];
        uint32_t nm = (uint32_t)p;
        if (slot != 0xFFFFu) nm = c.T[slot] & 0xFFFFu;
        D[p] = (uint16_t)nm;
        if (nm < (uint32_t)p) {
            c.mlen[p] = (uint8_t)(2 * k); c.mpos[p] = (uint16_t)nm;
            atomicOr(&folD[nm >> 5], 1u << (nm & 31));
            nonhead = 1;
        }
    }
    return __syncthreads_or(nonhead);
}

// lengths k+1 .. 2k-1.  S = name_k, D = name_2k, folS = follower bits of level k.
// Returns false on table overflow.
__device__ inline bool lz2_refine(ChunkCtx &c, const uint16_t *S, const uint16_t *D, const uint32_t *folS, int k)
{
    const int n = c.n, tid = threadIdx.x, lane = tid & 31;
    uint16_t *plist = (uint16_t *)(c.T + LZ2_RSLOTS);        // 4096 entries
    uint16_t *islot = plist + LZ2_NMAX;                      // 4096 entries
    volatile int *cnt = c.red + 30;
    volatile int *ovf = c.red + 31;
    if (tid == 0) { *cnt = 0; *ovf = 0; }
    __syncthreads();
    // participants: heads of their 2k-gram whose k-gram occurs elsewhere, with room for k+1 bytes
    const int Pmax = n - k - 1;
    for (int p0 = 0; p0 <= Pmax; p0 += AMBC_BLOCK) {
        const int p = p0 + tid;
        const bool pred = p <= Pmax && D[p] == p && lz2_ns(S, folS, p);
        const uint32_t m = __ballot_sync(FULL_MASK, pred);
        if (m) {
            int base = 0;
            if (lane == 0) base = atomicAdd((int *)cnt, __popc(m));
            base = __shfl_sync(FULL_MASK, base, 0);
            if (pred) plist[base + __popc(m & ((1u << lane) - 1))] = (uint16_t)p;
        }
    }
    __syncthreads();
    const int np = *cnt;
    if (np == 0) return true;
    // pass = (group of nj consecutive lengths) x (one of R key partitions)
    int G = LZ2_PART_TARGET / np, R = 1;
    if (G < 1) { G = 1; while (R * LZ2_PART_TARGET < np) R <<= 1; }
    for (int j0 = 1; j0 < k; j0 += G) {
        const int nj = min(G, k - j0);
        for (int r = 0; r < R; r++) {
            lz2_clear(c.T, LZ2_RSLOTS);
            __syncthreads();
            int overflow = 0;
            for (int jj = 0; jj < nj; jj++) {
                const int j = j0 + jj;
                for (int pi = tid; pi < np; pi += AMBC_BLOCK) {
                    const int p = plist[pi];
                    uint32_t slot = 0xFFFFu;
                    if (p + k + j <= n && lz2_ns(S, folS, p + j)) {
                        const uint32_t a = S[p], b = S[p + j];
                        const uint32_t h = (a | (b << 12) | ((uint32_t)jj << 24)) * LZ2_GOLD;
                        if (((h >> 16) & (uint32_t)(R - 1)) == (uint32_t)r)
                            slot = lz2_insert_pair(c.T, LZ2_RSLOTS - 1, h >> 20, S, a, b, p, j, (uint32_t)jj, &overflow);
                    }
                    islot[jj * np + pi] = (uint16_t)slot;
                }
            }
            if (overflow) *ovf = 1;
            __syncthreads();
            for (int jj = 0; jj < nj; jj++) { // ascending lengths; one thread owns all lengths of a participant
                const int L = k + j0 + jj;
                for (int pi = tid; pi < np; pi += AMBC_BLOCK) {
                    const uint32_t slot = islot[jj * np + pi];
                    if (slot != 0xFFFFu) {
                        const int p = plist[pi];
                        const uint32_t nm = c.T[slot] & 0xFFFFu;
                        if (nm < (uint32_t)p && L > (int)c.mlen[p]) { c.mlen[p] = (uint8_t)L; c.mpos[p] = (uint16_t)nm; }
                    }
                }
            }
            __syncthreads();
            if (*ovf) return false;
        }
    }
    return true;
}

// mlen / mpos for every position of the chunk (c.mlen zeroed by the caller).  n <= LZ2_NMAX.
__device__ inline bool lz2_match_all(ChunkCtx &c)
{
    uint16_t *A = c.nameA, *B = c.nameB;
    uint32_t *f0 = c.fol, *f1 = c.fol + LZ2_NMAX / 32;
    if (c.n < 3) return true;
    const int any4 = lz2_level4(c, A, f0);
    lz2_level3(c, A, B);
    if (!any4) return true;
    const int any8 = lz2_double(c, A, B, f0, f1, 4);
    if (!lz2_refine(c, A, B, f0, 4)) return false;
    if (!any8) return true;
    const int any16 = lz2_double(c, B, A, f1, f0, 8);
    if (!lz2_refine(c, B, A, f1, 8)) return false;
    if (!any16) return true;
    lz2_double(c, A, B, f0, f1, 16);
    return lz2_refine(c, A, B, f0, 16);
}
