// marker.cu -- MarkerFinder.find_marker on the GPU (marker_finder.py:22-123):
// the shortest bit string, and among those the smallest value, that does not occur in the
// MSB-first bit stream of the data.
//
//   k_marker_flags   presence of every L-bit window.  L <= 16: per-CTA bitmap in shared memory
//                    (test-before-set shared atomics over coalesced 16-byte loads), flushed to
//                    byte flags in global memory; L > 16: byte flags in global memory (L2), behind a per-CTA
//                    direct-mapped table of the values this CTA looked up last (48 KB of shared memory): a
//                    window value found there costs one shared load; anything else goes to the test-before-set
//                    on the global flag (round 1 went there for every window: 8 L2 sectors per input byte, the
//                    bound of that level.  Blind stores instead of test-before-set were measured slower).
//   k_marker_level   presence at length l-1 from length l (prefix OR + the stream's last window)
//                    and the smallest absent value per length.
// Byte flags (not packed bits) so that shards on several GPUs merge with one NCCL
// max-allreduce (NCCL has no bitwise OR).
#include "ambc_internal.h"
#include "common.cuh"

#define MK_THREADS 256

__device__ __forceinline__ uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

#define MK_CACHE 12288 // entries of the per-CTA "flagged lately" table (L > 16)
template <bool SMEM>
__global__ void __launch_bounds__(MK_THREADS)
k_marker_flags(const uint8_t *__restrict__ in, uint64_t n, uint32_t L, uint8_t *__restrict__ flags)
{
    extern __shared__ uint32_t bm[]; // SMEM: 2^L bits; else MK_CACHE values
    const uint32_t nwords = SMEM ? max(1u, (1u << L) >> 5) : 0;
    if (SMEM) {
        for (uint32_t i = threadIdx.x; i < nwords; i += MK_THREADS) bm[i] = 0;
        __syncthreads();
    } else {
        for (uint32_t i = threadIdx.x; i < MK_CACHE; i += MK_THREADS) bm[i] = 0xFFFFFFFFu; // (no L-bit value, L <= 31... see below)
        __syncthreads();
    }
    const uint64_t nbits = n * 8;
    const uint64_t ngroups = (n + 15) / 16; // 16 bytes = 128 window starts per thread step
    const bool aligned = (((uintptr_t)in) & 15) == 0;
    for (uint64_t g = (uint64_t)blockIdx.x * MK_THREADS + threadIdx.x; g < ngroups; g += (uint64_t)gridDim.x * MK_THREADS) {
        const uint64_t b0 = g * 16;
        uint32_t w[6]; // bytes b0 .. b0+24, zero beyond n
        if (aligned && b0 + 32 <= n) {
            uint4 a = __ldg((const uint4 *)(in + b0));
            uint2 c = __ldg((const uint2 *)(in + b0 + 16));
            w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = c.x; w[5] = c.y;
        } else {
#pragma unroll
            for (int k = 0; k < 6; k++) {
                uint32_t v = 0;
                for (int t = 0; t < 4; t++) {
                    uint64_t idx = b0 + 4 * k + t;
                    if (idx < n) v |= (uint32_t)in[idx] << (8 * t);
                }
                w[k] = v;
            }
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            // windows starting in word k: bits [32k+t, 32k+t+L) of the big-endian string
            const unsigned long long hi = ((unsigned long long)bswap32(w[k]) << 32) | bswap32(w[k + 1]);
            const uint64_t sbase = b0 * 8 + 32 * k;
#pragma unroll 8
            for (int t = 0; t < 32; t++) {
                if (sbase + t + L <= nbits) {
                    uint32_t v = (uint32_t)((hi << t) >> (64 - L));
                    if (SMEM) {
                        uint32_t bit = 1u << (v & 31);
                        if (!(bm[v >> 5] & bit)) atomicOr(&bm[v >> 5], bit);
                    } else {
                        // (racy on purpose: a lost update of the table only costs a redundant global store;
                        // for L == 32 the value 0xFFFFFFFF is never "found" in a fresh table either way, it is stored)
                        const uint32_t slot = (v * 2654435761u) >> 18; // 14 bits
                        const uint32_t sl = slot < MK_CACHE ? slot : slot - MK_CACHE / 2;
                        if (bm[sl] != v || v == 0xFFFFFFFFu) { bm[sl] = v; if (!flags[v]) flags[v] = 1; }
                    }
                }
            }
        }
    }
    if (SMEM) {
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < nwords; i += MK_THREADS) {
            uint32_t x = bm[i];
            while (x) {
                int b = __ffs(x) - 1;
                x &= x - 1;
                uint32_t v = (i << 5) + b;
                if (v < (1u << L)) flags[v] = 1;
            }
        }
    }
}

// windows that start inside the carried-in bits of the previous shard
__global__ void k_marker_carry(const uint8_t *__restrict__ in, uint64_t n, uint32_t L, uint64_t carry,
                               uint32_t carry_bits, uint8_t *__restrict__ flags)
{
    if (threadIdx.x || blockIdx.x) return;
    // first min(n,8) bytes of the shard as a big-endian 64-bit string
    unsigned long long head = 0;
    for (int k = 0; k < 8; k++) head = (head << 8) | (unsigned long long)((uint64_t)k < n ? in[k] : 0);
    for (uint32_t u = 0; u < carry_bits; u++) {
        uint32_t from_carry = carry_bits - u; // bits taken from the carry (its low from_carry bits)
        if (from_carry >= L) continue;
        uint32_t from_data = L - from_carry;
        if ((uint64_t)from_data > n * 8) continue;
        unsigned long long cpart = carry & ((1ull << from_carry) - 1);
        unsigned long long v = (cpart << from_data) | (head >> (64 - from_data));
        flags[v] = 1;
    }
}

extern "C" int ambc_marker_flags_dev(const void *in_dev, uint64_t n, uint32_t L, uint64_t carry,
                                     uint32_t carry_bits, uint8_t *flags_dev, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (L < 1 || L > 32 || !flags_dev || carry_bits >= L) return ambc_fail(AMBC_E_ARG, "ambc_marker_flags_dev: bad argument");
    if (n == 0) return AMBC_OK;
    if (!in_dev) return ambc_fail(AMBC_E_ARG, "null input");
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint64_t ngroups = (n + 15) / 16;
    unsigned grid = (unsigned)min<uint64_t>((ngroups + MK_THREADS - 1) / MK_THREADS, (uint64_t)sms * 8);
    if (L <= 16) {
        size_t smem = max<size_t>(4, ((size_t)1 << L) / 8);
        k_marker_flags<true><<<grid, MK_THREADS, smem, stream>>>((const uint8_t *)in_dev, n, L, flags_dev);
    } else {
        cudaFuncSetAttribute(k_marker_flags<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MK_CACHE * 4);
        k_marker_flags<false><<<grid, MK_THREADS, MK_CACHE * 4, stream>>>((const uint8_t *)in_dev, n, L, flags_dev);
    }
    ambc_count_launch();
    if (carry_bits) {
        k_marker_carry<<<1, 32, 0, stream>>>((const uint8_t *)in_dev, n, L, carry, carry_bits, flags_dev);
        ambc_count_launch();
    }
    CUDA_TRY(cudaGetLastError());
    return AMBC_OK;
}

// cur = presence bytes at length l (2^l entries).  Writes presence at length l-1 into nxt and
// the smallest absent value of length l into minabs[l].
__global__ void __launch_bounds__(256)
k_marker_level(const uint8_t *__restrict__ cur, uint32_t l, uint8_t *__restrict__ nxt, uint64_t tail_value,
               int has_tail, unsigned long long *minabs)
{
    const uint64_t half = 1ull << (l - 1);
    unsigned long long mymin = ~0ull;
    for (uint64_t v = (uint64_t)blockIdx.x * 256 + threadIdx.x; v < half; v += (uint64_t)gridDim.x * 256) {
        uint8_t a = cur[2 * v], b = cur[2 * v + 1];
        if (!a) mymin = min(mymin, (unsigned long long)(2 * v));
        else if (!b) mymin = min(mymin, (unsigned long long)(2 * v + 1));
        if (nxt) nxt[v] = (a | b | (has_tail && v == tail_value)) ? 1 : 0;
    }
    for (int d = 16; d > 0; d >>= 1) mymin = min(mymin, __shfl_xor_sync(FULL_MASK, mymin, d));
    if ((threadIdx.x & 31) == 0 && mymin != ~0ull) atomicMin(&minabs[l], mymin);
}

extern "C" int ambc_marker_pick_dev(const uint8_t *flags_dev, uint32_t L, uint32_t max_len, uint64_t total_bits,
                                    uint64_t tail, uint32_t tail_bits, uint32_t *out_len, uint64_t *out_value,
                                    void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (L < 1 || L > 32 || !flags_dev || !out_len || !out_value) return ambc_fail(AMBC_E_ARG, "ambc_marker_pick_dev: bad argument");
    uint8_t *scratch = nullptr;
    unsigned long long *minabs = nullptr;
    uint64_t sbytes = (1ull << L) + 64; // levels L-1 .. 0 need 2^(L-1) + ... + 1 < 2^L bytes
    CUDA_TRY(cudaMalloc(&scratch, sbytes));
    cudaError_t e = cudaMalloc(&minabs, 40 * sizeof(unsigned long long));
    if (e != cudaSuccess) { cudaFree(scratch); return ambc_fail(AMBC_E_CUDA, "cudaMalloc: %s", cudaGetErrorString(e)); }
    cudaMemsetAsync(minabs, 0xFF, 40 * sizeof(unsigned long long), stream);
    const uint8_t *cur = flags_dev;
    uint8_t *dst = scratch;
    for (uint32_t l = L; l >= 1; l--) {
        uint64_t half = 1ull << (l - 1);
        // last window of length l-1 of the whole stream exists when total_bits >= l-1 >= 1
        int has_tail = (l - 1 >= 1) && total_bits >= (uint64_t)(l - 1) && tail_bits >= l - 1;
        uint64_t tv = has_tail ? (tail & ((1ull << (l - 1)) - 1)) : 0;
        unsigned grid = (unsigned)min<uint64_t>((half + 255) / 256, 148 * 8);
        k_marker_level<<<grid, 256, 0, stream>>>(cur, l, l > 1 ? dst : nullptr, tv, has_tail, minabs);
        ambc_count_launch();
        cur = dst;
        dst += half;
    }
    unsigned long long h[40];
    e = cudaMemcpyAsync(h, minabs, sizeof h, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(scratch);
    cudaFree(minabs);
    if (e != cudaSuccess) return ambc_fail(AMBC_E_CUDA, "marker pick: %s", cudaGetErrorString(e));
    uint32_t top = L < max_len ? L : max_len;
    for (uint32_t l = 1; l <= top; l++) {
        if (h[l] != ~0ull) { *out_len = l; *out_value = h[l]; return AMBC_OK; }
    }
    *out_len = 0; *out_value = 0;
    return ambc_fail(AMBC_E_NO_MARKER, "Could not find a marker of length <= %u bits", top);
}

extern "C" int ambc_find_marker_dev(const void *in_dev, uint64_t n, uint32_t max_len, uint32_t *out_len,
                                    uint64_t *out_value, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (max_len < 1) return ambc_fail(AMBC_E_NO_MARKER, "Could not find a marker of length <= %u bits", max_len);
    if (max_len > 32) max_len = 32;
    const uint64_t total_bits = n * 8;
    // last (up to 31) bits of the stream, for the shorter-length derivation
    uint8_t last[4] = {0, 0, 0, 0};
    uint64_t nl = n < 4 ? n : 4;
    if (nl) CUDA_TRY(cudaMemcpyAsync(last + (4 - nl), (const uint8_t *)in_dev + (n - nl), nl, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    uint64_t tail32 = ((uint64_t)last[0] << 24) | ((uint64_t)last[1] << 16) | ((uint64_t)last[2] << 8) | last[3];
    uint32_t tail_have = (uint32_t)min<uint64_t>(31, total_bits);
    uint64_t tail = tail32 & ((1ull << tail_have) - 1);
    const uint32_t levels[3] = {16, 24, 32};
    for (int s = 0; s < 3; s++) {
        uint32_t L = levels[s] < max_len ? levels[s] : max_len;
        uint8_t *flags = nullptr;
        CUDA_TRY(cudaMalloc(&flags, (1ull << L) + 64));
        cudaMemsetAsync(flags, 0, 1ull << L, stream);
        int rc = ambc_marker_flags_dev(in_dev, n, L, 0, 0, flags, stream);
        if (rc == AMBC_OK) rc = ambc_marker_pick_dev(flags, L, max_len, total_bits, tail, tail_have, out_len, out_value, stream);
        cudaStreamSynchronize(stream);
        cudaFree(flags);
        if (rc != AMBC_E_NO_MARKER) return rc;
        if (L == max_len) return rc;
    }
    return ambc_fail(AMBC_E_NO_MARKER, "Could not find a marker of length <= %u bits", max_len);
}
