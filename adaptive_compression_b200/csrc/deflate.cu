// deflate.cu -- batch kernels of method id 5 (DeflateCompression plug-in, advanced_compression.py:71-107):
// one warp per item, lane 0 runs the serial coder of deflate.cuh, the other lanes stage, hash-clear and pad.
#include "ambc_internal.h"
#include "deflate.cuh"

#define DFL_WARPS 4
#define DFL_NMAX 8192

struct DflWarp {
    uint8_t data[DFL_NMAX + 16];
    uint16_t head[1 << DEF_HB];
};

__global__ void __launch_bounds__(DFL_WARPS * 32)
k_deflate_batch(const uint8_t *__restrict__ in, const uint64_t *__restrict__ in_off, uint32_t n_items,
                uint8_t *__restrict__ out, uint64_t out_stride, int32_t *__restrict__ out_len)
{
    extern __shared__ uint4 smem4[];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    DflWarp *s = (DflWarp *)smem4 + w;
    for (uint32_t i = blockIdx.x * DFL_WARPS + w; i < n_items; i += gridDim.x * DFL_WARPS) {
        const uint64_t a = in_off[i];
        const int n = (int)(in_off[i + 1] - a);
        int len = 0;
        if (n > 0 && n <= DFL_NMAX) { // DeflateCompression.compress: b'' for empty data (:78-79)
            unsigned long long s1 = 0, s2 = 0; // Adler-32 (RFC 1950): s1 = 1 + sum d, s2 = n + sum (n - k) d_k
            for (int k = lane; k < n; k += 32) {
                const uint32_t v = __ldg(in + a + k);
                s->data[k] = (uint8_t)v;
                s1 += v;
                s2 += (unsigned long long)(n - k) * v;
            }
            for (int k = lane; k < (1 << DEF_HB); k += 32) s->head[k] = 0xFFFF;
            for (int d = 16; d > 0; d >>= 1) { s1 += __shfl_xor_sync(FULL_MASK, s1, d); s2 += __shfl_xor_sync(FULL_MASK, s2, d); }
            const uint32_t adler = ((uint32_t)((s2 + (unsigned long long)n) % ADLER_MOD) << 16) | (uint32_t)((s1 + 1) % ADLER_MOD);
            __syncwarp();
            if (lane == 0) {
                const int cap = out_stride > 0x7fffffffull ? 0x7fffffff : (int)out_stride;
                len = deflate_fixed(s->data, n, out + (uint64_t)i * out_stride, cap, s->head, adler);
                if (len > cap) len = AMBC_E_CAPACITY;
            }
        } else if (n > DFL_NMAX) len = AMBC_E_TOO_LARGE;
        if (lane == 0) out_len[i] = len;
        __syncwarp();
    }
}

__global__ void __launch_bounds__(DFL_WARPS * 32)
k_inflate_batch(const uint8_t *__restrict__ in, const uint64_t *__restrict__ in_off, const uint32_t *__restrict__ orig_len,
                uint32_t n_items, uint8_t *__restrict__ out, uint64_t out_stride, int32_t *__restrict__ out_len)
{
    __shared__ InfCode codes[DFL_WARPS][2];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (uint32_t i = blockIdx.x * DFL_WARPS + w; i < n_items; i += gridDim.x * DFL_WARPS) {
        const uint64_t a = in_off[i];
        const long comp = (long)(in_off[i + 1] - a);
        const uint32_t orig = orig_len[i];
        uint8_t *dst = out + (uint64_t)i * out_stride;
        const uint32_t want = (uint64_t)orig < out_stride ? orig : (uint32_t)out_stride;
        int result = 0;
        if (comp > 0) { // DeflateCompression.decompress: b'' for empty data; zlib error -> zeros (:84-97)
            long got = 0;
            // (32 KiB of window behind the item's original_length bytes when the stride has room for it)
            const bool ring = out_stride >= (uint64_t)want + 32768u;
            if (lane == 0) got = inflate_zlib(in + a, comp, dst, (long)want, ring, &codes[w][0], &codes[w][1]);
            got = __shfl_sync(FULL_MASK, (long long)got, 0);
            const uint32_t good = got < 0 ? 0u : (got < (long)want ? (uint32_t)got : want);
            __syncwarp();
            for (uint32_t k = good + lane; k < want; k += 32) dst[k] = 0; // pad (or all zeros after an error)
            result = (int)orig;
        }
        if (lane == 0) out_len[i] = result;
        __syncwarp();
    }
}

int ambc_deflate_encode_batch(const void *in_dev, const uint64_t *in_off_dev, uint32_t n_items, void *out_dev,
                              uint64_t out_stride, int32_t *out_len_dev, cudaStream_t stream)
{
    const size_t smem = DFL_WARPS * sizeof(DflWarp);
    CUDA_TRY(cudaFuncSetAttribute(k_deflate_batch, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_deflate_batch<<<(n_items + DFL_WARPS - 1) / DFL_WARPS, DFL_WARPS * 32, smem, stream>>>(
        (const uint8_t *)in_dev, in_off_dev, n_items, (uint8_t *)out_dev, out_stride, out_len_dev);
    ambc_count_launch();
    CUDA_TRY(cudaGetLastError());
    return AMBC_OK;
}

int ambc_inflate_batch(const void *in_dev, const uint64_t *in_off_dev, const uint32_t *orig_len_dev, uint32_t n_items,
                       void *out_dev, uint64_t out_stride, int32_t *out_len_dev, cudaStream_t stream)
{
    k_inflate_batch<<<(n_items + DFL_WARPS - 1) / DFL_WARPS, DFL_WARPS * 32, 0, stream>>>(
        (const uint8_t *)in_dev, in_off_dev, orig_len_dev, n_items, (uint8_t *)out_dev, out_stride, out_len_dev);
    ambc_count_launch();
    CUDA_TRY(cudaGetLastError());
    return AMBC_OK;
}
