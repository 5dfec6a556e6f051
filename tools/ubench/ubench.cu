// dev tool: throughput of the warp / shared-memory primitives the chunk kernels are built from (sm_100a).
// Prints SM cycles per warp-instruction with W warps resident on one SM (one CTA per SM, 148 CTAs).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 512
__device__ __forceinline__ uint32_t rng(uint32_t &s) { s = s * 1664525u + 1013904223u; return s >> 8; }

template <int OP>
__global__ void k(unsigned long long *out, uint32_t *sink, int span)
{
    extern __shared__ uint32_t sm[];
    const int tid = threadIdx.x;
    for (int i = tid; i < 8192; i += blockDim.x) sm[i] = 0;
    __syncthreads();
    uint32_t s = tid * 2654435761u + 12345u, acc = 0;
    const uint32_t mask = (uint32_t)span - 1;
    long long t0 = clock64();
#pragma unroll 4
    for (int it = 0; it < ITERS; it++) {
        uint32_t r = rng(s);
        uint32_t a = r & mask;
        if (OP == 0) acc += sm[a];                               // LDS random
        if (OP == 1) sm[a] = r;                                  // STS random
        if (OP == 2) atomicAdd(&sm[a], 1u);                      // RED/ATOMS add, no return
        if (OP == 3) acc += atomicAdd(&sm[a], 1u);               // ATOMS add with return
        if (OP == 4) acc += atomicCAS(&sm[a], 0u, r);            // ATOMS CAS
        if (OP == 5) acc += atomicMin(&sm[a], r);                // ATOMS min
        if (OP == 6) acc += __match_any_sync(0xffffffffu, r & 0xFF);   // match 8-bit values
        if (OP == 7) acc += __match_any_sync(0xffffffffu, r & 0xFFF);  // match 12-bit values
        if (OP == 8) acc += __reduce_max_sync(0xffffffffu, r);
        if (OP == 9) acc += __ballot_sync(0xffffffffu, r & 1);
        if (OP == 10) acc += __shfl_xor_sync(0xffffffffu, r, 1);
        if (OP == 11) { __syncthreads(); acc += r; }
        if (OP == 12) acc += r;                                   // loop overhead only
        if (OP == 13) acc += ((uint8_t *)sm)[r & (4 * mask + 3)]; // LDS.U8 random
        if (OP == 14) { uint32_t v = sm[a]; sm[a] = v + 1; }      // non-atomic RMW
        if (OP == 15) acc += atomicAdd(&sm[(tid & 31) + 32 * (a & 63)], 1u); // conflict-free banks
        if (OP == 16) acc += __match_any_sync(0xffffffffu, r & 0x3); // match, 4 distinct values
        if (OP == 17) acc += atomicOr(&sm[a], r);
    }
    long long t1 = clock64();
    if (acc == 0x12345678u) sink[0] = acc;
    if (tid == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
}
template <int OP> void run(const char *name, int warps, int span, unsigned long long *d_out, uint32_t *d_sink)
{
    k<OP><<<148, warps * 32, 8192 * 4>>>(d_out, d_sink, span);
    k<OP><<<148, warps * 32, 8192 * 4>>>(d_out, d_sink, span);
    cudaDeviceSynchronize();
    unsigned long long h[148];
    cudaMemcpy(h, d_out, sizeof h, cudaMemcpyDeviceToHost);
    double s = 0; for (int i = 0; i < 148; i++) s += (double)h[i];
    s /= 148;
    printf("%-28s warps=%2d span=%5d  cycles/warp-instr/SM = %7.2f  (per-warp latency %7.1f)\n", name, warps, span,
           s / ((double)ITERS * warps), s / ITERS);
}
int main()
{
    unsigned long long *d_out; uint32_t *d_sink;
    cudaMalloc(&d_out, 148 * 8); cudaMalloc(&d_sink, 64);
    int ws[3] = {4, 16, 32};
    for (int wi = 0; wi < 3; wi++) {
        int w = ws[wi];
        run<12>("loop overhead", w, 8192, d_out, d_sink);
        run<0>("LDS random 32KB", w, 8192, d_out, d_sink);
        run<13>("LDS.U8 random", w, 8192, d_out, d_sink);
        run<1>("STS random", w, 8192, d_out, d_sink);
        run<14>("LDS+STS rmw random", w, 8192, d_out, d_sink);
        run<2>("atomicAdd noret random", w, 8192, d_out, d_sink);
        run<2>("atomicAdd noret 256 bins", w, 256, d_out, d_sink);
        run<3>("atomicAdd ret random", w, 8192, d_out, d_sink);
        run<15>("atomicAdd ret bank-free", w, 8192, d_out, d_sink);
        run<4>("atomicCAS random", w, 8192, d_out, d_sink);
        run<5>("atomicMin random", w, 8192, d_out, d_sink);
        run<17>("atomicOr random", w, 8192, d_out, d_sink);
        run<6>("match_any 8-bit", w, 8192, d_out, d_sink);
        run<7>("match_any 12-bit", w, 8192, d_out, d_sink);
        run<16>("match_any 2-bit", w, 8192, d_out, d_sink);
        run<8>("redux max", w, 8192, d_out, d_sink);
        run<9>("ballot", w, 8192, d_out, d_sink);
        run<10>("shfl xor", w, 8192, d_out, d_sink);
        run<11>("syncthreads", w, 8192, d_out, d_sink);
    }
    return 0;
}
