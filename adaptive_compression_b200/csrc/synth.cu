// synth.cu -- counter-based synthetic corpus ("mixed CSV / log / binary" of BASELINE.json).
// Every 16-byte unit is a pure function of (seed, byte offset), so any range can be generated
// on any GPU, and tests/synth_ref.py regenerates the same bytes with numpy.
//
// The corpus is a sequence of 64 KiB segments; segment s has kind
//   kinds[mix(seed ^ (s+1)*GOLD) % count]   over the kinds enabled in kind_mask:
//   0 csv      32-byte rows  "0001234,512.07,c05,2026-03-17,B\n"
//   1 log      64-byte lines "2026-10-18T03:25:41 INFO  svc07 GET /v1/items/01234 200 00123ms\n"
//   2 runs     256-byte blocks: la bytes of A then 256-la bytes of B
//   3 lowcard  bytes from a 12-symbol skewed alphabet
//   4 binrec   16-byte little-endian records (counter, small fields, flags)
//   5 random   uniform bytes
//   6 text     8-byte dictionary words
#include "ambc_internal.h"

#define SEG_BYTES 65536ull
#define GOLD 0x9E3779B97F4A7C15ull
#define MIXB 0xD1B54A32D192ED03ull

__host__ __device__ inline unsigned long long mix64(unsigned long long x)
{
    x += GOLD;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

__constant__ char c_csv_lit[33] = "0000000,000.00,c00,2026-00-00,A\n";
// field id per byte (0xFF = literal), and divisor
__constant__ unsigned char c_csv_fid[32] = {0,0,0,0,0,0,0,255, 1,1,1,255, 2,2,255,255, 3,3,255,255,255,255,255,255, 4,4,255, 5,5,255, 6,255};
__constant__ unsigned int c_csv_div[32] = {1000000,100000,10000,1000,100,10,1,0, 100,10,1,0, 10,1,0,0, 10,1,0,0,0,0,0,0, 10,1,0, 10,1,0, 0,0};
__constant__ char c_log_lit[65] = "2026-10-18T00:00:00 LLLLL svc00 GET /v1/items/00000 000 00000ms\n";
__constant__ unsigned char c_log_fid[64] = {
    255,255,255,255,255,255,255,255,255,255,255, 0,0,255, 1,1,255, 2,2,255, 3,3,3,3,3,255, 255,255,255, 4,4,255,
    255,255,255,255, 255,255,255,255,255,255,255,255,255,255, 5,5,5,5,5,255, 6,6,6,255, 7,7,7,7,7,255,255,255};
__constant__ unsigned int c_log_div[64] = {
    0,0,0,0,0,0,0,0,0,0,0, 10,1,0, 10,1,0, 10,1,0, 0,1,2,3,4,0, 0,0,0, 10,1,0,
    0,0,0,0, 0,0,0,0,0,0,0,0,0,0, 10000,1000,100,10,1,0, 100,10,1,0, 10000,1000,100,10,1,0,0,0};
__constant__ char c_levels[4][6] = {"INFO ", "WARN ", "ERROR", "DEBUG"};
__constant__ unsigned short c_status[8] = {200, 200, 200, 404, 500, 301, 200, 200};
__constant__ unsigned char c_skew[16] = {0, 0, 0, 0, 1, 1, 1, 2, 2, 3, 4, 5, 6, 7, 8, 9 + 2};
__constant__ char c_words[64][9] = {
    "the     ", "quick   ", "brown   ", "fox     ", "jumps   ", "over    ", "lazy    ", "dog     ",
    "adaptive", "marker  ", "based   ", "compress", "chunk   ", "method  ", "huffman ", "dictiona",
    "delta   ", "run     ", "length  ", "encoding", "stream  ", "header  ", "package ", "error   ",
    "warning ", "info    ", "debug   ", "request ", "response", "latency ", "status  ", "user    ",
    "session ", "and     ", "of      ", "to      ", "in      ", "is      ", "that    ", "for     ",
    "with    ", "as      ", "on      ", "be      ", "at      ", "by      ", "this    ", "have    ",
    "from    ", "or      ", "one     ", "had     ", "not     ", "but     ", "what    ", "all     ",
    "were    ", "when    ", "we      ", "there   ", "can     ", "an      ", "your    ", "which.\n "};

// 16 bytes of the corpus at 16-aligned offset `off`
__device__ void synth_unit(unsigned long long seed, unsigned kind, unsigned long long seg, unsigned off,
                           unsigned char *b)
{
    switch (kind) {
    case 0: { // csv, 32-byte rows
        unsigned rec = off >> 5, j0 = off & 31;
        unsigned long long h = mix64(seed + seg * GOLD + (unsigned long long)rec * MIXB);
        unsigned f[7];
        f[0] = (unsigned)((seg * 2048ull + rec) % 10000000ull);
        f[1] = (unsigned)(h % 1000ull);
        f[2] = (unsigned)((h >> 10) % 100ull);
        f[3] = (unsigned)((h >> 20) % 17ull);
        f[4] = 1 + (unsigned)((h >> 28) % 12ull);
        f[5] = 1 + (unsigned)((h >> 34) % 28ull);
        f[6] = (unsigned)((h >> 40) % 4ull);
        for (int k = 0; k < 16; k++) {
            unsigned j = j0 + k, fid = c_csv_fid[j];
            unsigned char ch = (unsigned char)c_csv_lit[j];
            if (fid != 255) ch = fid == 6 ? (unsigned char)('A' + f[6]) : (unsigned char)('0' + (f[fid] / c_csv_div[j]) % 10);
            b[k] = ch;
        }
        break;
    }
    case 1: { // log, 64-byte lines
        unsigned rec = off >> 6, j0 = off & 63;
        unsigned long long h = mix64(seed + seg * GOLD + (unsigned long long)rec * MIXB);
        unsigned long long t = seg * 1024ull + rec;
        unsigned f[8];
        f[0] = (unsigned)((t / 3600ull) % 24ull);
        f[1] = (unsigned)((t / 60ull) % 60ull);
        f[2] = (unsigned)(t % 60ull);
        f[3] = (unsigned)(h % 8ull); f[3] = f[3] < 5 ? 0 : f[3] - 4; // INFO x5, WARN, ERROR, DEBUG
        f[4] = (unsigned)((h >> 8) % 12ull);
        f[5] = (unsigned)((h >> 16) % 50000ull);
        f[6] = c_status[(h >> 36) % 8ull];
        f[7] = (unsigned)((h >> 40) % 100000ull);
        for (int k = 0; k < 16; k++) {
            unsigned j = j0 + k, fid = c_log_fid[j];
            unsigned char ch = (unsigned char)c_log_lit[j];
            if (fid == 3) ch = (unsigned char)c_levels[f[3]][c_log_div[j]];
            else if (fid != 255) ch = (unsigned char)('0' + (f[fid] / c_log_div[j]) % 10);
            b[k] = ch;
        }
        break;
    }
    case 2: { // runs, 256-byte blocks
        unsigned blk = off >> 8, j0 = off & 255;
        unsigned long long h = mix64(seed + seg * GOLD + (unsigned long long)blk * MIXB);
        unsigned la = (unsigned)(h % 257ull);
        unsigned char A = (unsigned char)(h >> 16), B = (unsigned char)(h >> 24);
        for (int k = 0; k < 16; k++) b[k] = (j0 + k) < la ? A : B;
        break;
    }
    case 3: { // lowcard: 4 bits of hash per byte through a skew table, per-segment alphabet
        unsigned u = off >> 4;
        unsigned long long h = mix64(seed + seg * GOLD + (unsigned long long)u * MIXB);
        unsigned long long sa = mix64(seed ^ (seg * MIXB));
        for (int k = 0; k < 16; k++) {
            unsigned s = c_skew[(h >> (4 * k)) & 15];
            b[k] = (unsigned char)(48 + ((sa >> (4 * (s % 12))) & 15) + 6 * s);
        }
        break;
    }
    case 4: { // binrec, 16-byte records
        unsigned rec = off >> 4;
        unsigned long long h = mix64(seed + seg * GOLD + (unsigned long long)rec * MIXB);
        unsigned cnt = (unsigned)(seg * 4096ull + rec);
        unsigned ts = rec * 10u + (unsigned)(h % 7ull);
        b[0] = cnt; b[1] = cnt >> 8; b[2] = cnt >> 16; b[3] = cnt >> 24;
        b[4] = (unsigned char)(h % 40ull); b[5] = 0;
        b[6] = (unsigned char)((h >> 8) % 3ull); b[7] = 0;
        b[8] = ts; b[9] = ts >> 8; b[10] = ts >> 16; b[11] = ts >> 24;
        b[12] = (unsigned char)((h >> 20) & 1); b[13] = 0; b[14] = 0xFF; b[15] = 0;
        break;
    }
    case 5: { // random
        unsigned u = off >> 4;
        unsigned long long h0 = mix64(seed + seg * GOLD + (unsigned long long)(2 * u) * MIXB);
        unsigned long long h1 = mix64(seed + seg * GOLD + (unsigned long long)(2 * u + 1) * MIXB);
        for (int k = 0; k < 8; k++) { b[k] = (unsigned char)(h0 >> (8 * k)); b[8 + k] = (unsigned char)(h1 >> (8 * k)); }
        break;
    }
    default: { // text: 8-byte words, 2 per unit
        unsigned u = off >> 4;
        unsigned long long h = mix64(seed + seg * GOLD + (unsigned long long)u * MIXB);
        for (int w = 0; w < 2; w++) {
            unsigned idx = (unsigned)((h >> (6 * w)) & 63);
            // skew towards the first 16 words
            if ((h >> (20 + w)) & 1) idx &= 15;
            for (int k = 0; k < 8; k++) b[8 * w + k] = (unsigned char)c_words[idx][k];
        }
        break;
    }
    }
}

__device__ __forceinline__ unsigned seg_kind(unsigned long long seed, unsigned long long seg, unsigned kind_mask)
{
    unsigned cnt = __popc(kind_mask & 0x7F);
    unsigned pick = (unsigned)(mix64(seed ^ ((seg + 1) * GOLD)) % (unsigned long long)cnt);
    unsigned m = kind_mask & 0x7F;
    for (unsigned k = 0; k < pick; k++) m &= m - 1;
    return __ffs(m) - 1;
}

__global__ void __launch_bounds__(256)
k_synth(unsigned char *__restrict__ out, unsigned long long offset, unsigned long long n, unsigned long long seed,
        unsigned kind_mask)
{
    // unit u covers corpus bytes [16u, 16u+16); generate every unit overlapping [offset, offset+n)
    const unsigned long long u0 = offset >> 4, u1 = (offset + n + 15) >> 4;
    for (unsigned long long u = u0 + (unsigned long long)blockIdx.x * 256 + threadIdx.x; u < u1;
         u += (unsigned long long)gridDim.x * 256) {
        unsigned long long pos = u << 4;
        unsigned long long seg = pos / SEG_BYTES;
        unsigned off = (unsigned)(pos % SEG_BYTES);
        unsigned char b[16];
        synth_unit(seed, seg_kind(seed, seg, kind_mask), seg, off, b);
        if (pos >= offset && pos + 16 <= offset + n && (((uintptr_t)(out + (pos - offset))) & 15) == 0) {
            uint4 v;
            v.x = b[0] | (b[1] << 8) | (b[2] << 16) | ((unsigned)b[3] << 24);
            v.y = b[4] | (b[5] << 8) | (b[6] << 16) | ((unsigned)b[7] << 24);
            v.z = b[8] | (b[9] << 8) | (b[10] << 16) | ((unsigned)b[11] << 24);
            v.w = b[12] | (b[13] << 8) | (b[14] << 16) | ((unsigned)b[15] << 24);
            *(uint4 *)(out + (pos - offset)) = v;
        } else {
            for (int k = 0; k < 16; k++) {
                unsigned long long p = pos + k;
                if (p >= offset && p < offset + n) out[p - offset] = b[k];
            }
        }
    }
}

extern "C" int ambc_synth_dev(void *out_dev, uint64_t offset, uint64_t n, uint64_t seed, uint32_t kind_mask,
                              void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n == 0) return AMBC_OK;
    if (!out_dev || !(kind_mask & 0x7F)) return ambc_fail(AMBC_E_ARG, "ambc_synth_dev: bad argument");
    uint64_t units = ((offset + n + 15) >> 4) - (offset >> 4);
    unsigned grid = (unsigned)min<uint64_t>((units + 255) / 256, 148 * 16);
    k_synth<<<grid, 256, 0, stream>>>((unsigned char *)out_dev, offset, n, seed, kind_mask);
    ambc_count_launch();
    CUDA_TRY(cudaGetLastError());
    return AMBC_OK;
}
