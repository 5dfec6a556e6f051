/*
 * include/ambc.h -- C ABI of libambc.so, the B200 (sm_100a) implementation of
 * adaptive-compression's chunked encode / select / decode hot path.
 *
 * This is the drop-in boundary: the reference is pure Python with no FFI of
 * its own, so these are the entry points a maintainer binds with ctypes from
 * adaptive_compressor.py / compression_methods.py / marker_finder.py (see
 * INTEGRATION.md for the stubs).  Each entry point cites the reference
 * function it replaces (file:line relative to the reference repo).
 *
 * Conventions
 *   - plain C symbols, plain pointers and sizes; no torch / C++ types.
 *   - return value: 0 = ok, <0 = error code below; text via ambc_last_error().
 *   - "_dev" pointers are CUDA device pointers owned by the caller (for
 *     example torch.empty(..., device="cuda").data_ptr()); `stream` is a
 *     cudaStream_t passed as void* (NULL = default stream).
 *   - there is NO CPU fallback: every call fails with AMBC_E_CUDA when no
 *     device / kernel image is available.
 */
#ifndef AMBC_H
#define AMBC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AMBC_OK 0
#define AMBC_E_CUDA (-1)      /* CUDA runtime error (no device, launch failure ...) */
#define AMBC_E_ARG (-2)       /* bad argument                                        */
#define AMBC_E_CAPACITY (-3)  /* output / workspace too small                        */
#define AMBC_E_MARKER (-4)    /* "Marker mismatch in chunk header."  (adaptive_compressor.py:406-407) */
#define AMBC_E_TOO_LARGE (-5) /* chunk larger than AMBC_MAX_CODEC_CHUNK for a codec call */
#define AMBC_E_NO_MARKER (-6) /* no absent bit string up to max length (marker_finder.py:123) */

/* method ids = package types (adaptive_compressor.py:97-110) */
#define AMBC_RLE 1
#define AMBC_DICT 2
#define AMBC_HUFFMAN 3
#define AMBC_DELTA 4
#define AMBC_RAW 255
#define AMBC_METHOD_BIT(id) (1u << (id))
#define AMBC_NATIVE_MASK (AMBC_METHOD_BIT(1) | AMBC_METHOD_BIT(2) | AMBC_METHOD_BIT(3) | AMBC_METHOD_BIT(4))

/* largest chunk any native method is eligible for (adaptive_compressor.py:114-127) */
#define AMBC_MAX_CODEC_CHUNK 8192

/* flags of ambc_compress_* */
#define AMBC_F_PER_CHUNK_RAW 1u /* labelled extension: a losing chunk becomes its own type-255
                                   package instead of the reference's "rest of file raw" rule
                                   (adaptive_compressor.py:586-590).  Off = reference behaviour. */

/* codec result codes in out_len[] of the *_batch calls (the Python exception the
 * reference raises for that input) */
#define AMBC_CODEC_INDEX_ERROR (-1)
#define AMBC_CODEC_VALUE_ERROR (-2)

const char *ambc_last_error(void);
int ambc_version(void);
/* number of CUDA devices, or AMBC_E_CUDA */
int ambc_device_count(void);
/* kernels launched by this library since load (bench.py's gpu_launches) */
uint64_t ambc_launch_count(void);

/* ------------------------------------------------------------------ */
/* compress: AdaptiveCompressor._adaptive_compress                      */
/*   adaptive_compressor.py:363-394 (chunk loop), :537-590 (trial +      */
/*   argmin + rest-of-file-raw rule), :631-700 (re-compress + benefit    */
/*   check + framing), :595-621 (package / END framing)                  */
/* ------------------------------------------------------------------ */

typedef struct {
    uint64_t body_len;      /* bytes written to out (all packages + 16-byte END)            */
    uint64_t n_chunks;      /* grid chunks = ceil(n / chunk)                                 */
    int64_t first_raw;      /* first grid chunk with no winning method, -1 if none           */
    uint64_t n_packages;    /* data packages written                                         */
    uint64_t map_type_off;  /* byte offset in the workspace of uint8  type[n_chunks]         */
    uint64_t map_comp_off;  /* byte offset in the workspace of uint32 comp_len[n_chunks]     */
    uint64_t payload_bytes; /* sum of comp_len over compressed (non-raw) packages            */
    uint64_t usage[5];      /* packages per method id 1..4, [0] = raw packages               */
} ambc_compress_result;

/* bytes of device workspace ambc_compress_dev needs for an input of n bytes */
uint64_t ambc_compress_workspace_bytes(uint64_t n, uint32_t chunk);
/* upper bound of the body length (out_cap to allocate) */
uint64_t ambc_compress_bound(uint64_t n, uint32_t chunk, uint32_t marker_bytes);

/*
 * Device-resident input -> device-resident body (packages + END).  `chunk` is the
 * single CHUNK_SIZE_CANDIDATES entry (adaptive_compressor.py:61, :548).
 * `method_mask` = enabled method ids (AMBC_METHOD_BIT), trial order 1,2,3,4 as in
 * the reference's list (adaptive_compressor.py:131-134).  `marker` =
 * marker_bytes_aligned (adaptive_compressor.py:196-219), 1..4 bytes.
 * Synchronises `stream` before returning and fills *res.
 */
int ambc_compress_dev(const void *in_dev, uint64_t n, uint32_t chunk, uint32_t method_mask, uint32_t flags,
                      const uint8_t *marker, uint32_t marker_bytes, void *out_dev, uint64_t out_cap,
                      void *work_dev, uint64_t work_bytes, ambc_compress_result *res, void *stream);

/*
 * Same with HOST buffers (pinned memory recommended): H2D copy, kernels, D2H copy
 * of the body and of the per-chunk method map.  map_type / map_comp may be NULL.
 * Device buffers are cached inside the library between calls.
 */
int ambc_compress_host(const void *in_host, uint64_t n, uint32_t chunk, uint32_t method_mask, uint32_t flags,
                       const uint8_t *marker, uint32_t marker_bytes, void *out_host, uint64_t out_cap,
                       uint8_t *map_type, uint32_t *map_comp, ambc_compress_result *res);

/*
 * Multi-candidate ("dynamic chunk size") mode: the reference's default
 * CHUNK_SIZE_CANDIDATES = [131072 .. 1024] (adaptive_compressor.py:61-62) or any
 * strictly descending list whose gcd is a multiple of 16 and >= 256.  At every
 * position each candidate (clamped to the remaining bytes) is tried, the smallest
 * ratio (len + overhead) / size wins, larger candidates win ties
 * (adaptive_compressor.py:548-584); no winner -> the rest is one raw package.
 * map_out (optional, map_cap entries): the package list in file order.
 */
typedef struct {
    uint64_t pos;       /* offset of the chunk in the input */
    uint32_t orig_len;  /* chunk size                        */
    uint32_t comp_len;  /* payload bytes                     */
    uint32_t type;      /* package type                      */
    uint32_t pad;
} ambc_chunk_info;
uint64_t ambc_compress_dynamic_workspace_bytes(uint64_t n, const uint32_t *cands, uint32_t n_cands);
int ambc_compress_dynamic_dev(const void *in_dev, uint64_t n, const uint32_t *cands, uint32_t n_cands,
                              uint32_t method_mask, uint32_t flags, const uint8_t *marker, uint32_t marker_bytes,
                              void *out_dev, uint64_t out_cap, void *work_dev, uint64_t work_bytes,
                              ambc_compress_result *res, ambc_chunk_info *map_out, uint64_t map_cap, void *stream);

/* ------------------------------------------------------------------ */
/* decompress: AdaptiveCompressor._adaptive_decompress                  */
/*   adaptive_compressor.py:396-454                                     */
/* ------------------------------------------------------------------ */

typedef struct {
    uint64_t src_off;  /* payload offset in the body                      */
    uint64_t dst_off;  /* output offset                                   */
    uint32_t comp_len; /* payload bytes                                   */
    uint32_t orig_len; /* original_length field of the package header     */
    uint32_t type;     /* package type; AMBC_RAW also for unknown types   */
    uint32_t out_len;  /* bytes this entry contributes to the output      */
} ambc_pkg;

/*
 * Walk the package chain on the host (adaptive_compressor.py:399-430 and the stop
 * rule :444-445): validates markers, stops at END / truncated header / truncated
 * payload.  Raw packages are split into entries of at most 64 KiB.  `known_mask`
 * = method ids present in method_lookup (others are copied through, :432-435).
 * Pass table=NULL to count entries.  Returns AMBC_E_MARKER on a marker mismatch.
 */
int ambc_index_host(const uint8_t *body, uint64_t body_len, const uint8_t *marker, uint32_t marker_bytes,
                    uint64_t orig_size, uint32_t known_mask, ambc_pkg *table, uint64_t table_cap,
                    uint64_t *n_entries, uint64_t *out_bytes);

/*
 * The same table built on the GPU from a device-resident body (no host copy of the body needed):
 * marker scan, header parse, successor links, pointer jumping from position 0, placement scans.
 * Only positions reachable from position 0 become packages, exactly as in the serial walk; a marker
 * mismatch the walk would hit returns AMBC_E_MARKER.  table_dev may be NULL to obtain *n_entries
 * first.  Synchronises `stream`.
 */
int ambc_index_dev(const void *body_dev, uint64_t body_len, const uint8_t *marker, uint32_t marker_bytes,
                   uint64_t orig_size, uint32_t known_mask, ambc_pkg *table_dev, uint64_t table_cap,
                   uint64_t *n_entries, uint64_t *out_bytes, void *stream);

/*
 * Decode every table entry into out_dev (orig_size bytes; zero padded / truncated
 * as adaptive_compressor.py:447-452).  status_dev (optional, uint32[2]):
 * [0] = entries whose codec raised (output zero-filled, :440-442),
 * [1] = entries whose decoded length differed from out_len (malformed stream).
 */
int ambc_decompress_dev(const void *body_dev, uint64_t body_len, const ambc_pkg *table_dev, uint64_t n_entries,
                        void *out_dev, uint64_t orig_size, uint32_t *status_dev, void *stream);

/* host buffers: index walk + H2D + kernels + D2H.  status (optional) as above. */
int ambc_decompress_host(const void *body_host, uint64_t body_len, const uint8_t *marker, uint32_t marker_bytes,
                         uint32_t known_mask, void *out_host, uint64_t orig_size, uint32_t *status);

/* ------------------------------------------------------------------ */
/* codec plug-ins: CompressionMethod.compress / decompress / should_use */
/*   compression_methods.py:78-180 (RLE), :195-343 (Dictionary),         */
/*   :354-574 (Huffman), :585-667 (Delta), :678-713 (NoCompression);     */
/*   method 5 = DeflateCompression, advanced_compression.py:71-107, where */
/*   the reference calls zlib: encode writes a conforming zlib stream     */
/*   (RFC 1950 / 1951, one fixed-Huffman block) that stock zlib reads --  */
/*   not zlib.compress(level=9)'s own bytes, hence a plug-in reported     */
/*   separately and no candidate of the chunk trial; decode inflates any  */
/*   zlib stream, pads / truncates to original_length, and returns        */
/*   original_length zero bytes where zlib.decompress raises (:93-97).    */
/*   known_mask bit 5 makes packages of type 5 decodable in a body.       */
/* ------------------------------------------------------------------ */

/*
 * Item i = in_dev[in_off[i] .. in_off[i+1]) (at most AMBC_MAX_CODEC_CHUNK bytes).
 * Payload i is written at out_dev + i*out_stride; out_len[i] = payload length or
 * AMBC_CODEC_*_ERROR.  out_stride must be >= ambc_codec_bound(method, max item).
 */
uint64_t ambc_codec_bound(int method, uint32_t n);
int ambc_codec_encode_batch(int method, const void *in_dev, const uint64_t *in_off_dev, uint32_t n_items,
                            void *out_dev, uint64_t out_stride, int32_t *out_len_dev, void *stream);
/*
 * Item i = payload in_dev[in_off[i] .. in_off[i+1]) with original_length
 * orig_len[i]; decoded bytes at out_dev + i*out_stride, out_len[i] = number of
 * bytes the reference's decompress() returns, or AMBC_CODEC_INDEX_ERROR.
 */
int ambc_codec_decode_batch(int method, const void *in_dev, const uint64_t *in_off_dev,
                            const uint32_t *orig_len_dev, uint32_t n_items, void *out_dev, uint64_t out_stride,
                            int32_t *out_len_dev, void *stream);
/*
 * should_use of all four methods: gates[i] bit (id) set when method id's gate is
 * true; entropy_dev (optional) receives the fp64 entropy of
 * compression_methods.py:566-574.
 */
int ambc_should_use_batch(const void *in_dev, const uint64_t *in_off_dev, uint32_t n_items, uint8_t *gates_dev,
                          double *entropy_dev, void *stream);

/* ------------------------------------------------------------------ */
/* marker search: MarkerFinder.find_marker  (marker_finder.py:22-123)   */
/* ------------------------------------------------------------------ */

/*
 * Presence flags of all L-bit windows of the MSB-first bit stream of
 * in_dev[0..n): flags_dev[v] = 1 when value v occurs.  flags_dev holds 2^L bytes
 * and must be zeroed by the caller (so that several shards / GPUs can be merged
 * with a byte-wise max all-reduce before ambc_marker_pick).  `carry` = the
 * (L-1) bits that precede this shard in the global stream (low bits, MSB first),
 * carry_bits = how many of them exist (0 for the first shard).
 */
int ambc_marker_flags_dev(const void *in_dev, uint64_t n, uint32_t L, uint64_t carry, uint32_t carry_bits,
                          uint8_t *flags_dev, void *stream);
/*
 * From the flags of length L derive every shorter length (a shorter value is
 * present iff it prefixes a present L-bit value or equals the stream's last
 * window, given by `tail` / tail_bits = the last min(L-1, nbits) bits of the whole
 * stream) and return the smallest length in [1, min(L, max_len)] with an absent
 * value: *out_len bits, *out_value the smallest absent value.  AMBC_E_NO_MARKER
 * when every value up to min(L, max_len) is present.  total_bits = bits of the
 * whole stream.
 */
int ambc_marker_pick_dev(const uint8_t *flags_dev, uint32_t L, uint32_t max_len, uint64_t total_bits,
                         uint64_t tail, uint32_t tail_bits, uint32_t *out_len, uint64_t *out_value, void *stream);
/* single-GPU convenience: device input -> (marker value, length) */
int ambc_find_marker_dev(const void *in_dev, uint64_t n, uint32_t max_len, uint32_t *out_len, uint64_t *out_value,
                         void *stream);

/* ------------------------------------------------------------------ */
/* synthetic corpus (bench / tests): counter-based, any byte range       */
/* ------------------------------------------------------------------ */
/* fills out_dev[0..n) with bytes [offset, offset+n) of corpus (seed, kind_mask) */
int ambc_synth_dev(void *out_dev, uint64_t offset, uint64_t n, uint64_t seed, uint32_t kind_mask, void *stream);

/* ------------------------------------------------------------------ */
/* measurement support (bench.py): per-kernel device times              */
/* ------------------------------------------------------------------ */
/* when enabled, ambc_compress_dev / ambc_decompress_dev bracket their kernels with CUDA events
 * on the caller's stream; ambc_last_timing returns the most recent call's milliseconds:
 * ms[0] k_select, ms[1] size scan (3 small kernels), ms[2] k_pack, ms[3] k_decode (+tail zero). */
void ambc_enable_timing(int on);
int ambc_last_timing(float *ms4);

/* ---- multi-GPU placement ------------------------------------------------------------
 * Shards are contiguous chunk ranges compressed independently ("as if no earlier shard had hit a chunk
 * without a winner").  Every rank all-gathers one 16-byte record (the only exchange of the data path;
 * torch.distributed / NCCL moves it, this call folds it): packed_bytes = bytes of its local body before
 * its first raw chunk (END and local raw package removed), first_raw = GLOBAL index of its first chunk
 * without a winner or -1.  The fold applies the rest-of-file-raw rule (adaptive_compressor.py:586-590)
 * across shards: out[r].offset = byte offset of rank r's contribution in the global body and
 * out[r].state = AMBC_SHARD_PACKED (its packed fragment), AMBC_SHARD_RAW_STARTS (packed part, then the
 * header of the one global raw package and its input from the raw chunk on) or AMBC_SHARD_IN_RAW_TAIL
 * (its input bytes, verbatim, inside that package).  first_byte[r] = input byte offset of rank r's shard
 * (a multiple of the chunk size, or the input length for an empty trailing rank -- byte offsets rather than
 * chunk indices so that such a rank lands at the end of the raw data when the last chunk is partial).
 * Pure host arithmetic; needs no device. */
typedef struct { uint64_t packed_bytes; int64_t first_raw; } ambc_shard_rec;
typedef struct { uint64_t offset; uint32_t state; uint32_t reserved; } ambc_shard_slot;
#define AMBC_SHARD_PACKED 0
#define AMBC_SHARD_RAW_STARTS 1
#define AMBC_SHARD_IN_RAW_TAIL 2
int ambc_shard_place(const ambc_shard_rec *recs, uint32_t n_ranks, const uint64_t *first_byte, uint32_t chunk,
                     uint32_t marker_bytes, ambc_shard_slot *out);

/* ---- threading ----------------------------------------------------------------------------------------
 * The *_host calls serialise on a library mutex.  The *_dev calls keep per-DEVICE scratch (side stream and
 * events of the decoders, index scratch, kernel timers): call them from ONE host thread per device at a
 * time (different devices may be driven concurrently, one thread each). */

/* ---- test and experiment hooks (exported, not part of the reference-facing surface) -------------------
 * ambc_set_lz_levels / ambc_set_lz_coop_threshold / ambc_set_lz_force_buckets steer the round-1 bucket search
 * (kept behind AMBC_SELECT=old and for the plug-in codec kernels); ambc_set_walk_threads lets small bodies
 * exercise the helper threads of ambc_index_host (min_bytes = body size from which helpers start, threads =
 * 0 for the host cores divided by LOCAL_WORLD_SIZE, at most 8); ambc_scan_state_bytes = size of the per-piece
 * state record ambc_compress_host downloads.  None of them changes any result. */
int ambc_set_lz_levels(const int *levels, int n);
int ambc_set_lz_coop_threshold(int t);
int ambc_set_lz_force_buckets(int on);
void ambc_set_walk_threads(uint64_t min_bytes, unsigned threads);
uint64_t ambc_scan_state_bytes(void);

/* pinned host memory helpers for callers without their own allocator */
void *ambc_host_alloc(uint64_t bytes);
void ambc_host_free(void *p);

#ifdef __cplusplus
}
#endif
#endif /* AMBC_H */
