/*
 * zstd_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C) of the zstd frame/block decoder that the reference
 * batch path runs per segment (python-zstandard c-ext/decompressor.c:1150 ->
 * vendored zstd/zstd.c ZSTD_decompressStream -> ZSTD_decompressFrame).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it.
 *
 * Parity is PINNED: tests/test_oracle.py checks this file against the reference's
 * own golden vectors (reference tests/test_compressor_compress.py:19,28,34,58-68,
 * tests/test_compressor_multi_compress_to_buffer.py:45,64) and against frames
 * produced by the compiled reference itself (oracle/_ref/libzstd_ref.so).
 */
#ifndef ZSTD_ORACLE_H
#define ZSTD_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* error codes; names follow zstd_errors.h so messages can be compared */
enum {
    ZO_OK = 0,
    ZO_ERR_GENERIC = 1,
    ZO_ERR_PREFIX_UNKNOWN = 10,
    ZO_ERR_FRAMEPARAM_UNSUPPORTED = 14,
    ZO_ERR_WINDOW_TOO_LARGE = 16,
    ZO_ERR_CORRUPTION = 20,
    ZO_ERR_CHECKSUM_WRONG = 22,
    ZO_ERR_LITERALS_HEADER_WRONG = 24,
    ZO_ERR_DICT_CORRUPTED = 30,
    ZO_ERR_DICT_WRONG = 32,
    ZO_ERR_TABLELOG_TOO_LARGE = 44,
    ZO_ERR_MAXSYMBOL_TOO_SMALL = 48,
    ZO_ERR_DSTSIZE_TOO_SMALL = 70,
    ZO_ERR_SRCSIZE_WRONG = 72
};

typedef struct {
    uint64_t content_size;     /* UINT64_MAX when unknown */
    uint64_t window_size;
    uint32_t dict_id;
    uint32_t header_size;
    int      has_checksum;
    int      single_segment;
    int      is_skippable;
    uint32_t skippable_size;   /* payload bytes of a skippable frame */
} zo_frame_header;

/* one decoded sequence, after repcode resolution (offset = real distance) */
typedef struct { uint32_t lit_len, match_len, offset; } zo_seq;

/* optional trace of the entropy stage, for debugging GPU kernels */
typedef struct {
    zo_seq*  seqs;      size_t seq_cap,  n_seqs;     /* all blocks, concatenated */
    uint8_t* lits;      size_t lit_cap,  n_lits;     /* all blocks, concatenated */
    uint32_t* block_nseq; uint32_t* block_nlit; size_t block_cap, n_blocks;
} zo_trace;

const char* zo_error_name(int code);

/* parse a frame header. returns 0, or ZO_ERR_*; ZO_ERR_SRCSIZE_WRONG if truncated */
int zo_get_frame_header(zo_frame_header* h, const void* src, size_t src_size);

/* total compressed size of the first frame in src (header+blocks+checksum); 0 on error */
size_t zo_find_frame_compressed_size(const void* src, size_t src_size, int* err);

/* decode ONE frame. dict may be NULL. returns bytes written, sets *err (0 ok).
 * *consumed (optional) receives the compressed bytes read. */
size_t zo_decompress_frame(void* dst, size_t dst_cap,
                           const void* src, size_t src_size,
                           const void* dict, size_t dict_size,
                           size_t* consumed, zo_trace* trace, int* err);

/* decode a whole segment the way the batch path does: skippable frames are
 * skipped, concatenated frames are all decoded. */
size_t zo_decompress(void* dst, size_t dst_cap, const void* src, size_t src_size,
                     const void* dict, size_t dict_size, int* err);

uint64_t zo_xxh64(const void* data, size_t len, uint64_t seed);

#ifdef __cplusplus
}
#endif
#endif
