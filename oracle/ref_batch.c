/*
 * ref_batch.c -- TEST/BENCH INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A thin pthread driver around the UNMODIFIED reference codec
 * (oracle/_ref/libzstd_ref.so, built from /root/reference/zstd/zstd.c by
 * oracle/Makefile).  It restates the orchestration of the reference batch
 * path so the CPU baseline can be timed without the Python extension:
 *   compress_worker / compress_from_datasources   c-ext/compressor.c:856-1076, :1084-1336
 *   decompress_worker / decompress_from_framesources c-ext/decompressor.c:944-1181, :1186-1455
 * i.e. a static contiguous partition of the segments by input bytes, one
 * context per worker, ZSTD_CCtx_setPledgedSrcSize + ZSTD_compressStream2(e_end)
 * (c-ext/compressor.c:1035-1044) or ZSTD_decompressStream (c-ext/decompressor.c:1150)
 * per segment, every worker malloc()ing its own output arena.
 */
#define ZSTD_STATIC_LINKING_ONLY
#include "zstd.h"
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    /* inputs */
    const uint8_t* src; const uint64_t* off; const uint64_t* len; const uint64_t* dst_len;
    size_t first, last;              /* segment range [first, last) */
    int level, checksum, compress;
    const void* dict; size_t dict_size;
    /* outputs */
    uint8_t* arena; size_t arena_size;
    uint64_t* out_off; uint64_t* out_len;   /* indexed by global segment id */
    int err; size_t err_item; const char* err_msg;
} rb_worker;

typedef struct { rb_worker* w; int n_workers; size_t n; uint64_t* out_off; uint64_t* out_len; int* owner; } rb_result;

static void* run_worker(void* arg)
{
    rb_worker* w = (rb_worker*)arg; size_t i, pos = 0, cap = 0;
    if (w->compress) {
        ZSTD_CCtx* c = ZSTD_createCCtx();
        ZSTD_CCtx_setParameter(c, ZSTD_c_compressionLevel, w->level);
        ZSTD_CCtx_setParameter(c, ZSTD_c_checksumFlag, w->checksum);
        if (w->dict) ZSTD_CCtx_loadDictionary_byReference(c, w->dict, w->dict_size);
        for (i = w->first; i < w->last; i++) cap += ZSTD_compressBound(w->len[i]);
        w->arena = (uint8_t*)malloc(cap ? cap : 1); w->arena_size = cap;
        for (i = w->first; i < w->last; i++) {
            ZSTD_inBuffer in = { w->src + w->off[i], w->len[i], 0 };
            ZSTD_outBuffer out = { w->arena + pos, cap - pos, 0 };
            size_t r;
            ZSTD_CCtx_setPledgedSrcSize(c, w->len[i]);
            r = ZSTD_compressStream2(c, &out, &in, ZSTD_e_end);
            if (ZSTD_isError(r) || r != 0) { w->err = 1; w->err_item = i; w->err_msg = ZSTD_getErrorName(r); break; }
            w->out_off[i] = pos; w->out_len[i] = out.pos; pos += out.pos;
        }
        ZSTD_freeCCtx(c);
    } else {
        ZSTD_DCtx* d = ZSTD_createDCtx();
        if (w->dict) ZSTD_DCtx_loadDictionary_byReference(d, w->dict, w->dict_size);
        for (i = w->first; i < w->last; i++) {
            uint64_t sz = w->dst_len ? w->dst_len[i] : ZSTD_getFrameContentSize(w->src + w->off[i], w->len[i]);
            if (sz == ZSTD_CONTENTSIZE_UNKNOWN || sz == ZSTD_CONTENTSIZE_ERROR) { w->err = 2; w->err_item = i; w->err_msg = "unknown size"; }
            else cap += sz;
        }
        w->arena = (uint8_t*)malloc(cap ? cap : 1); w->arena_size = cap;
        for (i = w->first; i < w->last && !w->err; i++) {
            uint64_t sz = w->dst_len ? w->dst_len[i] : ZSTD_getFrameContentSize(w->src + w->off[i], w->len[i]);
            ZSTD_inBuffer in = { w->src + w->off[i], w->len[i], 0 };
            ZSTD_outBuffer out = { w->arena + pos, sz, 0 };
            size_t r = ZSTD_decompressStream(d, &out, &in);
            if (ZSTD_isError(r)) { w->err = 1; w->err_item = i; w->err_msg = ZSTD_getErrorName(r); break; }
            if (r != 0 || out.pos != sz) { w->err = 3; w->err_item = i; w->err_msg = "size mismatch"; break; }
            w->out_off[i] = pos; w->out_len[i] = out.pos; pos += out.pos;
        }
        ZSTD_freeDCtx(d);
    }
    return NULL;
}

/* returns a handle (rb_result*), NULL on allocation failure */
void* rb_run(int compress, int level, int checksum, const void* dict, size_t dict_size,
             const uint8_t* src, const uint64_t* off, const uint64_t* len, const uint64_t* dst_len,
             size_t n, int threads)
{
    rb_result* R = (rb_result*)calloc(1, sizeof(*R)); size_t i; uint64_t total = 0, per, acc = 0; int t = 0;
    pthread_t* th;
    if (threads < 1) threads = 1;
    if ((size_t)threads > n) threads = (int)(n ? n : 1);
    R->w = (rb_worker*)calloc((size_t)threads, sizeof(rb_worker)); R->n = n;
    R->out_off = (uint64_t*)calloc(n ? n : 1, 8); R->out_len = (uint64_t*)calloc(n ? n : 1, 8);
    R->owner = (int*)calloc(n ? n : 1, sizeof(int));
    for (i = 0; i < n; i++) total += len[i];
    per = total / (uint64_t)threads;
    /* static contiguous partition: cut when the worker holds >= total/threads bytes (c-ext/compressor.c:1183-1200) */
    R->w[0].first = 0;
    for (i = 0; i < n; i++) {
        acc += len[i]; R->owner[i] = t;
        if (acc >= per && t < threads - 1 && i + 1 < n) { R->w[t].last = i + 1; t++; R->w[t].first = i + 1; acc = 0; }
    }
    R->w[t].last = n; R->n_workers = t + 1;
    th = (pthread_t*)calloc((size_t)R->n_workers, sizeof(pthread_t));
    for (t = 0; t < R->n_workers; t++) {
        rb_worker* w = &R->w[t];
        w->src = src; w->off = off; w->len = len; w->dst_len = dst_len; w->level = level; w->checksum = checksum;
        w->compress = compress; w->dict = dict_size ? dict : NULL; w->dict_size = dict_size;
        w->out_off = R->out_off; w->out_len = R->out_len;
        if (R->n_workers == 1) run_worker(w); else pthread_create(&th[t], NULL, run_worker, w);
    }
    if (R->n_workers > 1) for (t = 0; t < R->n_workers; t++) pthread_join(th[t], NULL);
    free(th);
    return R;
}

int rb_error(void* h, size_t* item, const char** msg)
{
    rb_result* R = (rb_result*)h; int t;
    for (t = 0; t < R->n_workers; t++) if (R->w[t].err) { *item = R->w[t].err_item; *msg = R->w[t].err_msg; return R->w[t].err; }
    return 0;
}
uint64_t rb_total(void* h) { rb_result* R = (rb_result*)h; uint64_t s = 0; size_t i; for (i = 0; i < R->n; i++) s += R->out_len[i]; return s; }
const uint8_t* rb_get(void* h, size_t i, uint64_t* len)
{
    rb_result* R = (rb_result*)h; *len = R->out_len[i]; return R->w[R->owner[i]].arena + R->out_off[i];
}
/* gather every output into one contiguous buffer + lengths (convenience for tests) */
void rb_gather(void* h, uint8_t* dst, uint64_t* lens)
{
    rb_result* R = (rb_result*)h; size_t i; uint64_t pos = 0;
    for (i = 0; i < R->n; i++) { memcpy(dst + pos, R->w[R->owner[i]].arena + R->out_off[i], R->out_len[i]); lens[i] = R->out_len[i]; pos += R->out_len[i]; }
}
void rb_free(void* h)
{
    rb_result* R = (rb_result*)h; int t;
    for (t = 0; t < R->n_workers; t++) free(R->w[t].arena);
    free(R->w); free(R->out_off); free(R->out_len); free(R->owner); free(R);
}
