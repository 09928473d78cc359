/* enc_model.c -- CPU model of the GPU match finder (development tool: ratio experiments).
 * Lock-step simulation: T lanes, lane t greedily parses unit [t*U, (t+1)*U) of a block; all lanes
 * share one hash table; per round every lane does one step (reads of a round see the writes of
 * earlier rounds and of lower lanes in the same round -- a stand-in for the SIMT order).
 * Sequences are entropy-coded with the reference's ZSTD_compressSequences to get a size. */
#define ZSTD_STATIC_LINKING_ONLY
#include "zstd.h"
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { uint32_t pos, ll, ml, off; } mseq;   /* pos = start of literals */

static uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static uint64_t rd64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }
static uint32_t hash4(uint32_t v, int bits) { return (v * 2654435761u) >> (32 - bits); }
static uint32_t hash5(uint64_t v, int bits) { return (uint32_t)(((v << 24) * 889523592379ULL) >> (64 - bits)); }
static uint32_t hash8(uint64_t v, int bits) { return (uint32_t)((v * 0xCF1BBCDCB7A56463ULL) >> (64 - bits)); }

typedef struct { uint32_t ip, anchor, end, rep1, rep2, nseq; int done; } lane_t;

/* params: U unit bytes, hlog, mml (4|5), use_long (second table with 8-byte hash, hlog_long) */
size_t model_parse(const uint8_t* src, uint32_t n, uint32_t U, int hlog, int mml, int hlog_long,
                   ZSTD_Sequence* out, size_t out_cap, int lazy)
{
    uint32_t T = (n + U - 1) / U, t;
    uint32_t* tab = (uint32_t*)malloc(sizeof(uint32_t) << hlog);
    uint32_t* tabL = hlog_long ? (uint32_t*)malloc(sizeof(uint32_t) << hlog_long) : NULL;
    lane_t* L = (lane_t*)calloc(T, sizeof(lane_t));
    mseq** sq = (mseq**)calloc(T, sizeof(mseq*));
    uint32_t live = T; size_t total = 0;
    memset(tab, 0xFF, sizeof(uint32_t) << hlog);
    if (tabL) memset(tabL, 0xFF, sizeof(uint32_t) << hlog_long);
    for (t = 0; t < T; t++) { L[t].ip = L[t].anchor = t * U; L[t].end = (t + 1) * U < n ? (t + 1) * U : n; sq[t] = (mseq*)malloc(sizeof(mseq) * (U / 3 + 2)); }
    uint32_t const ilimit = n >= 8 ? n - 8 : 0;          /* last position where 8 bytes can be read */
    while (live) {
        for (t = 0; t < T; t++) {
            lane_t* l = &L[t];
            if (l->done) continue;
            uint32_t ip = l->ip;
            if (ip >= l->end || ip > ilimit) { l->done = 1; live--; continue; }
            uint32_t start = 0, ml = 0, off = 0;
            uint32_t h = mml == 5 ? hash5(rd64(src + ip), hlog) : hash4(rd32(src + ip), hlog);
            uint32_t cand = tab[h]; tab[h] = ip;
            uint32_t candL = 0xFFFFFFFFu;
            if (tabL) { uint32_t hl = hash8(rd64(src + ip), hlog_long); candL = tabL[hl]; tabL[hl] = ip; }
            if (l->rep1 && ip + 1 <= ilimit && ip + 1 >= l->rep1 && rd32(src + ip + 1) == rd32(src + ip + 1 - l->rep1)) {
                start = ip + 1; off = l->rep1;
                ml = 4; while (start + ml < l->end && src[start + ml] == src[start + ml - off]) ml++;
            } else {
                uint32_t bestml = 0, bestc = 0;
                if (candL < ip && rd64(src + candL) == rd64(src + ip)) {
                    uint32_t m = 8; while (ip + m < l->end && src[ip + m] == src[candL + m]) m++;
                    bestml = m; bestc = candL;
                }
                if (cand < ip && rd32(src + cand) == rd32(src + ip) && (mml == 4 || src[cand + 4] == src[ip + 4])) {
                    uint32_t m = 4; while (ip + m < l->end && src[ip + m] == src[cand + m]) m++;
                    if (m > bestml) { bestml = m; bestc = cand; }
                }
                if (bestml >= (uint32_t)mml) {
                    start = ip; off = ip - bestc; ml = bestml;
                    while (start > l->anchor && start - off > 0 && src[start - 1] == src[start - off - 1]) { start--; ml++; }
                }
            }
            if (!ml) { l->ip = ip + 1 + ((ip - l->anchor) >> 8); continue; }
            if (start + ml > l->end) ml = l->end - start;
            if (ml < 4) { l->ip = ip + 1; continue; }
            { mseq* s = &sq[t][l->nseq++]; s->pos = l->anchor; s->ll = start - l->anchor; s->ml = ml; s->off = off; }
            if (off != l->rep1) { l->rep2 = l->rep1; l->rep1 = off; }
            ip = start + ml; l->anchor = ip;
            /* fill the table at a couple of covered positions (cheap, helps later matches) */
            if (ip - 2 <= ilimit && ip >= 2) { uint32_t q = ip - 2; uint32_t hq = mml == 5 ? hash5(rd64(src + q), hlog) : hash4(rd32(src + q), hlog); tab[hq] = q;
                if (tabL) tabL[hash8(rd64(src + q), hlog_long)] = q; }
            /* immediate repcode-2 matches */
            while (l->rep2 && ip <= ilimit && ip < l->end && ip >= l->rep2 && rd32(src + ip) == rd32(src + ip - l->rep2)) {
                uint32_t m = 4, o = l->rep2; while (ip + m < l->end && src[ip + m] == src[ip + m - o]) m++;
                if (ip + m > l->end) m = l->end - ip; if (m < 4) break;
                { mseq* s = &sq[t][l->nseq++]; s->pos = ip; s->ll = 0; s->ml = m; s->off = o; }
                l->rep2 = l->rep1; l->rep1 = o;
                ip += m; l->anchor = ip;
            }
            l->ip = ip;
        }
    }
    /* concatenate the lanes' sequences; leftover literals of a unit join the next sequence */
    { uint32_t carry_pos = 0;
      for (t = 0; t < T; t++) { uint32_t i;
        for (i = 0; i < L[t].nseq; i++) { mseq* s = &sq[t][i];
            if (total < out_cap) { out[total].offset = s->off; out[total].litLength = s->pos + s->ll - carry_pos; out[total].matchLength = s->ml; out[total].rep = 0; }
            total++; carry_pos = s->pos + s->ll + s->ml; }
        free(sq[t]); } }
    free(sq); free(L); free(tab); free(tabL);
    (void)lazy;
    return total;
}

/* compress n bytes as independent blocks of `blk` bytes each in its own frame; returns total size */
size_t model_compress(const uint8_t* src, size_t n, uint32_t blk, uint32_t U, int hlog, int mml, int hlog_long, size_t* nseq_out)
{
    ZSTD_CCtx* c = ZSTD_createCCtx(); size_t pos, total = 0, ns = 0;
    ZSTD_Sequence* seqs = (ZSTD_Sequence*)malloc(sizeof(ZSTD_Sequence) * (blk / 3 + 16));
    size_t cap = ZSTD_compressBound(blk) + 64; void* dst = malloc(cap);
    ZSTD_CCtx_setParameter(c, ZSTD_c_compressionLevel, 3);
    ZSTD_CCtx_setParameter(c, ZSTD_c_blockDelimiters, ZSTD_sf_noBlockDelimiters);
    ZSTD_CCtx_setParameter(c, ZSTD_c_validateSequences, 1);
    for (pos = 0; pos < n; pos += blk) {
        uint32_t len = (uint32_t)(n - pos < blk ? n - pos : blk);
        size_t k = model_parse(src + pos, len, U, hlog, mml, hlog_long, seqs, blk / 3 + 16, 0);
        size_t r = ZSTD_compressSequences(c, dst, cap, seqs, k, src + pos, len);
        if (ZSTD_isError(r)) { total = (size_t)-1; break; }
        total += r; ns += k;
    }
    if (nseq_out) *nseq_out = ns;
    free(seqs); free(dst); ZSTD_freeCCtx(c);
    return total;
}
