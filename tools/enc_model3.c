/* enc_model2.c -- CPU model of the two-phase GPU match finder (development tool).
 * Phase A: for every position p, cand[p] = nearest earlier position with the same hash, as the
 *          round-parallel kernel finds it: exact for positions of earlier rounds (W positions per
 *          round), and within the round only among the same 32-position group (warp match_any).
 * Phase C: unit-parallel greedy/lazy walk over cand[], units of U bytes, matches truncated at unit ends. */
#define ZSTD_STATIC_LINKING_ONLY
#include "zstd.h"
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static uint64_t rd64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }
static uint32_t hash4(uint32_t v, int bits) { return (v * 2654435761u) >> (32 - bits); }
static uint32_t hash5(uint64_t v, int bits) { return (uint32_t)(((v << 24) * 889523592379ULL) >> (64 - bits)); }

static int g_llog = 0;           /* 0 = no long table; else log2 of the 8-byte-hash table */
void model3_set_long(int llog) { g_llog = llog; }
static uint32_t hash8(uint64_t v, int bits) { return (uint32_t)((v * 0xCF1BBCDCB7A56463ULL) >> (64 - bits)); }
static uint32_t count(const uint8_t* src, uint32_t a, uint32_t b, uint32_t end) { uint32_t m = 0; while (b + m < end && src[a + m] == src[b + m]) m++; return m; }

size_t model2_parse(const uint8_t* src, uint32_t n, uint32_t U, int hlog, int mml, uint32_t W, int lazy, int maxdist,
                    ZSTD_Sequence* out, size_t out_cap)
{
    uint32_t* head = (uint32_t*)malloc(sizeof(uint32_t) << hlog);
    uint32_t* cand = (uint32_t*)malloc(sizeof(uint32_t) * (n + 8));
    uint32_t* candL = (uint32_t*)malloc(sizeof(uint32_t) * (n + 8));
    uint32_t* headL = g_llog ? (uint32_t*)malloc(sizeof(uint32_t) << g_llog) : 0;
    if (headL) memset(headL, 0xFF, sizeof(uint32_t) << g_llog);
    uint32_t p; size_t total = 0;
    uint32_t const ilimit = n >= 8 ? n - 8 : 0;
    memset(head, 0xFF, sizeof(uint32_t) << hlog);
    /* phase A */
    if (W == 0) W = 1;
    for (uint32_t r0 = 0; r0 < n; r0 += W) {
        uint32_t r1 = r0 + W < n ? r0 + W : n;
        for (p = r0; p < r1; p++) {
            if (p > ilimit) { cand[p] = 0xFFFFFFFFu; continue; }
            uint32_t h = mml == 5 ? hash5(rd64(src + p), hlog) : hash4(rd32(src + p), hlog);
            uint32_t c = head[h];                       /* from earlier rounds */
            /* same 32-group, earlier lane with the same hash */
            uint32_t g0 = p & ~31u; if (g0 < r0) g0 = r0;
            for (uint32_t q = p; q-- > g0;) { uint32_t hq = mml == 5 ? hash5(rd64(src + q), hlog) : hash4(rd32(src + q), hlog); if (hq == h) { c = q; break; } }
            cand[p] = c;
            candL[p] = headL ? headL[hash8(rd64(src + p), g_llog)] : 0xFFFFFFFFu;
        }
        for (p = r0; p < r1 && p <= ilimit; p++) { uint32_t h = mml == 5 ? hash5(rd64(src + p), hlog) : hash4(rd32(src + p), hlog); head[h] = p; if (headL) headL[hash8(rd64(src + p), g_llog)] = p; }
    }
    /* phase C */
    uint32_t carry = 0;
    for (uint32_t u0 = 0; u0 < n; u0 += U) {
        uint32_t end = u0 + U < n ? u0 + U : n, ip = u0, anchor = u0, rep1 = 0, rep2 = 0;
        while (ip < end && ip <= ilimit) {
            uint32_t start = 0, ml = 0, off = 0;
            if (rep1 && ip + 1 <= ilimit && ip + 1 >= rep1 && rd32(src + ip + 1) == rd32(src + ip + 1 - rep1)) {
                start = ip + 1; off = rep1; ml = count(src, start - off, start, end);
            } else {
                uint32_t c = cand[ip];
                { uint32_t cl = candL[ip]; if (cl < ip && ip - cl <= (uint32_t)maxdist && rd64(src + cl) == rd64(src + ip)) c = cl; }
                if (c < ip && ip - c <= (uint32_t)maxdist && rd32(src + c) == rd32(src + ip)) {
                    uint32_t m = count(src, c, ip, end);
                    if (m >= (uint32_t)mml) {
                        if (lazy && ip + 1 <= ilimit && ip + 1 < end) {      /* one-step lazy: a better match one byte later? */
                            uint32_t c2 = cand[ip + 1];
                            { uint32_t cl = candL[ip + 1]; if (cl < ip + 1 && ip + 1 - cl <= (uint32_t)maxdist && rd64(src + cl) == rd64(src + ip + 1)) c2 = cl; }
                            if (c2 < ip + 1 && ip + 1 - c2 <= (uint32_t)maxdist && rd32(src + c2) == rd32(src + ip + 1)) {
                                uint32_t m2 = count(src, c2, ip + 1, end);
                                if (m2 > m + 1) { ip++; continue; }
                            }
                        }
                        start = ip; off = ip - c; ml = m;
                        while (start > anchor && start - off > 0 && src[start - 1] == src[start - off - 1]) { start--; ml++; }
                    }
                }
            }
            if (ml < 4) { ip += 1 + ((ip - anchor) >> 8); continue; }
            if (total < out_cap) { out[total].offset = off; out[total].litLength = start - carry; out[total].matchLength = ml; out[total].rep = 0; }
            total++; carry = start + ml;
            if (off != rep1) { rep2 = rep1; rep1 = off; }
            ip = start + ml; anchor = ip;
            while (rep2 && ip <= ilimit && ip < end && ip >= rep2 && rd32(src + ip) == rd32(src + ip - rep2)) {
                uint32_t o = rep2, m = count(src, ip - o, ip, end);
                if (m < 4) break;
                if (total < out_cap) { out[total].offset = o; out[total].litLength = 0; out[total].matchLength = m; out[total].rep = 0; }
                total++; carry = ip + m; rep2 = rep1; rep1 = o; ip += m; anchor = ip;
            }
        }
    }
    free(head); free(cand); free(candL); if (headL) free(headL);
    return total;
}

size_t model2_compress(const uint8_t* src, size_t n, uint32_t blk, uint32_t U, int hlog, int mml, uint32_t W, int lazy, int maxdist, size_t* nseq_out)
{
    ZSTD_CCtx* c = ZSTD_createCCtx(); size_t pos, total = 0, ns = 0;
    ZSTD_Sequence* seqs = (ZSTD_Sequence*)malloc(sizeof(ZSTD_Sequence) * (blk / 3 + 16));
    size_t cap = ZSTD_compressBound(blk) + 64; void* dst = malloc(cap);
    ZSTD_CCtx_setParameter(c, ZSTD_c_compressionLevel, 3);
    ZSTD_CCtx_setParameter(c, ZSTD_c_blockDelimiters, ZSTD_sf_noBlockDelimiters);
    for (pos = 0; pos < n; pos += blk) {
        uint32_t len = (uint32_t)(n - pos < blk ? n - pos : blk);
        size_t k = model2_parse(src + pos, len, U, hlog, mml, W, lazy, maxdist, seqs, blk / 3 + 16);
        size_t r = ZSTD_compressSequences(c, dst, cap, seqs, k, src + pos, len);
        if (ZSTD_isError(r)) { total = (size_t)-1; break; }
        total += r; ns += k;
    }
    if (nseq_out) *nseq_out = ns;
    free(seqs); free(dst); ZSTD_freeCCtx(c);
    return total;
}
