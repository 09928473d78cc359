/* enc_model4.c -- CPU ratio model of the round-2 block compressor (development tool; not product code).
 *
 * Models what zb_compress_smem does so that table sizes / unit sizes can be chosen before the kernel is written:
 *   links   near: every SC-byte sub-chunk has a private table of 2^near_bits slots (+ tag bits), walked 32 positions per
 *                 step (the 32 positions of a step do not see each other);
 *           far : one table of 2^far_bits slots holding the latest position before the current ROUND of R positions
 *   verify  every position: match length of both candidates (capped at CAP), longer wins -> mlen[p], dist[p]
 *   parse   lanes own units of U bytes; a match may START only inside the unit but extends freely; one-step lazy;
 *           optional repcode probe at ip+1; units are stitched afterwards (front-trim / drop what an earlier unit covers)
 * The sequences are entropy-coded by the reference (ZSTD_compressSequences), which also finds the repcodes.
 */
#define ZSTD_STATIC_LINKING_ONLY
#include "zstd.h"
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int SC, near_bits, tag_bits, far_bits, far_mml, R, cap, lazy, U, rep_probe, backext, maxdist, near_mml, exact_step, step, small_mask;
} M4Params;

static uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static uint64_t rd64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }
static uint32_t hashN(const uint8_t* p, int mml, int bits)
{
    if (bits <= 0) return 0;
    if (mml == 4) return (rd32(p) * 2654435761u) >> (32 - bits);
    if (mml == 5) return (uint32_t)(((rd64(p) << 24) * 889523592379ULL) >> (64 - bits));
    if (mml == 6) return (uint32_t)(((rd64(p) << 16) * 227718039650203ULL) >> (64 - bits));
    return (uint32_t)((rd64(p) * 0xCF1BBCDCB7A56463ULL) >> (64 - bits));
}
static uint32_t count(const uint8_t* s, uint32_t a, uint32_t b, uint32_t end, uint32_t cap)
{
    uint32_t m = 0; while (b + m < end && m < cap && s[a + m] == s[b + m]) m++; return m;
}

typedef struct { uint32_t start, len, dist; } Rec;

size_t model4_parse(const uint8_t* src, uint32_t n, const M4Params* P, ZSTD_Sequence* out, size_t out_cap, uint32_t* tail_lits)
{
    uint32_t* nearT = (uint32_t*)malloc(sizeof(uint32_t) << P->near_bits);
    uint32_t* nearTag = (uint32_t*)malloc(sizeof(uint32_t) << P->near_bits);
    uint32_t* farT = (uint32_t*)malloc(sizeof(uint32_t) << P->far_bits);
    uint8_t* mlen = (uint8_t*)calloc(n + 16, 1);
    uint32_t* dist = (uint32_t*)calloc(n + 16, 4);
    uint32_t* ncand = (uint32_t*)malloc(4 * (n + 16));
    uint32_t* fcand = (uint32_t*)malloc(4 * (n + 16));
    uint32_t const NONE = 0xFFFFFFFFu;
    uint32_t const lim = n >= 8 ? n - 8 : 0;          /* positions that may be hashed (8 readable bytes) */
    uint32_t p;
    /* near */
    for (uint32_t c0 = 0; c0 < n; c0 += (uint32_t)P->SC) {
        uint32_t c1 = c0 + P->SC < n ? c0 + P->SC : n;
        memset(nearT, 0xFF, sizeof(uint32_t) << P->near_bits);
        for (uint32_t s0 = c0; s0 < c1; s0 += (uint32_t)P->step) {
            uint32_t s1 = s0 + P->step < c1 ? s0 + P->step : c1;
            for (p = s0; p < s1; p++) {
                ncand[p] = NONE;
                if (p > lim) continue;
                uint32_t h = hashN(src + p, P->near_mml, P->near_bits + P->tag_bits);
                uint32_t slot = h >> P->tag_bits, tag = h & ((1u << P->tag_bits) - 1);
                if (nearT[slot] != NONE && nearTag[slot] == tag) ncand[p] = nearT[slot];
                if (P->exact_step) for (uint32_t q = p; q-- > s0;) if (q <= lim && hashN(src + q, P->near_mml, P->near_bits + P->tag_bits) == h) { ncand[p] = q; break; }
            }
            for (p = s0; p < s1; p++) {
                if (p > lim) continue;
                uint32_t h = hashN(src + p, P->near_mml, P->near_bits + P->tag_bits);
                nearT[h >> P->tag_bits] = p; nearTag[h >> P->tag_bits] = h & ((1u << P->tag_bits) - 1);
            }
        }
    }
    /* far */
    memset(farT, 0xFF, sizeof(uint32_t) << P->far_bits);
    for (uint32_t r0 = 0; r0 < n; r0 += (uint32_t)P->R) {
        uint32_t r1 = r0 + P->R < n ? r0 + P->R : n;
        for (p = r0; p < r1; p++) { fcand[p] = NONE; if (p <= lim && P->far_bits > 0) fcand[p] = farT[hashN(src + p, P->far_mml, P->far_bits)]; }
        for (p = r0; p < r1; p++) if (p <= lim && P->far_bits > 0) farT[hashN(src + p, P->far_mml, P->far_bits)] = p;
    }
    /* verify */
    for (p = 0; p < n; p++) {
        uint32_t best = 0, bd = 0;
        if (ncand[p] != NONE && p - ncand[p] <= (uint32_t)P->maxdist) { uint32_t m = count(src, ncand[p], p, n, P->cap); if (m >= 4) { best = m; bd = p - ncand[p]; } }
        if (fcand[p] != NONE && p - fcand[p] <= (uint32_t)P->maxdist) { uint32_t m = count(src, fcand[p], p, n, P->cap); if (m >= 4 && m > best) { best = m; bd = p - fcand[p]; } }
        for (uint32_t d = 1; d <= 32; d++) if (((uint32_t)P->small_mask >> (d - 1)) & 1) { if (p >= d) { uint32_t m = count(src, p - d, p, n, P->cap); if (m >= 4 && m > best) { best = m; bd = d; } } }
        mlen[p] = (uint8_t)best; dist[p] = bd;
    }
    /* parse per unit */
    uint32_t const units = (n + P->U - 1) / P->U;
    Rec* recs = (Rec*)malloc(sizeof(Rec) * (n / 3 + 64));
    uint32_t* ufirst = (uint32_t*)malloc(4 * (units + 1));
    uint32_t nrec = 0;
    for (uint32_t u = 0; u < units; u++) {
        uint32_t u0 = u * P->U, u1 = u0 + P->U < n ? u0 + P->U : n, ip = u0, anchor = u0, r0 = 0;
        ufirst[u] = nrec;
        if (u == 0) ip = 1;
        while (ip < u1) {
            uint32_t start = 0, len = 0, d = 0;
            if (P->rep_probe && r0 && ip + 1 < u1 && ip + 1 + 4 <= n && ip + 1 >= r0 && rd32(src + ip + 1) == rd32(src + ip + 1 - r0)) {
                start = ip + 1; d = r0; len = count(src, start - d, start, n, 1u << 30);
            } else {
                uint32_t m = mlen[ip];
                if (m >= 4) {
                    if (P->lazy && m < (uint32_t)P->cap && ip + 1 < u1 && mlen[ip + 1] > m + 1) { ip++; continue; }
                    start = ip; d = dist[ip]; len = m;
                    if (m == (uint32_t)P->cap) len = count(src, start - d, start, n, 1u << 30);
                    if (P->backext) while (start > anchor && start > d && src[start - 1] == src[start - d - 1]) { start--; len++; }
                }
            }
            if (len < 4) { ip++; continue; }
            recs[nrec].start = start; recs[nrec].len = len; recs[nrec].dist = d; nrec++;
            r0 = d; ip = start + len; anchor = ip;
        }
    }
    ufirst[units] = nrec;
    /* stitch */
    size_t total = 0; uint32_t cover = 0;
    for (uint32_t i = 0; i < nrec; i++) {
        uint32_t s = recs[i].start, e = s + recs[i].len;
        if (s < cover) { if (e < cover + 3) continue; s = cover; }
        if (total < out_cap) { out[total].offset = recs[i].dist; out[total].litLength = s - cover; out[total].matchLength = e - s; out[total].rep = 0; }
        total++; cover = e;
    }
    *tail_lits = n - cover;
    free(nearT); free(nearTag); free(farT); free(mlen); free(dist); free(ncand); free(fcand); free(recs); free(ufirst);
    return total;
}

size_t model4_compress(const uint8_t* src, size_t n, uint32_t blk, const M4Params* P, size_t* nseq_out)
{
    ZSTD_CCtx* c = ZSTD_createCCtx(); size_t pos, total = 0, ns = 0;
    ZSTD_Sequence* seqs = (ZSTD_Sequence*)malloc(sizeof(ZSTD_Sequence) * (blk / 3 + 16));
    size_t cap = ZSTD_compressBound(blk) + 64; void* dst = malloc(cap);
    ZSTD_CCtx_setParameter(c, ZSTD_c_compressionLevel, 3);
    ZSTD_CCtx_setParameter(c, ZSTD_c_blockDelimiters, ZSTD_sf_noBlockDelimiters);
    for (pos = 0; pos < n; pos += blk) {
        uint32_t len = (uint32_t)(n - pos < blk ? n - pos : blk), tail;
        size_t k = model4_parse(src + pos, len, P, seqs, blk / 3 + 16, &tail);
        size_t r = ZSTD_compressSequences(c, dst, cap, seqs, k, src + pos, len);
        if (ZSTD_isError(r)) { fprintf(stderr, "compressSequences: %s\n", ZSTD_getErrorName(r)); total = (size_t)-1; break; }
        total += r; ns += k;
    }
    if (nseq_out) *nseq_out = ns;
    free(seqs); free(dst); ZSTD_freeCCtx(c);
    return total;
}
