/*
 * zstd_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see zstd_oracle.h).
 *
 * Plain-C restatement of the decoder half of the reference hot path.  Each
 * function names the reference function (file:line in /root/reference) whose
 * behaviour it restates.  Written from the format rules (RFC 8878) in a
 * position-counting style (a signed "bits remaining" counter instead of the
 * reference's pointer/container juggling) so that it is an independent check,
 * not a transliteration.
 *
 * Parity pinned by tests/test_oracle.py (golden vectors of the reference's own
 * tests + frames produced by oracle/_ref/libzstd_ref.so).
 */
#include "zstd_oracle.h"
#include <string.h>
#include <stdlib.h>

#define ZO_MAGIC           0xFD2FB528u   /* zstd/zstd.c:4440 */
#define ZO_MAGIC_DICT      0xEC30A437u   /* zstd/zstd.c:4441 */
#define ZO_MAGIC_SKIP      0x184D2A50u   /* zstd/zstd.c:4442 */
#define ZO_BLOCK_MAX       (128u << 10)
#define ZO_WINDOWLOG_MAX   31

typedef uint8_t  u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;

static u32 rd16(const u8* p) { return (u32)p[0] | ((u32)p[1] << 8); }
static u32 rd24(const u8* p) { return rd16(p) | ((u32)p[2] << 16); }
static u32 rd32(const u8* p) { return rd16(p) | (rd16(p + 2) << 16); }
static u64 rd64(const u8* p) { return (u64)rd32(p) | ((u64)rd32(p + 4) << 32); }
static int hibit(u32 v) { int n = 0; while (v >>= 1) n++; return n; }

const char* zo_error_name(int code)
{
    /* strings of ERR_getErrorString, zstd/zstd.c (error_private.c) */
    switch (code) {
    case ZO_OK: return "No error detected";
    case ZO_ERR_GENERIC: return "Error (generic)";
    case ZO_ERR_PREFIX_UNKNOWN: return "Unknown frame descriptor";
    case ZO_ERR_FRAMEPARAM_UNSUPPORTED: return "Unsupported frame parameter";
    case ZO_ERR_WINDOW_TOO_LARGE: return "Frame requires too much memory for decoding";
    case ZO_ERR_CORRUPTION: return "Data corruption detected";
    case ZO_ERR_CHECKSUM_WRONG: return "Restored data doesn't match checksum";
    case ZO_ERR_LITERALS_HEADER_WRONG: return "Header of Literals' block doesn't respect format specification";
    case ZO_ERR_DICT_CORRUPTED: return "Dictionary is corrupted";
    case ZO_ERR_DICT_WRONG: return "Dictionary mismatch";
    case ZO_ERR_TABLELOG_TOO_LARGE: return "tableLog requires too much memory : unsupported";
    case ZO_ERR_MAXSYMBOL_TOO_SMALL: return "Specified maxSymbolValue is too small";
    case ZO_ERR_DSTSIZE_TOO_SMALL: return "Destination buffer is too small";
    case ZO_ERR_SRCSIZE_WRONG: return "Src size is incorrect";
    default: return "Unspecified error code";
    }
}

/* ------------------------------------------------------------------ XXH64 */
/* restates XXH64 as used for the frame checksum, zstd/zstd.c:44260-44277 */
#define P1 0x9E3779B185EBCA87ULL
#define P2 0xC2B2AE3D27D4EB4FULL
#define P3 0x165667B19E3779F9ULL
#define P4 0x85EBCA77C2B2AE63ULL
#define P5 0x27D4EB2F165667C5ULL
static u64 rotl(u64 x, int r) { return (x << r) | (x >> (64 - r)); }
static u64 xround(u64 acc, u64 in) { return rotl(acc + in * P2, 31) * P1; }
static u64 xmerge(u64 h, u64 v) { return (h ^ xround(0, v)) * P1 + P4; }
uint64_t zo_xxh64(const void* data, size_t len, uint64_t seed)
{
    const u8* p = (const u8*)data; const u8* end = p + len; u64 h;
    if (len >= 32) {
        u64 v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
        do { v1 = xround(v1, rd64(p)); v2 = xround(v2, rd64(p + 8));
             v3 = xround(v3, rd64(p + 16)); v4 = xround(v4, rd64(p + 24)); p += 32;
        } while (p + 32 <= end);
        h = rotl(v1, 1) + rotl(v2, 7) + rotl(v3, 12) + rotl(v4, 18);
        h = xmerge(h, v1); h = xmerge(h, v2); h = xmerge(h, v3); h = xmerge(h, v4);
    } else h = seed + P5;
    h += (u64)len;
    while (p + 8 <= end) { h ^= xround(0, rd64(p)); h = rotl(h, 27) * P1 + P4; p += 8; }
    if (p + 4 <= end) { h ^= (u64)rd32(p) * P1; h = rotl(h, 23) * P2 + P3; p += 4; }
    while (p < end) { h ^= (*p++) * P5; h = rotl(h, 11) * P1; }
    h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
    return h;
}

/* ------------------------------------------------ backward bit reader */
/* restates BIT_initDStream / BIT_readBits / BIT_endOfDStream
 * (zstd/zstd.c:2359-2412, :2435-2482, :2517-2570).  `left` is the number of
 * unread bits; it may go negative (reads past the start return zeros), which
 * is the reference's BIT_DStream_overflow state. */
typedef struct { const u8* base; long left; } rbits;

static int rb_init(rbits* b, const u8* src, size_t n)
{
    if (n == 0) return ZO_ERR_SRCSIZE_WRONG;
    if (src[n - 1] == 0) return ZO_ERR_CORRUPTION;      /* end mark missing */
    b->base = src;
    b->left = (long)(n - 1) * 8 + hibit(src[n - 1]);     /* bits below the end mark */
    return 0;
}
/* peek nb (<=32) bits below the cursor without consuming */
static u32 rb_peek(const rbits* b, int nb)
{
    u64 acc = 0; long lo = b->left - nb;                 /* lowest bit wanted */
    int i;
    if (nb == 0) return 0;
    /* gather bytes covering [lo, lo+nb) ; bits at negative positions are 0 */
    for (i = 0; i < 6; i++) {
        long byte = (lo >> 3) + i;                       /* arithmetic shift: floor */
        u64 v = (byte >= 0 && byte * 8 < b->left) ? b->base[byte] : 0;
        acc |= v << (8 * i);
    }
    return (u32)((acc >> (lo & 7)) & ((1ULL << nb) - 1));
}
static u32 rb_read(rbits* b, int nb) { u32 v = rb_peek(b, nb); b->left -= nb; return v; }

/* --------------------------------------------------------------- FSE */
/* restates FSE_readNCount_body, zstd/zstd.c:3256-3413 : returns bytes used or <0 */
static int fse_read_ncount(short* norm, u32* max_sym, u32* table_log,
                           const u8* src, size_t n)
{
    /* forward LSB-first bit cursor over src, zero padded */
    size_t bitpos = 0;
    u32 const max_sv1 = *max_sym + 1;
    u32 sym = 0; int remaining, threshold, nbits, prev0 = 0;
#define PEEK(nb) ({ u64 a_ = 0; size_t by_ = bitpos >> 3; int i_;               \
        for (i_ = 0; i_ < 5; i_++) a_ |= (u64)(by_ + i_ < n ? src[by_ + i_] : 0) << (8 * i_); \
        (u32)((a_ >> (bitpos & 7)) & ((1ULL << (nb)) - 1)); })
    if (n == 0) return -ZO_ERR_SRCSIZE_WRONG;
    memset(norm, 0, max_sv1 * sizeof(short));
    nbits = (int)PEEK(4) + 5; bitpos += 4;
    if (nbits > 15) return -ZO_ERR_TABLELOG_TOO_LARGE;
    *table_log = (u32)nbits;
    remaining = (1 << nbits) + 1; threshold = 1 << nbits; nbits++;
    for (;;) {
        if (prev0) {
            /* 2-bit repeat codes: 0b11 = three more zero symbols, keep going */
            for (;;) {
                u32 r = PEEK(2); bitpos += 2; sym += r;
                if (r != 3) break;
                if (bitpos > 8 * n + 64) return -ZO_ERR_CORRUPTION;
            }
            if (sym >= max_sv1) break;
        }
        {   int const max = (2 * threshold - 1) - remaining;
            int count; u32 low = PEEK(nbits - 1);
            if ((int)low < max) { count = (int)low; bitpos += (size_t)(nbits - 1); }
            else { count = (int)PEEK(nbits); if (count >= threshold) count -= max; bitpos += (size_t)nbits; }
            count--;
            remaining -= count < 0 ? -count : count;
            norm[sym++] = (short)count;
            prev0 = !count;
            if (remaining < threshold) {
                if (remaining <= 1) break;
                nbits = hibit((u32)remaining) + 1; threshold = 1 << (nbits - 1);
            }
            if (sym >= max_sv1) break;
        }
    }
#undef PEEK
    if (remaining != 1) return -ZO_ERR_CORRUPTION;
    if (sym > max_sv1) return -ZO_ERR_MAXSYMBOL_TOO_SMALL;
    if (bitpos > 8 * n) return -ZO_ERR_CORRUPTION;
    *max_sym = sym - 1;
    return (int)((bitpos + 7) >> 3);
}

typedef struct { u16 next; u8 nb; u8 add_bits; u32 base; } fse_cell;   /* ZSTD_seqSymbol, zstd/zstd.c:41301 */
typedef struct { fse_cell cell[512]; int log; } fse_table;

/* restates the symbol spreading + state assignment shared by
 * FSE_buildDTable_internal (zstd/zstd.c:3680) and ZSTD_buildFSETable_body (:46118).
 * sym_of[] receives the symbol of every cell; next/nb are filled in cell[]. */
static void fse_spread(fse_table* t, u8* sym_of, const short* norm, u32 max_sym, int log)
{
    u32 const size = 1u << log, mask = size - 1, step = (size >> 1) + (size >> 3) + 3;
    u32 high = size - 1, s, pos = 0, u; u16 next[256];
    for (s = 0; s <= max_sym; s++) {
        if (norm[s] == -1) { sym_of[high--] = (u8)s; next[s] = 1; }
        else next[s] = (u16)norm[s];
    }
    for (s = 0; s <= max_sym; s++) {
        int i;
        for (i = 0; i < norm[s]; i++) {
            sym_of[pos] = (u8)s;
            do pos = (pos + step) & mask; while (pos > high);
        }
    }
    for (u = 0; u < size; u++) {
        u32 const x = next[sym_of[u]]++;
        t->cell[u].nb = (u8)(log - hibit(x));
        t->cell[u].next = (u16)((x << t->cell[u].nb) - size);
    }
    t->log = log;
}

static const u8 LL_bits_tab[36] = {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,1,1,1,1,2,2,3,3,4,6,7,8,9,10,11,12,13,14,15,16};
static const u8 ML_bits_tab[53] = {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,
                                   1,1,1,1,2,2,3,3,4,4,5,7,8,9,10,11,12,13,14,15,16};
static const short LL_defnorm[36] = {4,3,2,2,2,2,2,2,2,2,2,2,2,1,1,1,2,2,2,2,2,2,2,2,2,3,2,1,1,1,1,1,-1,-1,-1,-1};
static const short ML_defnorm[53] = {1,4,3,2,2,2,2,2,2,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,
                                     1,1,1,1,1,1,1,1,1,1,1,1,1,1,-1,-1,-1,-1,-1,-1,-1};
static const short OF_defnorm[29] = {1,1,1,1,1,1,2,2,2,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,-1,-1,-1,-1,-1};

/* baselines: LL_base/ML_base/OF_base, zstd/zstd.c:41266-41290, derived from the bit counts */
static u32 ll_base(u32 c) { u32 b = 0, i; for (i = 0; i < c; i++) b += 1u << LL_bits_tab[i]; return b; }
static u32 ml_base(u32 c) { u32 b = 3, i; for (i = 0; i < c; i++) b += 1u << ML_bits_tab[i]; return b; }
static u32 of_base(u32 c) { return c < 2 ? c : (1u << c) - 3; }

enum { K_LL = 0, K_OF = 1, K_ML = 2 };
static void seq_table_fill(fse_table* t, const u8* sym_of, int kind)
{
    u32 u, size = 1u << t->log;
    for (u = 0; u < size; u++) {
        u32 s = sym_of[u];
        if (kind == K_LL) { t->cell[u].base = ll_base(s); t->cell[u].add_bits = LL_bits_tab[s]; }
        else if (kind == K_ML) { t->cell[u].base = ml_base(s); t->cell[u].add_bits = ML_bits_tab[s]; }
        else { t->cell[u].base = of_base(s); t->cell[u].add_bits = (u8)s; }
    }
}
static void seq_table_build(fse_table* t, const short* norm, u32 max_sym, int log, int kind)
{
    u8 sym_of[512];
    fse_spread(t, sym_of, norm, max_sym, log);
    seq_table_fill(t, sym_of, kind);
}
/* restates ZSTD_buildSeqTable_rle, zstd/zstd.c:46100 */
static void seq_table_rle(fse_table* t, u32 sym, int kind)
{
    u8 s = (u8)sym; t->log = 0; t->cell[0].nb = 0; t->cell[0].next = 0;
    seq_table_fill(t, &s, kind);
}

/* ---------------------------------------------------------- Huffman */
typedef struct { u8 sym[4096]; u8 nb[4096]; int log; } huf_table;

/* restates FSE_decompress_wksp_body + FSE_decompress_usingDTable_generic
 * (zstd/zstd.c:3865, :3785) for the Huffman weight stream */
static int fse_decode_weights(u8* out, size_t cap, const u8* src, size_t n)
{
    short norm[256]; u32 max_sym = 255, log; fse_table t; u8 sym_of[512]; rbits b;
    size_t o = 0; u32 s1, s2; int r;
    int used = fse_read_ncount(norm, &max_sym, &log, src, n);
    if (used < 0) return used;
    if (log > 6) return -ZO_ERR_TABLELOG_TOO_LARGE;
    fse_spread(&t, sym_of, norm, max_sym, (int)log);
    if ((r = rb_init(&b, src + used, n - (size_t)used)) != 0) return -r;
    s1 = rb_read(&b, (int)log); s2 = rb_read(&b, (int)log);
    if (b.left < 0) return -ZO_ERR_CORRUPTION;
    for (;;) {
        if (o + 2 > cap) return -ZO_ERR_DSTSIZE_TOO_SMALL;
        out[o++] = sym_of[s1]; s1 = t.cell[s1].next + rb_read(&b, t.cell[s1].nb);
        if (b.left < 0) { out[o++] = sym_of[s2]; break; }
        if (o + 2 > cap) return -ZO_ERR_DSTSIZE_TOO_SMALL;
        out[o++] = sym_of[s2]; s2 = t.cell[s2].next + rb_read(&b, t.cell[s2].nb);
        if (b.left < 0) { out[o++] = sym_of[s1]; break; }
    }
    return (int)o;
}

/* restates HUF_readStats_body (zstd/zstd.c:3457-3521) + the cell layout of
 * HUF_readDTableX1_wksp (:39651-39783).  returns header bytes or <0 */
static int huf_read_table(huf_table* t, const u8* src, size_t n)
{
    u8 w[256]; u32 rank[16] = {0}; u32 nsym, total = 0, i, log; size_t hdr;
    if (n == 0) return -ZO_ERR_SRCSIZE_WRONG;
    if (src[0] >= 128) {                      /* 4-bit weights, high nibble first */
        nsym = (u32)src[0] - 127; hdr = (nsym + 1) / 2;
        if (hdr + 1 > n) return -ZO_ERR_SRCSIZE_WRONG;
        for (i = 0; i < nsym; i++) w[i] = (i & 1) ? (src[1 + i / 2] & 15) : (src[1 + i / 2] >> 4);
    } else {                                  /* FSE-compressed weights */
        int r; hdr = src[0];
        if (hdr + 1 > n) return -ZO_ERR_SRCSIZE_WRONG;
        r = fse_decode_weights(w, 255, src + 1, hdr);
        if (r < 0) return r;
        nsym = (u32)r;
    }
    for (i = 0; i < nsym; i++) {
        if (w[i] > 12) return -ZO_ERR_CORRUPTION;
        rank[w[i]]++; total += (1u << w[i]) >> 1;
    }
    if (total == 0) return -ZO_ERR_CORRUPTION;
    log = (u32)hibit(total) + 1;
    if (log > 12) return -ZO_ERR_CORRUPTION;
    {   u32 rest = (1u << log) - total, last = (u32)hibit(rest) + 1;
        if ((1u << hibit(rest)) != rest) return -ZO_ERR_CORRUPTION;
        w[nsym] = (u8)last; rank[last]++; nsym++;
    }
    if (rank[1] < 2 || (rank[1] & 1)) return -ZO_ERR_CORRUPTION;
    /* cells: weights ascending, symbols ascending within a weight, 2^(w-1) cells each */
    {   u32 wt, pos = 0;
        for (wt = 1; wt <= log; wt++) {
            u32 len = (1u << wt) >> 1;
            for (i = 0; i < nsym; i++) if (w[i] == wt) {
                u32 k; for (k = 0; k < len; k++) { t->sym[pos + k] = (u8)i; t->nb[pos + k] = (u8)(log + 1 - wt); }
                pos += len;
            }
        }
    }
    t->log = (int)log;
    return (int)(hdr + 1);
}

/* restates HUF_decompress1X1_usingDTable_internal_body, zstd/zstd.c:39845 */
static int huf_decode_stream(u8* out, size_t n_out, const u8* src, size_t n, const huf_table* t)
{
    rbits b; size_t i; int r = rb_init(&b, src, n);
    if (r) return r == ZO_ERR_SRCSIZE_WRONG ? ZO_ERR_SRCSIZE_WRONG : ZO_ERR_CORRUPTION;
    for (i = 0; i < n_out; i++) {
        u32 v = rb_peek(&b, t->log);
        out[i] = t->sym[v]; b.left -= t->nb[v];
    }
    return b.left == 0 ? 0 : ZO_ERR_CORRUPTION;
}
/* restates HUF_decompress4X1_usingDTable_internal_body, zstd/zstd.c:39868-39964 */
static int huf_decode_4(u8* out, size_t n_out, const u8* src, size_t n, const huf_table* t)
{
    size_t l1, l2, l3, l4, seg; int r;
    if (n < 10 || n_out < 6) return ZO_ERR_CORRUPTION;
    l1 = rd16(src); l2 = rd16(src + 2); l3 = rd16(src + 4);
    if (6 + l1 + l2 + l3 > n) return ZO_ERR_CORRUPTION;
    l4 = n - 6 - l1 - l2 - l3; seg = (n_out + 3) / 4;
    if (seg * 3 > n_out) return ZO_ERR_CORRUPTION;
    if ((r = huf_decode_stream(out, seg, src + 6, l1, t))) return ZO_ERR_CORRUPTION;
    if ((r = huf_decode_stream(out + seg, seg, src + 6 + l1, l2, t))) return ZO_ERR_CORRUPTION;
    if ((r = huf_decode_stream(out + 2 * seg, seg, src + 6 + l1 + l2, l3, t))) return ZO_ERR_CORRUPTION;
    if ((r = huf_decode_stream(out + 3 * seg, n_out - 3 * seg, src + 6 + l1 + l2 + l3, l4, t))) return ZO_ERR_CORRUPTION;
    return 0;
}

/* ------------------------------------------------------ frame decoder */
typedef struct {
    huf_table huf; int huf_valid;
    fse_table ll, of, ml; int fse_valid;      /* "repeat" tables, zstd/zstd.c:46301-46310 */
    u32 rep[3];
    const u8* dict; size_t dict_size;         /* raw content treated as history */
    u32 dict_id;
    u8* lits;                                 /* 128 KiB literal buffer */
    zo_trace* trace;
} zo_ctx;

static void default_tables(zo_ctx* c)
{
    seq_table_build(&c->ll, LL_defnorm, 35, 6, K_LL);   /* zstd/zstd.c:15622-15659 */
    seq_table_build(&c->of, OF_defnorm, 28, 5, K_OF);
    seq_table_build(&c->ml, ML_defnorm, 52, 6, K_ML);
}

/* restates ZSTD_loadDEntropy (zstd/zstd.c:44673-44757) + ZSTD_decompress_insertDictionary (:44760) */
static int load_dict(zo_ctx* c, const u8* dict, size_t n)
{
    const u8* p = dict; const u8* end = dict + n; int r; u32 i;
    c->dict = dict; c->dict_size = n; c->dict_id = 0;
    if (n < 8 || rd32(dict) != ZO_MAGIC_DICT) return 0;          /* raw-content dictionary */
    c->dict_id = rd32(dict + 4); p += 8;
    if ((r = huf_read_table(&c->huf, p, (size_t)(end - p))) < 0) return ZO_ERR_DICT_CORRUPTED;
    p += r;
    {   short norm[64]; u32 max, log;
        max = 31; if ((r = fse_read_ncount(norm, &max, &log, p, (size_t)(end - p))) < 0 || log > 8) return ZO_ERR_DICT_CORRUPTED;
        seq_table_build(&c->of, norm, max, (int)log, K_OF); p += r;
        max = 52; if ((r = fse_read_ncount(norm, &max, &log, p, (size_t)(end - p))) < 0 || log > 9) return ZO_ERR_DICT_CORRUPTED;
        seq_table_build(&c->ml, norm, max, (int)log, K_ML); p += r;
        max = 35; if ((r = fse_read_ncount(norm, &max, &log, p, (size_t)(end - p))) < 0 || log > 9) return ZO_ERR_DICT_CORRUPTED;
        seq_table_build(&c->ll, norm, max, (int)log, K_LL); p += r;
    }
    if (p + 12 > end) return ZO_ERR_DICT_CORRUPTED;
    {   size_t content = (size_t)(end - (p + 12));
        for (i = 0; i < 3; i++) {
            u32 rep = rd32(p + 4 * i);
            if (rep == 0 || rep > content) return ZO_ERR_DICT_CORRUPTED;
            c->rep[i] = rep;
        }
    }
    p += 12;
    c->dict = p; c->dict_size = (size_t)(end - p);
    c->huf_valid = 1; c->fse_valid = 1;
    return 0;
}

/* restates ZSTD_decodeLiteralsBlock, zstd/zstd.c:45767-45973.
 * returns section size or <0; *lit / *n_lit describe the regenerated literals */
static long decode_literals(zo_ctx* c, const u8* src, size_t n, size_t block_max,
                            const u8** lit, size_t* n_lit)
{
    u32 type, sf; size_t hdr, regen, csize;
    if (n < 2) return -ZO_ERR_CORRUPTION;                  /* MIN_CBLOCK_SIZE */
    type = src[0] & 3; sf = (src[0] >> 2) & 3;
    if (type == 0 || type == 1) {                          /* raw / RLE */
        if (sf == 1) { hdr = 2; regen = rd16(src) >> 4; }
        else if (sf == 3) { hdr = 3; if (n < 3) return -ZO_ERR_CORRUPTION; regen = rd24(src) >> 4; }
        else { hdr = 1; regen = src[0] >> 3; }
        if (regen > block_max) return -ZO_ERR_CORRUPTION;
        if (type == 0) {
            if (hdr + regen > n) return -ZO_ERR_CORRUPTION;
            *lit = src + hdr; *n_lit = regen; return (long)(hdr + regen);
        }
        if (hdr + 1 > n) return -ZO_ERR_CORRUPTION;
        memset(c->lits, src[hdr], regen);
        *lit = c->lits; *n_lit = regen; return (long)(hdr + 1);
    }
    /* Huffman (2) or treeless/repeat (3) */
    if (type == 3 && !c->huf_valid) return -ZO_ERR_DICT_CORRUPTED;
    if (n < 5) return -ZO_ERR_CORRUPTION;
    {   u32 lhc = rd32(src); int single = 0; int r;
        if (sf < 2) { single = (sf == 0); hdr = 3; regen = (lhc >> 4) & 0x3FF; csize = (lhc >> 14) & 0x3FF; }
        else if (sf == 2) { hdr = 4; regen = (lhc >> 4) & 0x3FFF; csize = lhc >> 18; }
        else { hdr = 5; regen = (lhc >> 4) & 0x3FFFF; csize = (lhc >> 22) + ((size_t)src[4] << 10); }
        if (regen > block_max) return -ZO_ERR_CORRUPTION;
        if (!single && regen < 6) return -ZO_ERR_LITERALS_HEADER_WRONG;
        if (csize + hdr > n) return -ZO_ERR_CORRUPTION;
        {   const u8* p = src + hdr; size_t left = csize;
            if (type == 2) {
                r = huf_read_table(&c->huf, p, left);
                if (r < 0 || (size_t)r >= left) return -ZO_ERR_CORRUPTION;
                p += r; left -= (size_t)r;
            }
            if (regen == 0) return -ZO_ERR_CORRUPTION;     /* HUF: dstSize==0 is an error */
            r = single ? huf_decode_stream(c->lits, regen, p, left, &c->huf)
                       : huf_decode_4(c->lits, regen, p, left, &c->huf);
            if (r) return -ZO_ERR_CORRUPTION;
        }
        c->huf_valid = 1;
        *lit = c->lits; *n_lit = regen;
        return (long)(hdr + csize);
    }
}

/* restates ZSTD_buildSeqTable, zstd/zstd.c:46280-46326 */
static long build_seq_table(fse_table* t, u32 mode, u32 max_sym, u32 max_log, int kind,
                            const u8* src, size_t n, int repeat_ok)
{
    switch (mode) {
    case 1: if (n == 0) return -ZO_ERR_SRCSIZE_WRONG;
            if (src[0] > max_sym) return -ZO_ERR_CORRUPTION;
            seq_table_rle(t, src[0], kind); return 1;
    case 0: if (kind == K_LL) seq_table_build(t, LL_defnorm, 35, 6, K_LL);
            else if (kind == K_OF) seq_table_build(t, OF_defnorm, 28, 5, K_OF);
            else seq_table_build(t, ML_defnorm, 52, 6, K_ML);
            return 0;
    case 3: return repeat_ok ? 0 : -ZO_ERR_CORRUPTION;
    default: {
            short norm[64]; u32 log; int r = fse_read_ncount(norm, &max_sym, &log, src, n);
            if (r < 0 || log > max_log) return -ZO_ERR_CORRUPTION;
            seq_table_build(t, norm, max_sym, (int)log, kind); return r; }
    }
}

/* restates ZSTD_decompressBlock_internal (zstd/zstd.c:47699) =
 * ZSTD_decodeSeqHeaders (:46328) + ZSTD_decompressSequences_body (:47248) with
 * ZSTD_decodeSequence (:46862) and ZSTD_execSequence (:46634) */
static long decode_block(zo_ctx* c, u8* out_base, size_t out_pos, size_t out_cap,
                         const u8* src, size_t n, size_t block_max)
{
    const u8* lit = NULL; size_t n_lit = 0, lit_pos = 0; size_t op = out_pos;
    long r = decode_literals(c, src, n, block_max, &lit, &n_lit);
    const u8* ip; const u8* iend = src + n; u32 nseq;
    zo_trace* tr = c->trace;
    if (n > block_max) return -ZO_ERR_SRCSIZE_WRONG;
    if (r < 0) return r;
    ip = src + r;
    if (ip >= iend) return -ZO_ERR_SRCSIZE_WRONG;           /* MIN_SEQUENCES_SIZE */
    nseq = *ip++;
    if (nseq > 0x7F) {
        if (nseq == 0xFF) { if (ip + 2 > iend) return -ZO_ERR_SRCSIZE_WRONG; nseq = rd16(ip) + 0x7F00; ip += 2; }
        else { if (ip >= iend) return -ZO_ERR_SRCSIZE_WRONG; nseq = ((nseq - 0x80) << 8) + *ip++; }
    }
    if (tr && tr->n_blocks < tr->block_cap) { tr->block_nseq[tr->n_blocks] = nseq; tr->block_nlit[tr->n_blocks] = (u32)n_lit; }
    if (tr) { tr->n_blocks++; if (tr->n_lits + n_lit <= tr->lit_cap) memcpy(tr->lits + tr->n_lits, lit, n_lit); tr->n_lits += n_lit; }
    if (nseq == 0) {
        if (ip != iend) return -ZO_ERR_CORRUPTION;
    } else {
        u32 modes; rbits b; u32 sll, sof, sml; u32 i; int e;
        if (ip + 1 > iend) return -ZO_ERR_SRCSIZE_WRONG;
        modes = *ip++;
        if (modes & 3) return -ZO_ERR_CORRUPTION;
        r = build_seq_table(&c->ll, modes >> 6, 35, 9, K_LL, ip, (size_t)(iend - ip), c->fse_valid);
        if (r < 0) return -ZO_ERR_CORRUPTION; ip += r;
        r = build_seq_table(&c->of, (modes >> 4) & 3, 31, 8, K_OF, ip, (size_t)(iend - ip), c->fse_valid);
        if (r < 0) return -ZO_ERR_CORRUPTION; ip += r;
        r = build_seq_table(&c->ml, (modes >> 2) & 3, 52, 9, K_ML, ip, (size_t)(iend - ip), c->fse_valid);
        if (r < 0) return -ZO_ERR_CORRUPTION; ip += r;
        c->fse_valid = 1;
        if ((e = rb_init(&b, ip, (size_t)(iend - ip))) != 0) return -ZO_ERR_CORRUPTION;
        sll = rb_read(&b, c->ll.log); sof = rb_read(&b, c->of.log); sml = rb_read(&b, c->ml.log);
        for (i = 0; i < nseq; i++) {
            const fse_cell* L = &c->ll.cell[sll]; const fse_cell* O = &c->of.cell[sof]; const fse_cell* M = &c->ml.cell[sml];
            u32 ll = L->base, ml = M->base, off;
            if (O->add_bits > 1) {
                off = O->base + rb_read(&b, O->add_bits);
                c->rep[2] = c->rep[1]; c->rep[1] = c->rep[0]; c->rep[0] = off;
            } else {
                u32 const ll0 = (L->base == 0);
                if (O->add_bits == 0) {
                    off = c->rep[ll0]; c->rep[1] = c->rep[!ll0]; c->rep[0] = off;
                } else {
                    u32 idx = O->base + ll0 + rb_read(&b, 1);
                    u32 tmp = (idx == 3) ? c->rep[0] - 1 : c->rep[idx];
                    if (tmp == 0) tmp = 0xFFFFFFFFu;          /* invalid -> caught below */
                    if (idx != 1) c->rep[2] = c->rep[1];
                    c->rep[1] = c->rep[0]; c->rep[0] = off = tmp;
                }
            }
            ml += rb_read(&b, M->add_bits);
            ll += rb_read(&b, L->add_bits);
            if (i + 1 < nseq) {
                sll = L->next + rb_read(&b, L->nb);
                sml = M->next + rb_read(&b, M->nb);
                sof = O->next + rb_read(&b, O->nb);
            }
            if (tr) { if (tr->n_seqs < tr->seq_cap) { tr->seqs[tr->n_seqs].lit_len = ll; tr->seqs[tr->n_seqs].match_len = ml; tr->seqs[tr->n_seqs].offset = off; } tr->n_seqs++; }
            /* execute */
            if ((size_t)ll + ml > out_cap - op) return -ZO_ERR_DSTSIZE_TOO_SMALL;
            if (ll > n_lit - lit_pos) return -ZO_ERR_CORRUPTION;
            memcpy(out_base + op, lit + lit_pos, ll); op += ll; lit_pos += ll;
            if ((size_t)off > op + c->dict_size) return -ZO_ERR_CORRUPTION;
            {   u32 k;
                for (k = 0; k < ml; k++, op++) {
                    if (off > op) out_base[op] = c->dict[c->dict_size - (off - op)];
                    else out_base[op] = out_base[op - off];
                }
            }
        }
        if (b.left != 0) return -ZO_ERR_CORRUPTION;
    }
    {   size_t last = n_lit - lit_pos;
        if (last > out_cap - op) return -ZO_ERR_DSTSIZE_TOO_SMALL;
        memcpy(out_base + op, lit + lit_pos, last); op += last;
    }
    if (op - out_pos > block_max) return -ZO_ERR_CORRUPTION;
    return (long)(op - out_pos);
}

int zo_get_frame_header(zo_frame_header* h, const void* vsrc, size_t n)
{
    /* restates ZSTD_getFrameHeader_advanced, zstd/zstd.c:43668-43778 */
    const u8* src = (const u8*)vsrc; u32 magic; u8 fhd; size_t pos = 5, need;
    static const u8 did_size[4] = {0, 1, 2, 4}; static const u8 fcs_size[4] = {0, 2, 4, 8};
    memset(h, 0, sizeof(*h));
    if (n < 5) {                      /* too short: still reject a wrong magic prefix (:43680-43697) */
        u8 zm[4] = {0x28, 0xB5, 0x2F, 0xFD}, sm[4] = {0x50, 0x2A, 0x4D, 0x18}; size_t k = n < 4 ? n : 4;
        if (n && memcmp(src, zm, k) != 0 && !((k < 1 || (src[0] & 0xF0) == sm[0]) && (k < 2 || memcmp(src + 1, sm + 1, k - 1) == 0)))
            return ZO_ERR_PREFIX_UNKNOWN;
        return ZO_ERR_SRCSIZE_WRONG;
    }
    magic = rd32(src);
    if (magic != ZO_MAGIC) {
        if ((magic & 0xFFFFFFF0u) == ZO_MAGIC_SKIP) {
            if (n < 8) return ZO_ERR_SRCSIZE_WRONG;
            h->is_skippable = 1; h->skippable_size = rd32(src + 4); h->header_size = 8;
            h->content_size = 0; return 0;
        }
        return ZO_ERR_PREFIX_UNKNOWN;
    }
    fhd = src[4];
    h->single_segment = (fhd >> 5) & 1;
    need = 5 + (h->single_segment ? 0 : 1) + did_size[fhd & 3] + fcs_size[fhd >> 6] + ((h->single_segment && !(fhd >> 6)) ? 1 : 0);
    if (n < need) return ZO_ERR_SRCSIZE_WRONG;
    h->header_size = (u32)need;
    if (fhd & 0x08) return ZO_ERR_FRAMEPARAM_UNSUPPORTED;
    h->has_checksum = (fhd >> 2) & 1;
    h->content_size = UINT64_MAX;
    if (!h->single_segment) {
        u8 wl = src[pos++]; u32 wlog = (wl >> 3) + 10;
        if (wlog > ZO_WINDOWLOG_MAX) return ZO_ERR_WINDOW_TOO_LARGE;
        h->window_size = 1ULL << wlog; h->window_size += (h->window_size >> 3) * (wl & 7);
    }
    switch (fhd & 3) { case 1: h->dict_id = src[pos]; pos += 1; break;
                       case 2: h->dict_id = rd16(src + pos); pos += 2; break;
                       case 3: h->dict_id = rd32(src + pos); pos += 4; break; default: break; }
    switch (fhd >> 6) { case 0: if (h->single_segment) h->content_size = src[pos]; break;
                        case 1: h->content_size = rd16(src + pos) + 256; break;
                        case 2: h->content_size = rd32(src + pos); break;
                        default: h->content_size = rd64(src + pos); break; }
    if (h->single_segment) h->window_size = h->content_size;
    return 0;
}

size_t zo_find_frame_compressed_size(const void* vsrc, size_t n, int* err)
{
    /* restates ZSTD_findFrameSizeInfo, zstd/zstd.c:43905 */
    const u8* src = (const u8*)vsrc; zo_frame_header h; size_t pos; int e = zo_get_frame_header(&h, src, n);
    *err = e; if (e) return 0;
    if (h.is_skippable) { if ((size_t)h.skippable_size + 8 > n) { *err = ZO_ERR_SRCSIZE_WRONG; return 0; } return (size_t)h.skippable_size + 8; }
    pos = h.header_size;
    for (;;) {
        u32 bh, type, bsize;
        if (pos + 3 > n) { *err = ZO_ERR_SRCSIZE_WRONG; return 0; }
        bh = rd24(src + pos); type = (bh >> 1) & 3; bsize = bh >> 3; pos += 3;
        if (type == 3) { *err = ZO_ERR_CORRUPTION; return 0; }
        if (type == 1) bsize = 1;
        if (pos + bsize > n) { *err = ZO_ERR_SRCSIZE_WRONG; return 0; }
        pos += bsize;
        if (bh & 1) break;
    }
    if (h.has_checksum) { if (pos + 4 > n) { *err = ZO_ERR_SRCSIZE_WRONG; return 0; } pos += 4; }
    return pos;
}

size_t zo_decompress_frame(void* vdst, size_t dst_cap, const void* vsrc, size_t n,
                           const void* dict, size_t dict_size,
                           size_t* consumed, zo_trace* trace, int* err)
{
    /* restates ZSTD_decompressFrame, zstd/zstd.c:44174-44288 */
    const u8* src = (const u8*)vsrc; u8* dst = (u8*)vdst; zo_frame_header h; zo_ctx* c;
    size_t pos, op = 0, block_max; int e;
    *err = 0; if (consumed) *consumed = 0;
    if ((e = zo_get_frame_header(&h, src, n)) != 0) { *err = e; return 0; }
    if (n < 6 + 3) { *err = ZO_ERR_SRCSIZE_WRONG; return 0; }      /* :44188 */
    if (h.is_skippable) { *err = ZO_ERR_PREFIX_UNKNOWN; return 0; }
    if (n < (size_t)h.header_size + 3) { *err = ZO_ERR_SRCSIZE_WRONG; return 0; }
    c = (zo_ctx*)calloc(1, sizeof(zo_ctx));
    c->lits = (u8*)malloc(ZO_BLOCK_MAX + 32);
    c->rep[0] = 1; c->rep[1] = 4; c->rep[2] = 8;                 /* zstd/zstd.c:15561 */
    c->trace = trace;
    default_tables(c);
    if (dict && dict_size) {
        if ((e = load_dict(c, (const u8*)dict, dict_size)) != 0) { *err = e; goto done; }
        if (h.dict_id && c->dict_id && h.dict_id != c->dict_id) { *err = ZO_ERR_DICT_WRONG; goto done; }
    }
    block_max = h.window_size < ZO_BLOCK_MAX ? (size_t)h.window_size : ZO_BLOCK_MAX;
    pos = h.header_size;
    for (;;) {
        u32 bh, type, bsize; int last; long r;
        if (pos + 3 > n) { *err = ZO_ERR_SRCSIZE_WRONG; goto done; }
        bh = rd24(src + pos); pos += 3;
        last = bh & 1; type = (bh >> 1) & 3; bsize = bh >> 3;
        if (type == 3) { *err = ZO_ERR_CORRUPTION; goto done; }
        if (type == 1) {                                          /* RLE, ZSTD_setRleBlock :44130 */
            if (pos + 1 > n) { *err = ZO_ERR_SRCSIZE_WRONG; goto done; }
            if (bsize > block_max) { *err = ZO_ERR_CORRUPTION; goto done; }
            if (bsize > dst_cap - op) { *err = ZO_ERR_DSTSIZE_TOO_SMALL; goto done; }
            memset(dst + op, src[pos], bsize); op += bsize; pos += 1;
        } else {
            if (pos + bsize > n) { *err = ZO_ERR_SRCSIZE_WRONG; goto done; }
            if (type == 0) {                                      /* raw, ZSTD_copyRawBlock :44117 */
                if (bsize > block_max) { *err = ZO_ERR_CORRUPTION; goto done; }
                if (bsize > dst_cap - op) { *err = ZO_ERR_DSTSIZE_TOO_SMALL; goto done; }
                memcpy(dst + op, src + pos, bsize); op += bsize;
            } else {
                r = decode_block(c, dst, op, dst_cap, src + pos, bsize, block_max);
                if (r < 0) { *err = (int)-r; goto done; }
                op += (size_t)r;
            }
            pos += bsize;
        }
        if (last) break;
    }
    if (h.content_size != UINT64_MAX && (u64)op != h.content_size) { *err = ZO_ERR_CORRUPTION; goto done; }
    if (h.has_checksum) {
        if (pos + 4 > n) { *err = ZO_ERR_CHECKSUM_WRONG; goto done; }
        if ((u32)zo_xxh64(dst, op, 0) != rd32(src + pos)) { *err = ZO_ERR_CHECKSUM_WRONG; goto done; }
        pos += 4;
    }
    if (consumed) *consumed = pos;
done:
    free(c->lits); free(c);
    return *err ? 0 : op;
}

size_t zo_decompress(void* dst, size_t dst_cap, const void* vsrc, size_t n,
                     const void* dict, size_t dict_size, int* err)
{
    /* restates ZSTD_decompressMultiFrame, zstd/zstd.c:44291-44390 */
    const u8* src = (const u8*)vsrc; size_t op = 0; int more = 0;
    *err = 0;
    while (n >= 5 || (n > 0 && !more)) {
        zo_frame_header h; size_t used, w; int e;
        if (n >= 8 && (rd32(src) & 0xFFFFFFF0u) == ZO_MAGIC_SKIP) {
            if ((e = zo_get_frame_header(&h, src, n)) != 0) { *err = e; return 0; }
            if ((size_t)h.skippable_size + 8 > n) { *err = ZO_ERR_SRCSIZE_WRONG; return 0; }
            src += h.skippable_size + 8; n -= h.skippable_size + 8; more = 1; continue;
        }
        w = zo_decompress_frame((u8*)dst + op, dst_cap - op, src, n, dict, dict_size, &used, NULL, &e);
        if (e) { *err = (e == ZO_ERR_PREFIX_UNKNOWN && more) ? ZO_ERR_SRCSIZE_WRONG : e; return 0; }
        op += w; src += used; n -= used; more = 1;
    }
    if (n) { *err = ZO_ERR_SRCSIZE_WRONG; return 0; }
    return op;
}
