// zb_common.cuh -- shared device-side definitions of the B200 zstd batch codec.
//
// Format constants are RFC 8878's (the reference holds them at zstd/zstd.c:15596-15662
// and :41266-41290); everything else here is this project's own design.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

typedef uint8_t  u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;

// ---- status codes: numeric values follow zstd_errors.h so host code can print
// ---- the reference's own error strings (c-ext/decompressor.c:1328-1371)
enum : u32 {
    ZB_OK = 0,
    ZB_E_GENERIC = 1,
    ZB_E_PREFIX_UNKNOWN = 10,
    ZB_E_FRAMEPARAM_UNSUPPORTED = 14,
    ZB_E_WINDOW_TOO_LARGE = 16,
    ZB_E_CORRUPTION = 20,
    ZB_E_CHECKSUM_WRONG = 22,
    ZB_E_LITERALS_HEADER_WRONG = 24,
    ZB_E_DICT_CORRUPTED = 30,
    ZB_E_DICT_WRONG = 32,
    ZB_E_TABLELOG_TOO_LARGE = 44,
    ZB_E_MAXSYMBOL_TOO_SMALL = 48,
    ZB_E_MEMORY = 64,
    ZB_E_DSTSIZE_TOO_SMALL = 70,
    ZB_E_SRCSIZE_WRONG = 72,
    // ours (python-zstandard worker errors, c-ext/decompressor.c:911-917)
    ZB_E_UNKNOWN_SIZE = 200,
    ZB_E_SIZE_MISMATCH = 201,
};

#define ZB_MAGIC       0xFD2FB528u
#define ZB_MAGIC_DICT  0xEC30A437u
#define ZB_MAGIC_SKIP  0x184D2A50u
#define ZB_BLOCK_MAX   (128u << 10)
#define ZB_CONTENT_UNKNOWN 0xFFFFFFFFFFFFFFFFull

struct ZbSegment { u64 offset, length; };       // == BufferSegment, c-ext/python-zstandard.h:307-313

// result of the frame scan (one per frame)
struct ZbFrameInfo {
    u64 content_size;     // from the header, ZB_CONTENT_UNKNOWN if absent
    u64 n_seq_rec;        // sequence records needed: sum(nbSeq + 1) over compressed blocks
    u64 n_lit;            // literal bytes that must be regenerated into scratch (Huffman coded); 64-bit: an 8 GiB frame can
                          //   hold more than 4 GiB of them
    u32 n_blocks;
    u32 status;
    u32 dict_id;
    u32 flags;            // bit0: checksum present
};

// per-frame placement, produced by the offsets scan
struct ZbFramePlace {
    u64 dst_off;          // byte offset of the frame's output in dst
    u64 dst_cap;          // bytes the frame may write
    u64 blk_off;          // first ZbBlock of the frame
    u64 seq_off;          // first sequence record
    u64 lit_off;          // first literal scratch byte
};

enum : u32 { ZB_BLK_RAW = 0, ZB_BLK_RLE = 1, ZB_BLK_COMPRESSED = 2 };
enum : u32 { ZB_LIT_RAW = 0, ZB_LIT_RLE = 1, ZB_LIT_SCRATCH = 2 };

// one block of a frame, written by the entropy stage for the execute stage
struct ZbBlock {
    u64 out_pos;          // frame-relative start of the block's output
    u64 src_pos;          // raw/rle block: payload position in src.  compressed: literal position
                          //   (in src for ZB_LIT_RAW, in the literal scratch for ZB_LIT_SCRATCH)
    u64 seq_pos;          // first sequence record (absolute index)
    u32 kind;             // ZB_BLK_*
    u32 regen;            // regenerated size of the block
    u32 lit_kind;         // ZB_LIT_*  (ZB_LIT_RLE: byte value in lit_byte)
    u32 n_lit;
    u32 n_seq;
    u32 lit_byte;
};

// sequence record: .x = literal start (block-relative index into the block's literals)
//                  .y = output start of the sequence's literals (block-relative)
//                  .z = match length, .w = match offset (real distance, repcodes resolved)
// A sentinel record {.x = literals consumed, .y = output produced} ends every block.
typedef uint4 ZbSeq;

// FSE decode cell, 32 bits: the information of ZSTD_seqSymbol (zstd/zstd.c:41301-41306) minus the
// baseline, which is looked up from the symbol (off the state chain):
//   bits 0-9 nextState base, 10-13 nbBits, 14-18 nbAdditionalBits, 19-24 symbol code
typedef u32 ZbFseCell;
#define ZB_CELL(next, nb, add, sym) ((u32)(next) | ((u32)(nb) << 10) | ((u32)(add) << 14) | ((u32)(sym) << 19))
#define ZB_CELL_NEXT(c) ((c) & 1023u)
#define ZB_CELL_NB(c)   (((c) >> 10) & 15u)
#define ZB_CELL_ADD(c)  (((c) >> 14) & 31u)
#define ZB_CELL_SYM(c)  ((c) >> 19)

// digested dictionary, device resident (restates what ZSTD_loadDEntropy keeps, zstd/zstd.c:44673-44757)
struct ZbDictDev {
    const u8* content; u32 content_size; u32 dict_id;
    const u16* huf; u32 huf_log; u32 has_entropy;
    const ZbFseCell* ll; const ZbFseCell* of; const ZbFseCell* ml;
    u32 ll_log, of_log, ml_log;
    u32 rep[3];
};

// dictionary digest as the device kernel writes it
struct ZbDictDigest {
    u16 huf[4096]; ZbFseCell ll[512]; ZbFseCell ml[512]; ZbFseCell of[256];
    u32 huf_log, ll_log, of_log, ml_log; u32 rep[3]; u32 dict_id; u32 content_off; u32 status; u32 has_entropy; u32 pad;
    // encoder view of the same entropy tables (ZSTD_loadCEntropy, zstd/zstd.c:28015): normalized counts and code lengths
    short c_norm_ll[36], c_norm_of[32], c_norm_ml[54]; u32 c_max_ll, c_max_of, c_max_ml;
    u8 c_huf_nb[256]; u32 c_huf_max;
};

__device__ __forceinline__ u32 zb_rd16(const u8* p) { return (u32)p[0] | ((u32)p[1] << 8); }
__device__ __forceinline__ u32 zb_rd24(const u8* p) { return zb_rd16(p) | ((u32)p[2] << 16); }
__device__ __forceinline__ u32 zb_rd32(const u8* p) { return zb_rd16(p) | (zb_rd16(p + 2) << 16); }
__device__ __forceinline__ u64 zb_rd64(const u8* p) { return (u64)zb_rd32(p) | ((u64)zb_rd32(p + 4) << 32); }
__device__ __forceinline__ int zb_hibit(u32 v) { return 31 - __clz(v); }

// ---------------------------------------------------------------------------
// Backward bit reader over an arbitrary byte range, built on ALIGNED 32-bit loads and 32-bit
// funnel shifts.  (hi:lo) holds the next unread bits top-aligned; `avail` counts them.  The two
// words below the window are loaded ahead (nx0, nx1) so that global-memory latency is off the
// decode chain.  left() is the number of unread stream bits; it goes negative when the stream is
// over-read (the reference's BIT_DStream_overflow, zstd/zstd.c:2517-2556).  Reads below the stream
// start return the neighbouring bytes (not zeros): harmless, because left() < 0 fails the block.
// ---------------------------------------------------------------------------
struct ZbBitR {
    const u32* w; int widx; u32 hi, lo; int avail; u32 nx0, nx1; int skew_bits;

    __device__ __forceinline__ bool init(const u8* s, u32 n) {
        if (n == 0) return false;
        u32 const last = s[n - 1];
        if (last == 0) return false;
        uintptr_t const a = (uintptr_t)s & ~(uintptr_t)3;
        w = (const u32*)a;
        int const skew = (int)((uintptr_t)s - a);
        skew_bits = skew * 8;
        int const P = (skew + (int)n - 1) * 8 + zb_hibit(last);   // bits from the aligned base up to the end mark
        nx0 = nx1 = 0; lo = 0;
        if (P == 0) { hi = 0; avail = 0; widx = -1; return true; }
        int const wi = (P - 1) >> 5, k = P - wi * 32;              // k in 1..32 valid bits in the top word
        hi = w[wi] << (32 - k); avail = k; widx = wi - 1;
        if (widx >= 0) nx0 = w[widx];
        if (widx >= 1) nx1 = w[widx - 1];
        refill();
        return true;
    }
    // branch-free: lanes of a warp refill at different moments, so the body is predicated, not branched
    __device__ __forceinline__ void refill() {
        bool const r = (avail <= 32) & (widx >= 0);                // lo is empty when r holds
        u32 const a = (u32)avail;
        u32 const add_hi = __funnelshift_rc(nx0, 0u, a);           // nx0 >> avail        (0 when avail == 32)
        u32 const new_lo = __funnelshift_lc(0u, nx0, 32u - a);     // nx0 << (32 - avail) (0 when avail == 0)
        hi |= r ? add_hi : 0u;
        lo = r ? new_lo : lo;
        avail += r ? 32 : 0;
        widx -= r ? 1 : 0;
        nx0 = r ? nx1 : nx0;
        if (r && widx >= 1) nx1 = w[widx - 1];
    }
    __device__ __forceinline__ u32 peek(u32 nb) const { return __funnelshift_rc(hi, 0u, 32u - nb); }   // nb in 0..32
    __device__ __forceinline__ void skip(u32 nb) {                                                      // nb in 0..32
        hi = __funnelshift_lc(lo, hi, nb); lo = __funnelshift_lc(0u, lo, nb); avail -= (int)nb;
    }
    __device__ __forceinline__ u32 read(u32 nb) { u32 const v = peek(nb); skip(nb); return v; }
    __device__ __forceinline__ int left() const { return avail + 32 * (widx + 1) - skew_bits; }
};
