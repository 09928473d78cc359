// zb_encode2.cuh -- the round-2 block compressor: ONE CTA of 1024 threads per SM, the <=128 KiB block resident in
// shared memory (TMA bulk copy), every dependent access (hash tables, candidate verification, parse, FSE states) served
// from shared memory; only streaming intermediates (unit records, compacted sequences) go through the L2-resident
// per-CTA scratch.  Included by zb_encode.cu (it reuses the table builders above).
//
// Stands in for ZSTD_compressBlock_doubleFast (zstd/zstd.c:31039) + ZSTD_entropyCompressSeqStore (:25842) per block:
//   load     cp.async.bulk global -> shared, mbarrier completion
//   links    passes of 16 KiB.  near: a warp per 1 KiB sub-chunk, private 512-slot table (position | 6-bit tag),
//            16 positions per half-step, exact order.  far: one 2^14-slot table keyed on 5 bytes, advanced in rounds
//            of 1024 positions (all threads): lookup -> verify both candidates + distances 1..4 against the input in
//            shared memory (<= 15 bytes) -> one-step lazy with the neighbour's length -> dist[] (0 = no match here)
//   parse    a lane per 132-byte unit walks dist[]: a match may START only in its unit but extends freely (forward
//            to the block end, backward to the unit's anchor); records {start, len, dist} -> scratch
//   stitch   fold of the units' last match ends -> what each unit must drop / front-trim; survivors are compacted;
//            repcode history is an MTF(3) list, so unit summaries (3 most recent distinct offsets) scan exactly and
//            every lane codes its unit's offsets with the true history (ZSTD_updateRep :19971)
//   entropy  tables by four warps (LL, OF, ML, Huffman); FSE state chains run SPECULATIVELY per lane range: a lane
//            warms its states up on the 24 sequences after its range, neighbours compare states and only lanes whose
//            guess was wrong redo their range (states forget their past after ~log/H symbols); bit counts are
//            prefix-scanned and every lane packs its own span; Huffman literals likewise from prefix-scanned lengths
//   assemble literal streams are OR-ed straight into the block's slot, the sequence stream is staged in shared
//            memory and copied; raw / RLE fallbacks
#pragma once

#define Z2_NT       1024
#define Z2_U        132                   // bytes per parse unit: 33 words, so the lanes of a warp start in 32 different banks
#define Z2_Q        16384                 // positions per pass
#define Z2_UPP      125                   // units per pass
#define Z2_MAXU     1000
#define Z2_RMAX     34                    // records per unit (a unit starts at most 33 matches)
#define Z2_FARLOG   14
#define Z2_MAXSEQ   33024
#define Z2_STAGE    57344                 // bytes of staging / literal buffer in shared memory
#define Z2_WARM     24                    // FSE warm-up symbols
#define Z2_SEQ_SMEM 16384                 // sequence records that fit the (by then dead) input buffer

struct Z2Scratch {
    u64 rec[Z2_MAXU * Z2_RMAX];           // unit records: start | len << 17 | dist << 35
    u64 fseq[Z2_MAXSEQ + 8];              // final sequences: ll | (ml - 3) << 17 | offBase << 34
    u8  lit[ZE_BLOCK + 64];               // gathered literals when they do not fit shared memory
    u32 stage[(ZE_BLOCK + 4096) / 4];     // sequence-stream staging when it does not fit shared memory
};

struct Z2Mf { u16 far[1 << Z2_FARLOG]; u16 near[16][512]; u16 dist[Z2_Q + 16]; };
struct Z2Ent {
    __align__(16) u32 stage[Z2_STAGE / 4];        // literals (gathered), later the sequence bitstream
    ZeCTable ct[4];                       // LL, OF, ML, Huffman-weight table
    ZeHuf huf;
    u32 wk[1600];
    u8 tmp_sym[4][512];
    u8 huf_tbl[160]; u8 seq_hdr_buf[256];
    u32 hist[8][256];
    u32 hLL[36], hOF[32], hML[56];
};
struct Z2Shared {
    __align__(16) u8 in[ZE_BLOCK + 48];   // the block, at in[skew ..]; after the literals section: the sequence records
    union { Z2Mf mf; Z2Ent en; };
    u32 x0[1024], x1[1024], x2[1024];     // per-unit exchange arrays (match ends, covers, repcode summaries, FSE states)
    u32 part[40];
    u8 llcode[64], mlcode[128];
    unsigned long long mbar;
    u32 job, tail_from, bad, lit_mode, huf_tbl_bytes, seq_hdr_bytes, all_same;
    u32 stream_bits[4];
    u32 mtmp[36];
};

// ---- shared-memory byte window helpers (buffer is 4-byte aligned, 12 bytes of slack after the last position read)
__device__ __forceinline__ u32 z2_ld32(const u8* b, u32 pos)
{
    const u32* const w = (const u32*)b + (pos >> 2); u32 const sh = (pos & 3) * 8;
    return __funnelshift_r(w[0], w[1], sh);
}
__device__ __forceinline__ u64 z2_ld64(const u8* b, u32 pos)
{
    const u32* const w = (const u32*)b + (pos >> 2); u32 const sh = (pos & 3) * 8;
    u32 const w0 = w[0], w1 = w[1], w2 = w[2];
    return (u64)__funnelshift_r(w0, w1, sh) | ((u64)__funnelshift_r(w1, w2, sh) << 32);
}
__device__ __forceinline__ u32 z2_hash5(u64 v) { return (u32)(((v << 24) * 889523592379ull) >> (64 - Z2_FARLOG)); }

// common prefix of the bytes at a and at c (c < a), looking at most 16 bytes ahead and never past n; A = the 8 bytes at a
__device__ __forceinline__ u32 z2_match16(const u8* b, u32 skew, u64 A, u32 a, u32 c, u32 n)
{
    u64 const x = A ^ z2_ld64(b, skew + c);
    if (x) return ze_common8(0, x);
    if (a + 16 > n) return 8;
    return 8 + ze_common8(z2_ld64(b, skew + a + 8), z2_ld64(b, skew + c + 8));
}

// exclusive scan of one u32 per thread over the CTA (1024 threads); part: 33+ words of shared memory
__device__ __forceinline__ u32 z2_scan(u32 v, u32* part, u32& total)
{
    u32 const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u32 x = v;
    #pragma unroll
    for (int d = 1; d < 32; d <<= 1) { u32 const y = __shfl_up_sync(0xFFFFFFFFu, x, d); if (lane >= (u32)d) x += y; }
    if (lane == 31) part[warp] = x;
    __syncthreads();
    if (warp == 0) {
        u32 y = part[lane];
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) { u32 const z = __shfl_up_sync(0xFFFFFFFFu, y, d); if (lane >= (u32)d) y += z; }
        part[lane] = y;
    }
    __syncthreads();
    total = part[31];
    u32 const r = (warp ? part[warp - 1] : 0u) + x - v;
    __syncthreads();
    return r;
}

// repcode history as an MTF(3) list: apply the (<= 3, most recent first, 0 = none) distinct offsets of a unit
__device__ __forceinline__ void z2_rep_apply(u32& r0, u32& r1, u32& r2, u32 a, u32 b, u32 c)
{
    u32 o[6]; u32 k = 0;
    if (a) o[k++] = a;
    if (b) o[k++] = b;
    if (c) o[k++] = c;
    u32 const t = k;
    u32 const old[3] = {r0, r1, r2};
    #pragma unroll
    for (int i = 0; i < 3; i++) {
        u32 const v = old[i]; bool dup = false;
        for (u32 q = 0; q < t; q++) if (o[q] == v && v) dup = true;
        if (!dup) o[k++] = v;
    }
    r0 = o[0]; r1 = k > 1 ? o[1] : 0; r2 = k > 2 ? o[2] : 0;
}
// the unit's most recent distinct offsets, collected from its LAST sequence backwards
__device__ __forceinline__ void z2_recent_push_back(u32& a, u32& b, u32& c, u32 d)
{
    if (d == a || d == b || d == c) return;
    if (!a) a = d; else if (!b) b = d; else if (!c) c = d;
}

// LSB-first bit accumulator over a zeroed word array (shared or global): the first and the last word of a lane's span
// may be shared with its neighbours (atomicOr), the words in between are its own (plain stores)
struct Z2Bits {
    u32* words; u32 w; u64 acc; u32 nacc; bool first;
    __device__ __forceinline__ void init(u32* base, u32 bitpos) { words = base; w = bitpos >> 5; acc = 0; nacc = bitpos & 31; first = true; }
    __device__ __forceinline__ void put(u32 v, u32 nb) {          // nb <= 26
        acc |= (u64)(v & ((1u << nb) - 1)) << nacc; nacc += nb;
        if (nacc >= 32) { if (first) { atomicOr(&words[w], (u32)acc); first = false; } else words[w] = (u32)acc; w++; acc >>= 32; nacc -= 32; }
    }
    __device__ __forceinline__ void flush() { if (nacc) atomicOr(&words[w], (u32)acc); }
};
__device__ __forceinline__ void z2_or_byte(u32* words, u32 byte_off, u32 v) { atomicOr(&words[byte_off >> 2], (v & 255u) << ((byte_off & 3) * 8)); }

__device__ __forceinline__ u32 z2_fse_init(const ZeCTable& ct, u32 s)      // FSE_initCState2, zstd/zstd.c:2770
{
    u32 const nbo = (u32)(ct.dnb[s] + (1 << 15)) >> 16; u32 const v = (nbo << 16) - (u32)ct.dnb[s];
    return ct.state[(v >> nbo) + ct.dfs[s]];
}

// Huffman code lengths, warp-cooperative front end: rank the present symbols by (count, symbol) -- every lane ranks
// eight of them against all 256 -- then lane 0 runs the serial tree construction of ze_huf_from_sorted
__device__ static bool z2_huf_build(ZeHuf& H, const u32* count, u32* wk, u32 lane)
{
    u32* const sym = wk; u32* const key = wk + 1100;             // key[]: 256 words inside the (not yet used) tail of wk
    u32 present = 0, mxs = 0;
    for (u32 k = 0; k < 8; k++) { u32 const s = lane * 8 + k, c = count[s]; key[s] = c ? ((c << 8) | s) : 0xFFFFFFFFu; if (c) { present++; mxs = s; } }
    #pragma unroll
    for (int d = 16; d > 0; d >>= 1) { present += __shfl_xor_sync(0xFFFFFFFFu, present, d); mxs = max(mxs, __shfl_xor_sync(0xFFFFFFFFu, mxs, d)); }
    __syncwarp();
    if (present < 2) return false;
    u32 mykey[8], rank[8];
    #pragma unroll
    for (int k = 0; k < 8; k++) { mykey[k] = key[lane * 8 + k]; rank[k] = 0; }
    for (u32 t = 0; t < 256; t++) {
        u32 const kt = key[t];
        #pragma unroll
        for (int k = 0; k < 8; k++) rank[k] += kt < mykey[k] ? 1u : 0u;
    }
    #pragma unroll
    for (int k = 0; k < 8; k++) if (mykey[k] != 0xFFFFFFFFu) sym[rank[k]] = lane * 8 + k;
    __syncwarp();
    u32 ok = 0;
    if (lane == 0) ok = ze_huf_from_sorted(H, count, wk, present, mxs) ? 1u : 0u;
    return __shfl_sync(0xFFFFFFFFu, ok, 0) != 0;
}

#ifdef ZB_PHASE_TIMERS
__device__ unsigned long long g_z2_phase[16];
#define Z2_MARK(k) do { if (tid == 0) { long long const t_ = clock64(); atomicAdd(&g_z2_phase[k], (unsigned long long)(t_ - t_phase)); t_phase = t_; } } while (0)
#else
#define Z2_MARK(k) do { } while (0)
#endif

// records of one unit, four at a time (two 16-byte loads in flight) so that the L2 latency is paid once per four
#define Z2_FOR_RECS(k0_, k1_, r_, q_, ...) \
    for (u32 b_ = (k0_) & ~1u; b_ < (k1_); b_ += 4) { \
        ulonglong2 const v0_ = *(const ulonglong2*)(myrec + b_); \
        ulonglong2 const v1_ = b_ + 2 < (k1_) ? *(const ulonglong2*)(myrec + b_ + 2) : make_ulonglong2(0, 0); \
        u64 const rr_[4] = {v0_.x, v0_.y, v1_.x, v1_.y}; \
        _Pragma("unroll") for (u32 i_ = 0; i_ < 4; i_++) { u32 const q_ = b_ + i_; if (q_ >= (k0_) && q_ < (k1_)) { u64 const r_ = rr_[i_]; __VA_ARGS__ } } }

__global__ void __launch_bounds__(Z2_NT, 1)
zb_compress_smem(const u8* __restrict__ src, const ZeBlockJob* __restrict__ jobs, u32 n_jobs, Z2Scratch* __restrict__ scratch,
                 u8* __restrict__ slots, u64 slot_bytes, ZeBlockOut* __restrict__ outs, u32* __restrict__ work_counter, ZeUpload up)
{
    extern __shared__ __align__(16) u8 z2_smem_raw[];
    Z2Shared& S = *(Z2Shared*)z2_smem_raw;
    Z2Scratch& G = scratch[blockIdx.x];
    u32 const tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#ifdef ZB_PHASE_TIMERS
    long long t_phase = clock64();
#endif

    // symbol code look-up tables (ZSTD_LLcode / ZSTD_MLcode, zstd/zstd.c:19738,:19755), derived from the baselines
    if (tid < 64) { u32 c = tid < 16 ? tid : 16; if (tid >= 16) while (c < 35 && tid >= e_LL_base[c + 1]) c++; S.llcode[tid] = (u8)c; }
    if (tid < 128) { u32 c = tid < 32 ? tid : 32; if (tid >= 32) while (c < 52 && tid + 3 >= e_ML_base[c + 1]) c++; S.mlcode[tid] = (u8)c; }
#ifdef __CUDA_ARCH__
    u32 const mbar = (u32)__cvta_generic_to_shared(&S.mbar);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(mbar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    u32 parity = 0;
#endif
    __syncthreads();

    for (;;) {
        // ---------------- next block; its bytes go global -> shared memory in one bulk copy
        if (tid == 0) {
            u32 const jn = atomicAdd(work_counter, 1u);
            if (up.progress && jn < n_jobs) {       // host input still being uploaded: wait until this block (and a margin) has landed
                unsigned long long const want = jobs[jn].src_pos + jobs[jn].size + 256u;
                unsigned long long const need = want < up.total ? want : up.total;
                long long t0 = clock64(); unsigned long long seen = 0;
                for (;;) {
                    unsigned long long const now = *(volatile const unsigned long long*)up.progress;
                    if (now >= need) break;
                    if (now != seen) { seen = now; t0 = clock64(); }                  // the upload is making progress
                    __nanosleep(400);
                    if (clock64() - t0 > 6000000000ll) { atomicExch(up.status, 1u); break; }
                }
                __threadfence();
            }
            S.job = jn;
#ifdef __CUDA_ARCH__
            if (jn < n_jobs && jobs[jn].size) {
                const u8* const p = src + jobs[jn].src_pos;
                const u8* const pal = (const u8*)((uintptr_t)p & ~(uintptr_t)15);
                u32 const bytes = ((u32)(p - pal) + jobs[jn].size + 15u) & ~15u;
                u32 const dst = (u32)__cvta_generic_to_shared(S.in);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             :: "r"(dst), "l"(pal), "r"(bytes), "r"(mbar) : "memory");
            }
#endif
        }
        __syncthreads();
        u32 const j = S.job;
        if (j >= n_jobs) return;
        ZeBlockJob const job = jobs[j];
        u32 const n = job.size;
        const u8* const gsrc = src + job.src_pos;
        u32 const skew = (u32)((uintptr_t)gsrc & 15);
        u8* const out = slots + (u64)j * slot_bytes;
        u32* const ow = (u32*)out;                                    // slots are 16-byte aligned
        const u8* const in = S.in;                                    // block byte i is in[skew + i]
        // while the copy is in flight: clear the far table
        for (u32 i = tid; i < (1u << Z2_FARLOG) / 2; i += Z2_NT) ((u32*)S.mf.far)[i] = 0xFFFFFFFFu;
        if (tid == 0) { S.bad = 0; S.all_same = 1; }
#ifdef __CUDA_ARCH__
        if (n) { u32 ok = 0;
          while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                                   : "=r"(ok) : "r"(mbar), "r"(parity) : "memory");
          parity ^= 1; }
#else
        for (u32 i = tid; i < n; i += Z2_NT) S.in[skew + i] = gsrc[i];
#endif
        __syncthreads();
        if (tid < 32) S.in[skew + n + tid] = 0;                       // slack after the block: defined bytes for window loads
        __syncthreads();

        // ---------------- trivial blocks: tiny -> raw, all bytes equal -> RLE (not as the first block of a frame: the
        // reference emits its first block compressed, zstd/zstd.c:27378)
        if (n >= 9) {
            u8 const b0 = in[skew]; bool same = true;
            for (u32 i = tid; i < n && same; i += Z2_NT) if (in[skew + i] != b0) same = false;
            if (!same) S.all_same = 0;
            __syncthreads();
            if (S.all_same && !job.first) {
                if (tid == 0) { u32 const bh = job.last | (1u << 1) | (n << 3); out[0] = (u8)bh; out[1] = (u8)(bh >> 8); out[2] = (u8)(bh >> 16); out[3] = b0; outs[j].csize = 4; }
                __syncthreads();
                continue;
            }
        } else {
            for (u32 i = tid; i < n; i += Z2_NT) out[3 + i] = in[skew + i];
            if (tid == 0) { u32 const bh = job.last | (n << 3); out[0] = (u8)bh; out[1] = (u8)(bh >> 8); out[2] = (u8)(bh >> 16); outs[j].csize = 3 + n; }
            __syncthreads();
            continue;
        }
        Z2_MARK(0);

        // ================================================================= match finding, a pass of Z2_Q positions at a time
        u32 const npass = (n + Z2_Q - 1) / Z2_Q;
        for (u32 ps = 0; ps < npass; ps++) {
            u32 const q0 = ps * Z2_Q, npos = min((u32)Z2_Q, n - q0);
            // ---- near links: warp w owns sub-chunk w
            if (warp < ((npos + 1023) >> 10)) {
                u16* const tab = S.mf.near[warp];
                for (u32 i = lane; i < 256; i += 32) ((u32*)tab)[i] = 0xFFFFFFFFu;
                __syncwarp();
                u32 const c0 = q0 + (warp << 10), c1 = min(c0 + 1024u, q0 + npos);
                for (u32 st = 0; st < 32; st++) {
                    u32 const p = c0 + st * 32 + lane, pl = st * 32 + lane;
                    bool const valid = p < c1 && p + 8 <= n;
                    u32 const v = valid ? z2_ld32(in, skew + p) : 0u;
                    u32 const h = (v * 2654435761u) >> 17;                       // 9 slot bits + 6 tag bits
                    u32 const slot = h >> 6, tag = h & 63u, entry = (pl << 6) | tag;
                    u32 dn = 0;
                    if (lane < 16 && valid) { u32 const e = tab[slot]; if ((e & 63u) == tag && (e >> 6) < pl) dn = pl - (e >> 6); tab[slot] = (u16)entry; }
                    __syncwarp();
                    if (lane >= 16 && valid) { u32 const e = tab[slot]; if ((e & 63u) == tag && (e >> 6) < pl) dn = pl - (e >> 6); tab[slot] = (u16)entry; }
                    __syncwarp();
                    if (p < c1) S.mf.dist[p - q0] = (u16)dn;
                }
            }
            __syncthreads();
            // ---- far links + verification, rounds of 1024 positions
            u32 const rounds = (npos + 1023) >> 10;
            for (u32 r = 0; r < rounds; r++) {
                u32 const p = q0 + (r << 10) + tid;
                bool const inr = p < q0 + npos, valid = inr && p + 8 <= n;
                u32 best = 0, bd = 0, hf = 0;
                if (valid) {
                    u64 const A = z2_ld64(in, skew + p);
                    hf = z2_hash5(A);
                    u32 const cf = S.mf.far[hf];
                    u32 const df = (p - cf) & 0xFFFFu;
                    if (cf != 0xFFFFu && df != 0 && df <= p) {
                        u32 const m = z2_match16(in, skew, A, p, p - df, n);
                        if (m >= 4) { best = m; bd = df; }
                    }
                    u32 const dn = S.mf.dist[p - q0];
                    if (dn && dn != bd) {
                        u32 const m = z2_match16(in, skew, A, p, p - dn, n);
                        if (m >= 4 && m >= best) { best = m; bd = dn; }
                    }
                    if (p >= 4) {      // distances 1..4: the four bytes in front of p are next to A's
                        u32 const W = z2_ld32(in, skew + p - 4);
                        u64 const WA = (u64)W | ((u64)(u32)A << 32);
                        #pragma unroll
                        for (u32 d = 1; d <= 4; d++) {
                            if ((u32)(WA >> (8 * (4 - d))) == (u32)A && d != bd) {
                                u32 const m = z2_match16(in, skew, A, p, p - d, n);
                                if (m > best) { best = m; bd = d; }
                            }
                        }
                    }
                    if (best > 15) best = 15;
                }
                u32 mn = __shfl_down_sync(0xFFFFFFFFu, best, 1);
                if (lane == 0) S.mtmp[warp] = best;
                __syncthreads();
                if (lane == 31) mn = warp < 31 ? S.mtmp[warp + 1] : 0u;
                bool const take = best >= 4 && !(best < 15 && mn > best + 1);        // one-step lazy
                if (inr) S.mf.dist[p - q0] = (u16)(take ? bd : 0u);
                if (valid && (p & 0xFFFFu) != 0xFFFFu) S.mf.far[hf] = (u16)p;
                __syncthreads();
            }
            // ---- parse: thread t (< 125) walks unit ps * 125 + t
            if (tid < Z2_UPP) {
                u32 const u0 = q0 + tid * Z2_U;
                if (u0 < q0 + npos) {
                    u32 const u1 = min(u0 + (u32)Z2_U, q0 + npos);
                    u64* const rec = G.rec + (u64)(ps * Z2_UPP + tid) * Z2_RMAX;
                    u32 ip = u0 ? u0 : 1u, anchor = u0, cnt = 0, endm = 0;
                    while (ip < u1) {
                        u32 const d = S.mf.dist[ip - q0];
                        if (!d) { ip++; continue; }
                        u32 start = ip, a = ip, c = ip - d;
                        for (;;) {          // forward length, 8 bytes at a time, never past n
                            if (a + 8 > n) { while (a < n && in[skew + a] == in[skew + c]) { a++; c++; } break; }
                            u64 const x = z2_ld64(in, skew + a) ^ z2_ld64(in, skew + c);
                            if (x) { a += ze_common8(0, x); break; }
                            a += 8; c += 8;
                        }
                        while (start > anchor && start > d && in[skew + start - 1] == in[skew + start - 1 - d]) start--;
                        u32 const len = a - start;
                        if (len < 4 || cnt >= Z2_RMAX - 1) { ip++; continue; }            // (verified >= 4 bytes: does not happen)
                        rec[cnt++] = (u64)start | ((u64)len << 17) | ((u64)d << 35);
                        ip = a; anchor = a; endm = a;
                    }
                    S.x0[ps * Z2_UPP + tid] = endm; S.x1[ps * Z2_UPP + tid] = cnt;        // for the unit's own thread
                }
            }
            __syncthreads();
        }
        Z2_MARK(1);
        u32 const nunits = (npass - 1) * Z2_UPP + ((n - (npass - 1) * Z2_Q) + Z2_U - 1) / Z2_U;
        u32 const my_cnt = tid < nunits ? S.x1[tid] : 0u;
        __syncthreads();

        // ================================================================= stitch: cover on entry of every unit
        if (tid == 0) {
            u32 c = 0;
            for (u32 u = 0; u < nunits; u++) { S.x1[u] = c; u32 const e = S.x0[u]; if (e >= c + 3) c = e; }
            S.tail_from = c;
        }
        __syncthreads();
        // survivors of this unit: drop what an earlier match covers, front-trim the first one that sticks out by >= 3
        u32 const cin = tid < nunits ? S.x1[tid] : 0u;
        u32 first_k = my_cnt, lits = 0, ra = 0, rb = 0, rc = 0;
        const u64* const myrec = G.rec + (u64)tid * Z2_RMAX;
        if (my_cnt) {
            u32 cover = cin;
            Z2_FOR_RECS(0u, my_cnt, r, q, {
                u32 s = (u32)r & 0x1FFFFu; u32 const e = s + ((u32)(r >> 17) & 0x3FFFFu);
                if (first_k == my_cnt && (s >= cover || e >= cover + 3)) first_k = q;
                if (q >= first_k) {
                    if (s < cover) s = cover;
                    lits += s - cover; cover = e;
                    u32 const d = (u32)(r >> 35);          // the unit's three most recent distinct offsets (move to front)
                    if (d == ra) { } else if (d == rb) { rb = ra; ra = d; } else { rc = rb; rb = ra; ra = d; }
                }
            })
        }
        u32 const nsv = my_cnt - first_k;
        __syncthreads();
        S.x0[tid] = ra; S.x1[tid] = rb; S.x2[tid] = rc;
        u32 nseq, nlit_seq;
        u32 const seq_base = z2_scan(nsv, S.part, nseq);
        u32 const lit_base = z2_scan(lits, S.part, nlit_seq);
        // repcode history on entry of every unit: warp 0, lane l folds units [32 l, 32 l + 32)
        if (warp == 0) {
            u32 r0 = 0, r1 = 0, r2 = 0;
            for (u32 u = lane * 32; u < lane * 32 + 32 && u < nunits; u++) z2_rep_apply(r0, r1, r2, S.x0[u], S.x1[u], S.x2[u]);
            u32 e0 = 0, e1 = 0, e2 = 0;
            u32 s0 = job.first ? 1u : 0u, s1 = job.first ? 4u : 0u, s2 = job.first ? 8u : 0u;      // a frame starts from {1, 4, 8}; later blocks: unknown
            for (u32 l = 0; l < 32; l++) {
                u32 const a = __shfl_sync(0xFFFFFFFFu, r0, l), b = __shfl_sync(0xFFFFFFFFu, r1, l), c = __shfl_sync(0xFFFFFFFFu, r2, l);
                if (lane == l) { e0 = s0; e1 = s1; e2 = s2; }
                z2_rep_apply(s0, s1, s2, a, b, c);
            }
            for (u32 u = lane * 32; u < lane * 32 + 32 && u < nunits; u++) {
                u32 const a = S.x0[u], b = S.x1[u], c = S.x2[u];
                S.x0[u] = e0; S.x1[u] = e1; S.x2[u] = e2;
                z2_rep_apply(e0, e1, e2, a, b, c);
            }
        }
        for (u32 i = tid; i < 8 * 256; i += Z2_NT) ((u32*)S.en.hist)[i] = 0;
        if (tid < 36) S.en.hLL[tid] = 0;
        if (tid < 32) S.en.hOF[tid] = 0;
        if (tid < 56) S.en.hML[tid] = 0;
        __syncthreads();
        u32 const tail_from = S.tail_from;
        u32 const nlit = nlit_seq + (n - tail_from);
        u8* const lit = nlit <= Z2_STAGE ? (u8*)S.en.stage : G.lit;
        // final sequences (offsets coded against the true history), code histograms, literal gather + histogram
        if (nsv) {
            u32 r0 = S.x0[tid], r1 = S.x1[tid], r2 = S.x2[tid];
            u32 cover = cin, lp = lit_base; u32* const hist = S.en.hist[warp & 7];
            Z2_FOR_RECS(first_k, my_cnt, r, q, {
                u32 s = (u32)r & 0x1FFFFu; u32 const e = s + ((u32)(r >> 17) & 0x3FFFFu), d = (u32)(r >> 35);
                if (s < cover) s = cover;
                u32 const ll = s - cover, ml = e - s;
                if (ll == 0 && d == r0) S.bad = 1;         // two adjacent matches with one offset: cannot be formed (see the notes); raw block if it ever is
                u32 const ob = ze_off_code(d, ll, r0, r1, r2);
                G.fseq[seq_base + (q - first_k)] = (u64)ll | ((u64)(ml - 3) << 17) | ((u64)ob << 34);
                u32 const lc = ll < 64 ? S.llcode[ll] : ze_hibit(ll) + 19, mc = ml - 3 < 128 ? S.mlcode[ml - 3] : ze_hibit(ml - 3) + 36, oc = ze_hibit(ob);
                atomicAdd(&S.en.hLL[lc], 1u); atomicAdd(&S.en.hML[mc], 1u); atomicAdd(&S.en.hOF[oc], 1u);
                for (u32 k = 0; k < ll; k++) { u8 const b = in[skew + cover + k]; lit[lp + k] = b; atomicAdd(&hist[b], 1u); }
                lp += ll; cover = e;
            })
        }
        {   // last literals of the block
            u32* const hist = S.en.hist[warp & 7];
            for (u32 k = tid; k < n - tail_from; k += Z2_NT) { u8 const b = in[skew + tail_from + k]; lit[nlit_seq + k] = b; atomicAdd(&hist[b], 1u); }
        }
        __syncthreads();
        for (u32 s = tid; s < 256; s += Z2_NT) { u32 t = 0; for (int k = 0; k < 8; k++) t += S.en.hist[k][s]; S.en.hist[0][s] = t; }
        __syncthreads();
        Z2_MARK(2);

        // ================================================================= entropy tables: four warps
        if (nseq) {
            if (tid == 0)  ze_make_table(S.en.ct[0], S.en.hLL, 35, nseq, 9, 6, e_LL_defnorm, 35, S.en.tmp_sym[0]);
            if (tid == 32) ze_make_table(S.en.ct[1], S.en.hOF, 31, nseq, 8, 5, e_OF_defnorm, 28, S.en.tmp_sym[1]);
            if (tid == 64) ze_make_table(S.en.ct[2], S.en.hML, 52, nseq, 9, 6, e_ML_defnorm, 52, S.en.tmp_sym[2]);
        }
        if (warp == 3) {
            u32 mode = 0, tb = 0;
            u32 most = 0; for (u32 s = lane; s < 256; s += 32) most = max(most, S.en.hist[0][s]);
            #pragma unroll
            for (int d = 16; d > 0; d >>= 1) most = max(most, __shfl_xor_sync(0xFFFFFFFFu, most, d));
            if (nlit >= 8 && most == nlit) mode = 1;                                   // RLE literals
            else if (nlit >= 64) {                                                     // ZSTD_minLiteralsToCompress, zstd/zstd.c:20918
                bool const ok = z2_huf_build(S.en.huf, S.en.hist[0], S.en.wk, lane);
                if (lane == 0 && ok) { tb = ze_huf_write_table(S.en.huf_tbl, S.en.huf, S.en.ct[3], S.en.tmp_sym[3]); if (tb) mode = 2; }
            }
            if (lane == 0) { S.lit_mode = mode; S.huf_tbl_bytes = tb; }
        }
        __syncthreads();
        Z2_MARK(3);

        // ================================================================= literals section, written straight into the slot
        bool const four = nlit >= 256;
        u32 const seg = four ? (nlit + 3) / 4 : nlit, nstreams = four ? 4u : 1u;
        u32 const nchunks = (nlit + 31) / 32, lrounds = (nchunks + Z2_NT - 1) / Z2_NT;
        u32 lit_mode = S.lit_mode, lit_payload = 0, lh = 0;
        u32 const tb = S.huf_tbl_bytes;
        // one chunk = 32 literals at an aligned index (two 16-byte loads); a chunk meets at most one stream border
        #define Z2_CHUNK_BITS(g_, sA_, bnd_, bA_, bB_, c_) \
            uint4 c_[2]; u32 sA_ = 0, bnd_ = 0xFFFFFFFFu, bA_ = 0, bB_ = 0; \
            if ((g_) < nchunks) { \
                c_[0] = *(const uint4*)(lit + 32 * (g_)); c_[1] = *(const uint4*)(lit + 32 * (g_) + 16); \
                if (four) { sA_ = min(32 * (g_) / seg, 3u); if (sA_ < 3 && (sA_ + 1) * seg < 32 * (g_) + 32) bnd_ = (sA_ + 1) * seg; } \
                const u32* const cw_ = (const u32*)c_; \
                _Pragma("unroll") for (u32 k_ = 0; k_ < 32; k_++) { u32 const i_ = 32 * (g_) + k_; if (i_ < nlit) { u32 const nb_ = S.en.huf.nb[(cw_[k_ >> 2] >> ((k_ & 3) * 8)) & 255u]; if (i_ < bnd_) bA_ += nb_; else bB_ += nb_; } } \
            } else { c_[0] = make_uint4(0, 0, 0, 0); c_[1] = c_[0]; }
        if (lit_mode == 2) {
            u32 run[4] = {0, 0, 0, 0};
            for (u32 rd = 0; rd < lrounds; rd++) {
                u32 const g = rd * Z2_NT + tid;
                Z2_CHUNK_BITS(g, sA, bnd, bA, bB, cdat)
                (void)cdat;
                for (u32 st = 0; st < nstreams; st++) {
                    u32 const v = (sA == st ? bA : 0u) + (sA + 1 == st ? bB : 0u);
                    u32 tot; (void)z2_scan(v, S.part, tot); run[st] += tot;
                }
            }
            if (tid < 4) S.stream_bits[tid] = run[tid];
            u32 est = tb + (four ? 6u : 0u);
            for (u32 st = 0; st < nstreams; st++) est += (run[st] + 1 + 7) / 8;
            if (est + (nlit >> 6) + 2 >= nlit) lit_mode = 0;                      // ZSTD_minGain, zstd/zstd.c:19831 (uniform: every thread holds the totals)
            else lit_payload = est;
            __syncthreads();
        }
        if (lit_mode == 2) {
            lh = 3 + (nlit >= 1024) + (nlit >= 16384);
            u32 sbyte[4], sbytes[4]; { u32 o = tb + (four ? 6u : 0u); for (u32 st = 0; st < nstreams; st++) { sbytes[st] = (S.stream_bits[st] + 1 + 7) / 8; sbyte[st] = o; o += sbytes[st]; } }
            u32 const o_pl = 3 + lh;                                                  // byte offset of the payload in the slot
            for (u32 i = tid; i < (o_pl + lit_payload) / 4 + 2; i += Z2_NT) ow[i] = 0;
            __syncthreads();
            for (u32 i = tid; i < tb; i += Z2_NT) z2_or_byte(ow, o_pl + i, S.en.huf_tbl[i]);
            if (four && tid < 3) { z2_or_byte(ow, o_pl + tb + 2 * tid, sbytes[tid]); z2_or_byte(ow, o_pl + tb + 2 * tid + 1, sbytes[tid] >> 8); }
            if (tid < nstreams) { Z2Bits w; w.init(ow, (o_pl + sbyte[tid]) * 8 + S.stream_bits[tid]); w.put(1, 1); w.flush(); }     // end marks
            u32 run[4] = {0, 0, 0, 0};
            for (u32 rd = 0; rd < lrounds; rd++) {
                u32 const g = rd * Z2_NT + tid;
                Z2_CHUNK_BITS(g, sA, bnd, bA, bB, cdat)
                u32 pA = 0, pB = 0;
                for (u32 st = 0; st < nstreams; st++) {
                    u32 const v = (sA == st ? bA : 0u) + (sA + 1 == st ? bB : 0u);
                    u32 tot; u32 const pre = z2_scan(v, S.part, tot) + run[st]; run[st] += tot;
                    if (sA == st) pA = pre;
                    if (sA + 1 == st) pB = pre;
                }
                if (g < nchunks) {      // symbols are written last to first: a span ends at (stream total - bits before it)
                    const u32* const cw = (const u32*)cdat;
                    if (bB) {
                        Z2Bits w; w.init(ow, (o_pl + sbyte[sA + 1]) * 8 + (S.stream_bits[sA + 1] - pB - bB));
                        #pragma unroll
                        for (u32 kk = 0; kk < 32; kk++) { u32 const k = 31 - kk, i = 32 * g + k; if (i < nlit && i >= bnd) { u32 const sym = (cw[k >> 2] >> ((k & 3) * 8)) & 255u; w.put(S.en.huf.code[sym], S.en.huf.nb[sym]); } }
                        w.flush();
                    }
                    if (bA) {
                        Z2Bits w; w.init(ow, (o_pl + sbyte[sA]) * 8 + (S.stream_bits[sA] - pA - bA));
                        #pragma unroll
                        for (u32 kk = 0; kk < 32; kk++) { u32 const k = 31 - kk, i = 32 * g + k; if (i < nlit && i < bnd) { u32 const sym = (cw[k >> 2] >> ((k & 3) * 8)) & 255u; w.put(S.en.huf.code[sym], S.en.huf.nb[sym]); } }
                        w.flush();
                    }
                }
            }
            if (tid == 0) {      // section header (ZSTD_compressLiterals, zstd/zstd.c:20932-21038), OR-ed: its word also holds payload bits
                u32 v; u32 hb[5];
                if (lh == 3) { v = 2u | ((four ? 1u : 0u) << 2) | (nlit << 4) | (lit_payload << 14); hb[0] = v; hb[1] = v >> 8; hb[2] = v >> 16; }
                else if (lh == 4) { v = 2u | (2u << 2) | (nlit << 4) | (lit_payload << 18); hb[0] = v; hb[1] = v >> 8; hb[2] = v >> 16; hb[3] = v >> 24; }
                else { v = 2u | (3u << 2) | (nlit << 4) | (lit_payload << 22); hb[0] = v; hb[1] = v >> 8; hb[2] = v >> 16; hb[3] = v >> 24; hb[4] = lit_payload >> 10; }
                for (u32 k = 0; k < lh; k++) z2_or_byte(ow, 3 + k, hb[k]);
            }
        } else {
            lh = 1 + (nlit > 31) + (nlit > 4095);
            u8* const o = out + 3;
            if (lit_mode == 0) { for (u32 i = tid; i < nlit; i += Z2_NT) o[lh + i] = lit[i]; lit_payload = nlit; }
            else { if (tid == 0) o[lh] = lit[0]; lit_payload = 1; }
            if (tid == 0) {      // ZSTD_noCompressLiterals / ZSTD_compressRleLiteralsBlock, zstd/zstd.c:20851-20930
                u32 const t = lit_mode;
                if (lh == 1) o[0] = (u8)(t | (nlit << 3));
                else if (lh == 2) { u32 const v = t | (1u << 2) | (nlit << 4); o[0] = (u8)v; o[1] = (u8)(v >> 8); }
                else { u32 const v = t | (3u << 2) | (nlit << 4); o[0] = (u8)v; o[1] = (u8)(v >> 8); o[2] = (u8)(v >> 16); }
            }
        }
        __syncthreads();
        Z2_MARK(4);

        // ================================================================= sequences section
        u32 seq_payload = 0;
        if (tid == 0) {      // header (ZSTD_entropyCompressSeqStore_internal, zstd/zstd.c:25893-25926)
            u8* const q = S.en.seq_hdr_buf; u32 k = 0;
            if (nseq < 128) q[k++] = (u8)nseq;
            else if (nseq < 0x7F00) { q[k++] = (u8)((nseq >> 8) + 0x80); q[k++] = (u8)nseq; }
            else { q[k++] = 0xFF; q[k++] = (u8)(nseq - 0x7F00); q[k++] = (u8)((nseq - 0x7F00) >> 8); }
            if (nseq) {
                q[k++] = (u8)((S.en.ct[0].mode << 6) | (S.en.ct[1].mode << 4) | (S.en.ct[2].mode << 2));
                for (int t = 0; t < 3; t++) for (u32 i = 0; i < S.en.ct[t].hdr_bytes; i++) q[k++] = S.en.ct[t].hdr[i];
            }
            S.seq_hdr_bytes = k;
        }
        // the input is no longer needed (a raw block is copied from global memory): its buffer takes the sequence records
        u64* const sq = (u64*)S.in;
        for (u32 i = tid; i < nseq && i < Z2_SEQ_SMEM; i += Z2_NT) sq[i] = G.fseq[i];
        __syncthreads();
        u32 const shb = S.seq_hdr_bytes;
        u32 const o_sh = 3 + lh + lit_payload;
        if (nseq) {
            ZeCTable const& cL = S.en.ct[0]; ZeCTable const& cO = S.en.ct[1]; ZeCTable const& cM = S.en.ct[2];
            bool const rL = cL.mode == 1, rO = cO.mode == 1, rM = cM.mode == 1;
            u32 const K = (nseq + Z2_NT - 1) / Z2_NT;
            u32 const lo = tid * K, hi = min(lo + K, nseq);
            bool const act = lo < nseq;
            #define Z2_SEQ(i_) ((i_) < Z2_SEQ_SMEM ? sq[i_] : G.fseq[i_])
            #define Z2_CODES(r_, lc_, oc_, mc_, ll_, mb_, ob_) \
                u32 const ll_ = (u32)(r_) & 0x1FFFFu, mb_ = (u32)((r_) >> 17) & 0x1FFFFu, ob_ = (u32)((r_) >> 34); \
                u32 const lc_ = ll_ < 64 ? S.llcode[ll_] : ze_hibit(ll_) + 19, mc_ = mb_ < 128 ? S.mlcode[mb_] : ze_hibit(mb_) + 36, oc_ = ze_hibit(ob_);
            // states on entry of this lane's range (= after encoding sequence `hi`): exact for the last range, a guess otherwise
            u32 inL = 0, inO = 0, inM = 0;
            if (act) {
                u32 const ws = hi == nseq ? nseq - 1 : min(nseq - 1, hi + Z2_WARM - 1);
                { u64 const r = Z2_SEQ(ws); Z2_CODES(r, lc, oc, mc, ll, mb, ob) (void)ll; (void)mb; (void)ob;
                  inL = rL ? 0u : z2_fse_init(cL, lc); inO = rO ? 0u : z2_fse_init(cO, oc); inM = rM ? 0u : z2_fse_init(cM, mc); }
                if (hi != nseq) for (u32 i = ws; i-- > hi;) {
                    u64 const r = Z2_SEQ(i); Z2_CODES(r, lc, oc, mc, ll, mb, ob) (void)ll; (void)mb; (void)ob;
                    if (!rO) { u32 const nb = (inO + (u32)cO.dnb[oc]) >> 16; inO = cO.state[(inO >> nb) + cO.dfs[oc]]; }
                    if (!rM) { u32 const nb = (inM + (u32)cM.dnb[mc]) >> 16; inM = cM.state[(inM >> nb) + cM.dfs[mc]]; }
                    if (!rL) { u32 const nb = (inL + (u32)cL.dnb[lc]) >> 16; inL = cL.state[(inL >> nb) + cL.dfs[lc]]; }
                }
            }
            u32 bits = 0;
            for (;;) {
                // run the range from its entry states: bit count and exit states
                u32 sL = inL, sO = inO, sM = inM; bits = 0;
                if (act) for (u32 i = hi; i-- > lo;) {
                    u64 const r = Z2_SEQ(i); Z2_CODES(r, lc, oc, mc, ll, mb, ob) (void)ll; (void)mb; (void)ob;
                    bits += e_LL_bits[lc] + e_ML_bits[mc] + oc;
                    if (i == nseq - 1) continue;                                         // the last sequence only initialises the states
                    if (!rO) { u32 const nb = (sO + (u32)cO.dnb[oc]) >> 16; bits += nb; sO = cO.state[(sO >> nb) + cO.dfs[oc]]; }
                    if (!rM) { u32 const nb = (sM + (u32)cM.dnb[mc]) >> 16; bits += nb; sM = cM.state[(sM >> nb) + cM.dfs[mc]]; }
                    if (!rL) { u32 const nb = (sL + (u32)cL.dnb[lc]) >> 16; bits += nb; sL = cL.state[(sL >> nb) + cL.dfs[lc]]; }
                }
                S.x0[tid] = sL; S.x1[tid] = sO; S.x2[tid] = sM;
                __syncthreads();
                bool wrong = false;
                if (act && hi != nseq) {       // the lane above me ends where I begin
                    u32 const tL = S.x0[tid + 1], tO = S.x1[tid + 1], tM = S.x2[tid + 1];
                    if (tL != inL || tO != inO || tM != inM) { wrong = true; inL = tL; inO = tO; inM = tM; }
                }
                if (!__syncthreads_or(wrong ? 1 : 0)) break;
            }
            // bit offset of this range: ranges are written from the last sequence down, so everything above me comes first
            u32 total_bits; u32 const before_fwd = z2_scan(bits, S.part, total_bits);
            u32 const my_off = total_bits - before_fwd - bits;
            u32 const logL = rL ? 0u : cL.log, logO = rO ? 0u : cO.log, logM = rM ? 0u : cM.log;
            seq_payload = (total_bits + logM + logO + logL + 1 + 7) / 8;
            u32* const stage = seq_payload + 8 <= Z2_STAGE ? S.en.stage : G.stage;
            for (u32 i = tid; i < seq_payload / 4 + 2; i += Z2_NT) stage[i] = 0;
            __syncthreads();
            if (act) {
                Z2Bits w; w.init(stage, my_off);
                u32 sL = inL, sO = inO, sM = inM;
                for (u32 i = hi; i-- > lo;) {
                    u64 const r = Z2_SEQ(i); Z2_CODES(r, lc, oc, mc, ll, mb, ob)
                    if (i != nseq - 1) {
                        if (!rO) { u32 const nb = (sO + (u32)cO.dnb[oc]) >> 16; w.put(sO, nb); sO = cO.state[(sO >> nb) + cO.dfs[oc]]; }
                        if (!rM) { u32 const nb = (sM + (u32)cM.dnb[mc]) >> 16; w.put(sM, nb); sM = cM.state[(sM >> nb) + cM.dfs[mc]]; }
                        if (!rL) { u32 const nb = (sL + (u32)cL.dnb[lc]) >> 16; w.put(sL, nb); sL = cL.state[(sL >> nb) + cL.dfs[lc]]; }
                    }
                    w.put(ll, e_LL_bits[lc]); w.put(mb, e_ML_bits[mc]); w.put(ob, oc);
                }
                if (lo == 0) { w.put(sM, logM); w.put(sO, logO); w.put(sL, logL); w.put(1, 1); }      // flush ML, OF, LL states + end mark
                w.flush();
            }
            __syncthreads();
            const u8* const ps8 = (const u8*)stage;
            for (u32 i = tid; i < seq_payload; i += Z2_NT) out[o_sh + shb + i] = ps8[i];
        }
        for (u32 i = tid; i < shb; i += Z2_NT) out[o_sh + i] = S.en.seq_hdr_buf[i];
        Z2_MARK(5);
        // ================================================================= block header; raw fallback (cSize >= srcSize - minGain, zstd/zstd.c:25987)
        u32 const body = lh + lit_payload + shb + seq_payload;
        bool const use_raw = S.bad || body + (n >> 7) + 2 >= n || body >= ZE_BLOCK;
        __syncthreads();
        if (use_raw) for (u32 i = tid; i < n; i += Z2_NT) out[3 + i] = gsrc[i];
        if (tid == 0) {
            u32 const bsz = use_raw ? n : body;
            u32 const bh = job.last | ((use_raw ? 0u : 2u) << 1) | (bsz << 3);
            out[0] = (u8)bh; out[1] = (u8)(bh >> 8); out[2] = (u8)(bh >> 16);
            outs[j].csize = 3 + bsz;
        }
        Z2_MARK(6);
        __syncthreads();
    }
}
