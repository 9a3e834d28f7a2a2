// mz_decode_pc.cuh -- MinLZ block decode, parser / copier kernel (sm_100a).
//
// Replaces minLZDecode (reference decode.go:178-622 / decodeBlockAsm).
//
// Why it looks the way it does
// ----------------------------
// A MinLZ token stream has two serial dependencies: where token i+1 starts
// (after token i's header and literals) and what a back-reference reads (bytes
// produced by earlier tokens).  A warp that walks one block token by token
// spends ~1400 cycles per token waiting for dependent loads (v0, 70 ms for
// 4096 x 1 MiB).  This kernel splits the two:
//
//   * PARSER warp: lane l walks the token stream of block slot l of the CTA
//     (up to 32 blocks per CTA, SIMT across blocks).  Walking is the only truly
//     serial part, and one lane per block is the cheapest place to do it: the
//     lane knows the output cursor and the repeat offset for free, validates
//     every token exactly like the reference, and emits fully resolved
//     descriptors {literal source, output position, literal length, match
//     length, offset} into shared memory, 32 tokens per block per round.
//   * COPIER warps: lane k owns token k of a batch: it decodes the token's
//     fields, a warp prefix sum gives its output position, a "last defined"
//     scan resolves repeat offsets, and every token is validated exactly like
//     the reference (decode.go:221,326,410,572).  Then all 32 back-reference
//     gathers of a batch go out together (16-byte cp.async per lane into
//     shared memory), literals come from one coalesced load of the batch's
//     stream span, everything is assembled in a shared-memory image of the
//     batch's output span and leaves with coalesced 16-byte stores.  Tokens
//     whose source lies inside the span being assembled are applied afterwards
//     in stream order, warp-cooperatively (overlapping copies replicate the
//     `offset`-byte pattern).
//   * LEXER warps take everything off the parser's chain that is not the chain itself.  Where
//     the next token starts depends on the previous token, but how long a token WOULD be if
//     one started at byte i depends on bytes i, i+1 only: the lexers prefetch every slot's
//     stream into a shared-memory ring (one cp.async.bulk per slot and round, completed on an
//     mbarrier; -DMZ_DEC_BULK=0: per-lane 16-byte cp.async) and fill adv[i] for every byte
//     offset of it, lanes in parallel.  The parser's step is then
//     i += adv[i] -- one shared-memory load -- plus the descriptor stores (~45 cycles per
//     token instead of ~350).  Extended lengths (adv 0) and bytes not lexed yet take the
//     parser's old path.
//   * parser and copiers run bulk-synchronously: round r+1 is parsed while
//     round r is copied (double-buffered descriptors, one __syncthreads per
//     round), so there is no fine-grained inter-warp signalling.
//
// Result contract = the reference's: status 0 and exactly dst_len bytes, or
// status 1 (decodeErrCodeCorrupt); a corrupt block never writes outside its
// own dst range (tokens are validated before their descriptors are emitted).
#pragma once

#include "mz_common.cuh"

namespace mz {

constexpr int kDecSlots = 28;      // block slots per CTA (one parser lane and one copier warp each)
constexpr int kDecCopiers = 28;    // copier warps per CTA
constexpr int kDecLexers = 3;      // lexer warps per CTA (stream prefetch + token lengths for the parser)
constexpr int kDecWarps = 1 + kDecCopiers + kDecLexers;
constexpr int kDecThreads = kDecWarps * 32;
constexpr int kDecTok = 32;        // tokens per batch
constexpr int kDecShort = 40;      // max literal / match length of a "short" token
constexpr int kDecStage = kDecTok * 2 * kDecShort + 32;       // output image of a batch (+ alignment slack)
constexpr int kDecLitStage = kDecTok * (8 + kDecShort) + 48;  // stream span of a batch
constexpr int kDecScratch = 32 * 48 + 16;                     // per-lane 48-byte gather landing zone
constexpr int kDescStride = kDecTok + 2;                      // uint16 per (slot, buffer); 17 words -> conflict free
constexpr int kDecRing = 1024;                                // per-slot ring of compressed bytes (+ as many token lengths)
constexpr int kDecRingAhead = kDecRing - 32;                  // bytes kept requested ahead of the batch being copied
constexpr uint32_t kDescGlobal = 0x8000;                      // descriptor flag: the token's header is not in the ring

constexpr uint32_t kBatchLong = 0x100;   // the batch is a single long token
constexpr uint32_t kBatchEnded = 0x200;  // the stream ended behind this batch (decode.go:615 check is due)
constexpr uint32_t kBatchBad = 0x400;    // the parser hit a token that does not fit the stream

struct DecSlotState {
    // written by the parser, read by the copiers after the barrier
    uint32_t count[2][kDecSlots];    // tokens in the batch | kBatch* flags
    uint32_t s_first[2][kDecSlots];  // stream span of the batch: [s_first, s_end)
    uint32_t s_end[2][kDecSlots];
    // owned by the copier warp of the slot: output cursor, repeat offset, failed
    uint32_t d[kDecSlots], off[kDecSlots], dead[kDecSlots];
    uint8_t lut[256];  // first token byte -> header + literal bytes (bit 7: needs the slow path)
    // lexer state, one entry per slot, owned by the slot's lexer warp; positions are relative to
    // lx_abase, the 16-byte aligned address at or below the stream start (a = lead + s)
    unsigned long long lx_abase[kDecSlots];
    int lx_lead[kDecSlots], lx_aend[kDecSlots];
    int lx_fill[kDecSlots];          // requested up to here
    int lx_done[kDecSlots];          // token lengths written up to here
    // published to the parser, double buffered by round parity
    int lx_ready[2][kDecSlots];      // ring bytes below this have landed
    int lx_lexed[2][kDecSlots];      // adv[] entries below this are valid
};

constexpr size_t kDecSmemBytes = sizeof(uint16_t) * 2 * kDecSlots * kDescStride + 16 + sizeof(DecSlotState) +
                                 (size_t)kDecCopiers * (kDecStage + kDecLitStage + kDecScratch) +
                                 (size_t)kDecSlots * kDecRing * 2 + 64;

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

#ifndef MZ_DEC_BULK
#define MZ_DEC_BULK 1
#endif
// ---- bulk copies (the TMA engine's 1-D form, SASS UBLKCP) with mbarrier completion -------------
// One thread hands the copy engine a contiguous, 16-byte aligned stretch of global memory; the bytes
// land in shared memory without passing through registers and signal an mbarrier by byte count.
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n .reg .pred p;\n"
        "W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        " @!p bra W;\n}" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void *smem, const void *gmem, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem)),
                 "l"(gmem), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}

// Warp-cooperative forward copy dst[0..n) = src[0..n) with memmove-forward
// semantics for overlapping ranges (dst - src = off > 0), as decode.go:339-358.
__device__ __forceinline__ void warp_copy_overlap(uint8_t *dstp, uint32_t off, uint32_t n, int lane) {
    const uint8_t *from = dstp - off;
    if (off >= 32) {
        const bool overlap = off < n;
        for (uint32_t base = 0; base < n; base += 32) {
            uint32_t i = base + lane;
            if (i < n) dstp[i] = from[i];
            if (overlap) __syncwarp();
        }
    } else {
        uint32_t r = lane % off;
        const uint32_t step = 32 % off;
        for (uint32_t base = 0; base < n; base += 32) {
            uint32_t i = base + lane;
            if (i < n) dstp[i] = from[r];
            r += step;
            if (r >= off) r -= off;
        }
    }
}

// Warp-cooperative copy of non-overlapping bytes (literal runs).
__device__ __forceinline__ void warp_copy(uint8_t *dstp, const uint8_t *srcp, uint32_t n, int lane) {
    // head: until dst is 16-byte aligned
    uint32_t head = (uint32_t)((16 - (reinterpret_cast<uintptr_t>(dstp) & 15)) & 15);
    if (head > n) head = n;
    if ((uint32_t)lane < head) dstp[lane] = srcp[lane];
    const uint32_t body = (n - head) / 16;
    const uint8_t *sb = srcp + head;
    uint4 *db = reinterpret_cast<uint4 *>(dstp + head);
    const unsigned mis = (unsigned)(reinterpret_cast<uintptr_t>(sb) & 3);
    const uint32_t *sw = reinterpret_cast<const uint32_t *>(sb - mis);
    const unsigned sh = mis * 8;
    for (uint32_t i = lane; i < body; i += 32) {
        uint32_t w0 = sw[4 * i], w1 = sw[4 * i + 1], w2 = sw[4 * i + 2], w3 = sw[4 * i + 3];
        uint4 v = make_uint4(w0, w1, w2, w3);
        if (mis) {
            uint32_t w4 = sw[4 * i + 4];
            v = make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh),
                           __funnelshift_r(w3, w4, sh));
        }
        db[i] = v;
    }
    const uint32_t done = head + body * 16;
    if (done + lane < n) dstp[done + lane] = srcp[done + lane];
}

// Branch-free token parse (layouts: SPEC.md section 2; decode.go:195-308).
// One lane per block runs this in lockstep with 27+ other lanes, so there is
// no tag dispatch: every field is computed with selects.
struct PTok {
    uint32_t hdr, lit, mlen, off;
    bool isrep;
};
__device__ __forceinline__ PTok parse_token_bf(uint64_t w) {
    const uint32_t lo = (uint32_t)w;
    const uint32_t b0 = lo & 0xff;
    const uint32_t tag = lo & 3;
    const bool t0 = tag == 0, t1 = tag == 1, t2 = tag == 2;
    const bool bit2 = (lo & 4) != 0;
    const bool c3 = tag == 3 && bit2;   // copy3
    const bool fz = tag == 3 && !bit2;  // fused copy2
    // length code, and how many extension bytes follow the fixed header
    const uint32_t code = t0 ? b0 >> 3 : t1 ? (b0 >> 2) & 15 : t2 ? b0 >> 2 : (lo >> 5) & 63;
    const uint32_t thr = t0 ? 28u : t1 ? 14u : 60u;
    uint32_t ext = code > thr ? code - thr : 0;
    if (fz) ext = 0;
    const uint32_t fixed = t0 ? 1u : t1 ? 2u : c3 ? 4u : 3u;
    const uint32_t E = (uint32_t)(w >> (8 * fixed)) & (0xffffffu >> (8 * (3 - ext)));
    const uint32_t addb = t0 ? 30u : t1 ? 18u : 64u;
    uint32_t len = ext ? E + addb : code + (t0 ? 1u : 4u);
    if (fz) len = 4 + ((lo >> 5) & 7);
    const bool islit = t0 && !bit2;
    PTok t;
    t.isrep = t0 && bit2;
    t.hdr = fixed + ext;
    t.lit = islit ? len : fz ? ((lo >> 3) & 3) + 1 : c3 ? (lo >> 3) & 3 : 0;
    t.mlen = islit ? 0 : len;
    t.off = t1 ? ((lo & 0xffff) >> 6) + 1 : c3 ? (lo >> 11) + kMinCopy3Offset : ((lo >> 8) & 0xffff) + kMinCopy2Offset;
    return t;
}

// Copies n (0..16) bytes between two shared-memory regions at arbitrary byte
// alignment with word accesses: 5 aligned loads + funnel shifts on the source,
// up to 3 head bytes, up to 4 aligned words and up to 3 tail bytes on the
// destination.  Byte ranges written by different lanes never overlap, and full
// words are only stored inside a lane's own range, so lanes do not race.
// `sbase` / `dbase` must be 4-byte aligned; both regions need 4 bytes of slack.
__device__ __forceinline__ void smem_put16(const uint8_t *sbase, uint32_t soff, uint8_t *dbase, uint32_t doff, uint32_t n) {
    if (n == 0) soff = doff = 0;  // idle lanes still execute the (unconditional) loads: keep them in bounds
    const uint32_t *sw = reinterpret_cast<const uint32_t *>(sbase) + (soff >> 2);
    const unsigned sh = (soff & 3) * 8;
    const uint32_t a0 = sw[0], a1 = sw[1], a2 = sw[2], a3 = sw[3], a4 = sw[4];
    const uint32_t S0 = __funnelshift_r(a0, a1, sh), S1 = __funnelshift_r(a1, a2, sh), S2 = __funnelshift_r(a2, a3, sh),
                   S3 = __funnelshift_r(a3, a4, sh);
    const uint32_t head = min(n, (4u - (doff & 3u)) & 3u);
    uint8_t *dp = dbase + doff;
    if (head > 0) dp[0] = (uint8_t)S0;
    if (head > 1) dp[1] = (uint8_t)(S0 >> 8);
    if (head > 2) dp[2] = (uint8_t)(S0 >> 16);
    const unsigned hs = head * 8;
    const uint32_t D0 = __funnelshift_r(S0, S1, hs), D1 = __funnelshift_r(S1, S2, hs), D2 = __funnelshift_r(S2, S3, hs),
                   D3 = S3 >> hs;
    const uint32_t rem = n - head;
    const uint32_t nw = rem >> 2, tail = rem & 3;
    uint32_t *dw = reinterpret_cast<uint32_t *>(dp + head);
    if (nw > 0) dw[0] = D0;
    if (nw > 1) dw[1] = D1;
    if (nw > 2) dw[2] = D2;
    if (nw > 3) dw[3] = D3;
    const uint32_t Dt = nw == 0 ? D0 : nw == 1 ? D1 : nw == 2 ? D2 : D3;
    uint8_t *tp = reinterpret_cast<uint8_t *>(dw + nw);
    if (tail > 0) tp[0] = (uint8_t)Dt;
    if (tail > 1) tp[1] = (uint8_t)(Dt >> 8);
    if (tail > 2) tp[2] = (uint8_t)(Dt >> 16);
}

__device__ __forceinline__ uint8_t dec_lut_entry(uint32_t b0) {
    const uint32_t tag = b0 & 3;
    if (tag == 0) {
        const uint32_t x = b0 >> 3;
        if (x >= 29) return 0x80;                                 // extended length
        return (uint8_t)((b0 & 4) ? 1 : 1 + x + 1);               // repeat: 1 byte; literal: header + x+1 bytes
    }
    if (tag == 1) return ((b0 >> 2) & 15) == 15 ? 0x80 : 2;
    if (tag == 2) return (b0 >> 2) > 60 ? 0x80 : 3;
    if ((b0 & 4) == 0) return (uint8_t)(3 + ((b0 >> 3) & 3) + 1);  // fused copy2 + 1..4 literals
    return (uint8_t)(4 + ((b0 >> 3) & 3));                         // copy3 + 0..3 literals (extension checked inline)
}

__global__ void __launch_bounds__(kDecThreads, 1)
decode_pc_kernel(int nblk, int slots_per_cta, const uint8_t *__restrict__ src, const uint64_t *__restrict__ sbeg,
                 const uint64_t *__restrict__ send, uint8_t *dst, const uint64_t *__restrict__ dbeg,
                 const uint64_t *__restrict__ dend, int32_t *__restrict__ status) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    // [2][kDecSlots][kDescStride]: offset of every token of the batch from the batch's first byte (| kDescGlobal)
    uint16_t *desc = reinterpret_cast<uint16_t *>(smem_raw);
    DecSlotState *st = reinterpret_cast<DecSlotState *>(smem_raw + ((sizeof(uint16_t) * 2 * kDecSlots * kDescStride + 15) & ~size_t(15)));
    uint8_t *copier_mem = reinterpret_cast<uint8_t *>(st + 1);
    copier_mem += (16 - (reinterpret_cast<uintptr_t>(copier_mem) & 15)) & 15;
    __shared__ int produced[2];  // produced[r & 1]: the parser emitted something in round r
    __shared__ __align__(8) uint64_t lexer_bar[kDecLexers];  // bulk-copy completion, one phase per round

    const int lane = lane_id();
    const int warp = threadIdx.x >> 5;
    const int first_blk = blockIdx.x * slots_per_cta;
    const int nslots = min(slots_per_cta, nblk - first_blk);

    // ---- parser lane state (warp 0) ----
    // The parser reads its stream and the token lengths through per-slot shared-memory rings
    // kept filled by the lexer warps (a step never waits on DRAM: with 28 lanes in lockstep
    // some lane would miss the cache on every step).  Ring positions are relative to the
    // 16-byte aligned address at or below the stream start: a = lead + s.
    uint8_t *rings = copier_mem + (size_t)kDecCopiers * (kDecStage + kDecLitStage + kDecScratch);  // [kDecSlots][kDecRing]
    uint8_t *advs = rings + (size_t)kDecSlots * kDecRing;                                          // [kDecSlots][kDecRing]
    const uint8_t *p_sp = nullptr;
    int p_slen = 0, p_s = 0;
    bool p_done = true;
    int p_lead = 0, p_aend = 0;
    const uint8_t *p_ring = rings + (size_t)lane * kDecRing;
    const uint8_t *p_adv = advs + (size_t)lane * kDecRing;
    if (warp == 0 && lane < nslots) {
        const int b = first_blk + lane;
        p_sp = src + sbeg[b];
        p_slen = (int)(send[b] - sbeg[b]);
        p_done = false;
        const uintptr_t abase = reinterpret_cast<uintptr_t>(p_sp) & ~uintptr_t(15);
        p_lead = (int)(reinterpret_cast<uintptr_t>(p_sp) - abase);
        p_aend = p_lead + p_slen;
        st->lx_abase[lane] = abase;
        st->lx_lead[lane] = p_lead;
        st->lx_aend[lane] = p_aend;
        st->lx_fill[lane] = 0;
        st->lx_done[lane] = 0;
        st->lx_ready[0][lane] = st->lx_ready[1][lane] = 0;
        st->lx_lexed[0][lane] = st->lx_lexed[1][lane] = 0;
    }
    if (threadIdx.x < kDecLexers) {
        mbar_init(&lexer_bar[threadIdx.x], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");  // visible to the copy engine and to the other warps
    }
    if (threadIdx.x < 256) st->lut[threadIdx.x] = dec_lut_entry(threadIdx.x);
    if (threadIdx.x < kDecSlots) {
        st->count[0][threadIdx.x] = st->count[1][threadIdx.x] = 0;
        st->s_end[0][threadIdx.x] = st->s_end[1][threadIdx.x] = 0;
        st->s_first[0][threadIdx.x] = st->s_first[1][threadIdx.x] = 0;
        st->d[threadIdx.x] = 0;
        st->off[threadIdx.x] = 1;  // decode.go:186
        st->dead[threadIdx.x] = 0;
    }
    __syncthreads();

    for (int round = 0;; round++) {
        const int wb = round & 1;  // buffer the parser fills
        const int rb = wb ^ 1;     // buffer the copiers drain (filled in the previous round)
        if (warp == 0) {
            // ================= PARSER =================
            uint16_t *my = desc + (wb * kDecSlots + min(lane, kDecSlots - 1)) * kDescStride;
            // what the lexers published at the end of the previous round
            const int p_ready = lane < nslots ? st->lx_ready[rb][lane] : 0;
            const int p_lexed = lane < nslots ? st->lx_lexed[rb][lane] : 0;
            int cnt = 0;
            uint32_t flags = 0;
            const int s_first = p_s;
            // end of stream is noticed at the start of a round (decode.go:615 is checked
            // by the copier that owns the output cursor)
            if (!p_done && p_s >= p_slen) {
                p_done = true;
                flags = kBatchEnded;
            }
            bool stop = p_done;  // no more tokens from this lane in this round
            // One token per lane per step, and as little as possible per step: every instruction
            // here is on the serial chain of 28 blocks.  The chain is  a -> adv[a] -> a + adv.
            // Tokens the lexers left alone (extended lengths), bytes not lexed yet and tokens
            // that overrun the stream take the cold path.
#pragma unroll 4
            for (int k = 0; k < kDecTok; k++) {
                const bool live = !stop && p_s < p_slen;
                const int a = p_lead + p_s;
                int adv = 0;
                if (live && a < p_lexed) adv = p_adv[a & (kDecRing - 1)];
                if (live) {
                    if (adv != 0 && adv <= p_slen - p_s) {
                        my[cnt] = (uint16_t)(p_s - s_first);
                        p_s += adv;
                        cnt++;
                    } else {
                        // ---- cold path ----
                        uint32_t where = 0;
                        uint64_t w8;
                        if (a + 12 <= p_ready || p_aend <= p_ready) {
                            const uint32_t *rw = reinterpret_cast<const uint32_t *>(p_ring);
                            const uint32_t a4 = (uint32_t)a >> 2;
                            const uint32_t w0 = rw[a4 & (kDecRing / 4 - 1)], w1 = rw[(a4 + 1) & (kDecRing / 4 - 1)],
                                           w2 = rw[(a4 + 2) & (kDecRing / 4 - 1)];
                            const unsigned sh = ((unsigned)a & 3u) * 8;
                            w8 = (uint64_t)__funnelshift_r(w1, w2, sh) << 32 | __funnelshift_r(w0, w1, sh);
                        } else {  // first rounds / after a long literal run: not in the ring (yet)
                            w8 = ldg_window(p_sp, p_s, p_slen);
                            where = kDescGlobal;
                        }
                        const uint32_t lo = (uint32_t)w8;
                        const uint32_t e = st->lut[lo & 0xff];
                        adv = (int)(e & 0x7f);
                        bool lng = false;
                        if ((e & 0x80) != 0 || ((lo & 7) == 7 && ((lo >> 5) & 63) > 60)) {  // extended length
                            const PTok t = parse_token_bf(w8);
                            adv = (int)(t.hdr + t.lit);
                            lng = t.lit > kDecShort || t.mlen > kDecShort;
                        }
                        // header and literals must lie inside the stream (decode.go:221,410 src side)
                        if (adv > p_slen - p_s) {
                            flags |= kBatchBad;
                            p_done = true;
                            stop = true;
                        } else if (lng && cnt > 0) {
                            stop = true;  // a long token travels alone: it starts the next batch
                        } else if ((uint32_t)(p_s - s_first) >= kDescGlobal) {
                            stop = true;  // the 15-bit offset is used up: next batch (cannot happen with 32 short tokens)
                        } else {
                            my[cnt] = (uint16_t)((uint32_t)(p_s - s_first) | where);
                            p_s += adv;
                            cnt++;
                            if (lng) {
                                flags |= kBatchLong;
                                stop = true;
                            }
                        }
                    }
                }
            }
            if (lane < kDecSlots) {
                st->count[wb][lane] = (uint32_t)cnt | flags;
                st->s_first[wb][lane] = (uint32_t)s_first;
                st->s_end[wb][lane] = (uint32_t)p_s;
            }
            // a round that emits nothing (no tokens, no end / error notices) means
            // every block of this CTA is finished
            const bool some = __any_sync(kFullMask, cnt > 0 || flags != 0);
            if (lane == 0) produced[wb] = some ? 1 : 0;
        } else if (warp > kDecCopiers) {
            // ================= LEXERS =================
            // pass 1: request the next stretch of every slot of mine.  The ring keeps the batch the
            // copiers work on in this round (they read the token headers from it) and runs ahead of it.
            uint64_t *lx_bar = &lexer_bar[warp - 1 - kDecCopiers];
            unsigned lx_tx = 0;
            for (int slot = warp - 1 - kDecCopiers; slot < nslots; slot += kDecLexers) {
                const int lead = st->lx_lead[slot], aend = st->lx_aend[slot];
                int afill = st->lx_fill[slot];
                const int keep = lead + (int)st->s_first[rb][slot];  // first byte of the batch being copied
                const int cur = lead + (int)st->s_end[rb][slot];     // the parser's cursor
                if (afill < (cur & ~15)) afill = cur & ~15;          // a long literal run skipped ahead of everything requested
                const int want = min((aend + 15) & ~15, (keep + kDecRingAhead) & ~15);
                uint8_t *ring = rings + (size_t)slot * kDecRing;
                const uintptr_t abase = (uintptr_t)st->lx_abase[slot];
#if MZ_DEC_BULK
                // [afill, want) is contiguous in memory and 16-byte aligned at both ends: one bulk copy,
                // two when it wraps around the ring
                if (lane == 0 && want > afill) {
                    const unsigned o = (unsigned)afill & (kDecRing - 1);
                    const unsigned len = (unsigned)(want - afill);
                    const unsigned first = min(len, (unsigned)kDecRing - o);
                    bulk_copy_g2s(ring + o, reinterpret_cast<const void *>(abase + (uintptr_t)afill), first, lx_bar);
                    if (len > first)
                        bulk_copy_g2s(ring, reinterpret_cast<const void *>(abase + (uintptr_t)afill + first), len - first, lx_bar);
                    lx_tx += len;
                }
#else  // the same bytes with per-lane 16-byte cp.async (kept for A/B and for racecheck, which models these)
                for (int chunk = afill + 16 * lane; chunk < want; chunk += 16 * 32)
                    cp_async16(ring + (chunk & (kDecRing - 1)), reinterpret_cast<const void *>(abase + (uintptr_t)chunk));
#endif
                if (want > afill) afill = want;
                __syncwarp();  // every lane has read lx_fill[slot]
                if (lane == 0) st->lx_fill[slot] = afill;
            }
            // rounds are long (the copiers): wait for the bytes and lex them right away
#if MZ_DEC_BULK
            if (lane == 0) mbar_arrive_expect_tx(lx_bar, lx_tx);
            mbar_wait(lx_bar, (unsigned)round & 1u);
#else
            (void)lx_bar;
            (void)lx_tx;
            cp_async_wait_all();
#endif
            __syncwarp();  // lane 0's lx_fill stores are read by every lane below
            // pass 2: token lengths.  adv[i] = header + literal bytes of a token that starts at byte i,
            // 0 when the length is extended (parser's cold path).  An entry needs bytes i, i+1; four
            // offsets per lane per step from two aligned words.
            for (int slot = warp - 1 - kDecCopiers; slot < nslots; slot += kDecLexers) {
                const int lead = st->lx_lead[slot], aend = st->lx_aend[slot];
                const int afill = st->lx_fill[slot];
                int done = st->lx_done[slot];
                const int cur = lead + (int)st->s_end[rb][slot];
                if (done < (cur & ~15)) done = cur & ~15;  // bytes behind the cursor are dead
                const uint32_t *ring = reinterpret_cast<const uint32_t *>(rings + (size_t)slot * kDecRing);
                uint32_t *adv = reinterpret_cast<uint32_t *>(advs + (size_t)slot * kDecRing);
                // The frontier stays 12 bytes behind what has landed: parser and copiers take the 8 header
                // bytes of a lexed token from the ring with three aligned word loads (<= 11 bytes from its
                // start), and those must never touch cells the copy engine is filling.  It moves in whole
                // words (an entry word is written once, never while the parser may read it) until it
                // reaches the end of the stream.
                const int lim = afill >= aend ? aend : (afill - 12) & ~3;
                if (lim > done)  // (an entry word is never rewritten: the parser may be reading it)
                for (int i = (done & ~3) + 4 * lane; i < lim; i += 4 * 32) {
                    const uint32_t w0 = ring[(i >> 2) & (kDecRing / 4 - 1)], w1 = ring[((i >> 2) + 1) & (kDecRing / 4 - 1)];
                    uint32_t out = 0;
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const uint32_t b01 = __funnelshift_r(w0, w1, 8 * j) & 0xffffu;  // bytes i+j, i+j+1
                        uint32_t e = st->lut[b01 & 0xff];
                        if ((e & 0x80) != 0 || ((b01 & 7) == 7 && ((b01 >> 5) & 63) > 60)) e = 0;
                        out |= e << (8 * j);
                    }
                    adv[(i >> 2) & (kDecRing / 4 - 1)] = out;
                }
                if (lim > done) done = lim;
                __syncwarp();
                if (lane == 0) {
                    st->lx_done[slot] = done;
                    st->lx_ready[wb][slot] = afill;
                    st->lx_lexed[wb][slot] = done;
                }
            }
        } else if (round > 0) {
            // ================= COPIERS =================
            const int cw = warp - 1;
            uint8_t *stage = copier_mem + (size_t)cw * (kDecStage + kDecLitStage + kDecScratch);
            uint8_t *lstage = stage + kDecStage;
            uint8_t *scratch = lstage + kDecLitStage;
#ifdef MZ_EXPERIMENT_NO_COPY
            for (int slot = nslots; slot < nslots; slot += kDecCopiers) {
#else
            for (int slot = cw; slot < nslots; slot += kDecCopiers) {
#endif
                const uint32_t cword = st->count[rb][slot];
                if (cword == 0 || st->dead[slot]) continue;
                int n = (int)(cword & 0xff);
                const int b = first_blk + slot;
                const uint8_t *sp = src + sbeg[b];
                uint8_t *dp = dst + dbeg[b];
                const uint32_t dlen = (uint32_t)(dend[b] - dbeg[b]);
                const uint32_t d_base = st->d[slot];
                const uint32_t off_carry = st->off[slot];
                // ---- decode my token ----
                const uint16_t *dsc = desc + (rb * kDecSlots + slot) * kDescStride;
                const uint32_t slen = (uint32_t)(send[b] - sbeg[b]);
                uint32_t tp = 0, lit = 0, mlen = 0, off_tok = 0, hdr = 0;
                bool isrep = false;
                if (lane < n) {
                    const uint32_t dw = dsc[lane];
                    tp = st->s_first[rb][slot] + (dw & (kDescGlobal - 1));
                    uint64_t w8;
                    if (dw & kDescGlobal) {  // parsed before its bytes were in the ring
                        w8 = ldg_window(sp, tp, slen);
                    } else {  // the 8 header bytes from the slot's ring (kept by the lexers while this batch is copied)
                        const uint32_t *rw = reinterpret_cast<const uint32_t *>(rings + (size_t)slot * kDecRing);
                        const uint32_t ap = (uint32_t)st->lx_lead[slot] + tp;
                        const uint32_t a4 = ap >> 2;
                        const uint32_t w0 = rw[a4 & (kDecRing / 4 - 1)], w1 = rw[(a4 + 1) & (kDecRing / 4 - 1)],
                                       w2 = rw[(a4 + 2) & (kDecRing / 4 - 1)];
                        const unsigned sh = (ap & 3u) * 8;
                        w8 = (uint64_t)__funnelshift_r(w1, w2, sh) << 32 | __funnelshift_r(w0, w1, sh);
                    }
                    const PTok t = parse_token_bf(w8);
                    lit = t.lit;
                    mlen = t.mlen;
                    off_tok = t.off;
                    hdr = t.hdr;
                    isrep = t.isrep;
                }
                const uint32_t litpos = tp + hdr;
                // output positions: exclusive prefix sum of lit + mlen
                const uint32_t olen = lit + mlen;
                uint32_t incl = olen;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t u = __shfl_up_sync(kFullMask, incl, o);
                    if (lane >= o) incl += u;
                }
                const uint32_t dpos = d_base + incl - olen;
                // repeat offsets: the last offset defined before me (decode.go:186,214)
                uint32_t lastdef = (mlen != 0 && !isrep) ? off_tok : 0;  // 0 = none (offsets are >= 1)
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t u = __shfl_up_sync(kFullMask, lastdef, o);
                    if (lane >= o && lastdef == 0) lastdef = u;
                }
                uint32_t prevdef = __shfl_up_sync(kFullMask, lastdef, 1);
                if (lane == 0 || prevdef == 0) prevdef = lane == 0 ? off_carry : prevdef;
                if (prevdef == 0) prevdef = off_carry;
                const uint32_t off = isrep ? prevdef : off_tok;
                // the reference's dst-side checks (decode.go:221,326,410,572), in stream order
                bool tbad = false;
                if (lane < n) {
                    tbad = dpos > dlen || lit > dlen - dpos;
                    if (!tbad && mlen != 0) tbad = off > dpos + lit || mlen > dlen - dpos - lit;
                }
                const unsigned badm = __ballot_sync(kFullMask, tbad);
                bool kill = (cword & kBatchBad) != 0;
                if (badm) {  // keep the tokens before the first bad one, then give up on the block
                    n = __ffs(badm) - 1;
                    kill = true;
                    if (lane >= n) lit = mlen = 0;
                }
                const uint32_t total = n > 0 ? __shfl_sync(kFullMask, incl, n - 1) : 0;
                {
                    // carry the cursor and the repeat offset to the next batch of this block
                    const unsigned defm = __ballot_sync(kFullMask, lane < n && mlen != 0 && !isrep);
                    const uint32_t newoff = defm ? __shfl_sync(kFullMask, off_tok, 31 - __clz(defm)) : off_carry;
                    __syncwarp();  // every lane has read st->d / st->off / st->dead of this slot
                    if (lane == 0) {
                        st->d[slot] = d_base + total;
                        st->off[slot] = newoff;
                        if ((cword & kBatchEnded) && d_base + total != dlen) kill = true;  // decode.go:615
                        if (kill) st->dead[slot] = 1;
                    }
                }
                if (n == 0) continue;
                if (cword & kBatchLong) {
                    // ---- one long token: cooperative copies straight to global memory ----
                    const uint32_t l_litpos = __shfl_sync(kFullMask, litpos, 0);
                    const uint32_t l_dpos = __shfl_sync(kFullMask, dpos, 0);
                    const uint32_t l_lit = __shfl_sync(kFullMask, lit, 0);
                    const uint32_t l_mlen = __shfl_sync(kFullMask, mlen, 0);
                    const uint32_t l_off = __shfl_sync(kFullMask, off, 0);
                    if (l_lit) warp_copy(dp + l_dpos, sp + l_litpos, l_lit, lane);
                    __syncwarp();
                    if (l_mlen) warp_copy_overlap(dp + l_dpos + l_lit, l_off, l_mlen, lane);
                    __syncwarp();
                    continue;
                }
                // ---- batch of short tokens ----
                const uint32_t x0 = d_base;
                const uint32_t xend = d_base + total;
                const uint32_t shift = (uint32_t)(reinterpret_cast<uintptr_t>(dp + x0) & 15);  // stage mirrors dst alignment
                // 1. stream span -> lstage (aligned 16-byte chunks of the absolute address)
                const uint32_t s_first = st->s_first[rb][slot], s_end = st->s_end[rb][slot];
                const uint8_t *span0 = sp + s_first;
                const uintptr_t a0 = reinterpret_cast<uintptr_t>(span0) & ~uintptr_t(15);
                const uint32_t lshift = (uint32_t)(reinterpret_cast<uintptr_t>(span0) - a0);
                const uint32_t span_chunks = (lshift + (s_end - s_first) + 15) / 16;
                for (uint32_t c = lane; c < span_chunks; c += 32)
                    cp_async16(lstage + 16 * c, reinterpret_cast<const void *>(a0 + 16 * (uintptr_t)c));
                // The image is flushed in whole 16-byte chunks.  Its first chunk also holds
                // the last `shift` bytes of the previous batch: fetch them back (not at the
                // very start of a block, where those bytes belong to a neighbour).
                const bool head_ok = x0 > 0;
                if (lane == 0 && shift != 0 && head_ok) cp_async16(stage, dp + x0 - shift);
                asm volatile("cp.async.commit_group;\n" ::: "memory");  // group 1: stream span + head chunk
                // 2. back-reference gathers of tokens whose source is complete (ends before
                //    this span): up to 32 source bytes = three aligned 16-byte chunks per lane
                const uint32_t m = dpos + lit;    // output position of the match
                const uint32_t srcpos = m - off;  // its source
                const bool has_m = lane < n && mlen > 0;
                const bool indep = has_m && srcpos + mlen <= x0;
                const bool dep = has_m && !indep;
                uint8_t *myscr = scratch + lane * 48;
                uint32_t g15 = 0;
                if (indep) {
                    const uint8_t *g = dp + srcpos;
                    const uintptr_t ga = reinterpret_cast<uintptr_t>(g) & ~uintptr_t(15);
                    g15 = (uint32_t)(reinterpret_cast<uintptr_t>(g) - ga);
                    const uint32_t need = g15 + min(mlen, 32u);
                    cp_async16(myscr, reinterpret_cast<const void *>(ga));
                    if (need > 16) cp_async16(myscr + 16, reinterpret_cast<const void *>(ga + 16));
                    if (need > 32) cp_async16(myscr + 32, reinterpret_cast<const void *>(ga + 32));
                }
                asm volatile("cp.async.commit_group;\n" ::: "memory");  // group 2: back-reference gathers
                asm volatile("cp.async.wait_group 1;\n" ::: "memory");   // the span has landed; gathers still fly
                __syncwarp();
                const uint32_t T = shift + (dpos - x0);  // stage offset of this token's output
                // 3. literals: lstage -> stage, 16 bytes per pass
                {
                    const uint32_t loff = lshift + (litpos - s_first);
                    for (uint32_t done = 0;; done += 16) {
                        const uint32_t nb = lit > done ? min(lit - done, 16u) : 0;
                        if (!__any_sync(kFullMask, nb != 0)) break;
                        smem_put16(lstage, loff + done, stage, T + done, nb);
                    }
                }
                // 4. independent matches: scratch -> stage; 32 source bytes per gather
                cp_async_wait_all();
                __syncwarp();
                {
                    const uint32_t Tm = T + lit;
                    for (uint32_t done = 0;;) {
                        const uint32_t n0 = indep && mlen > done ? min(mlen - done, 16u) : 0;
                        smem_put16(myscr, g15, stage, Tm + done, n0);
                        const uint32_t n1 = indep && mlen > done + 16 ? min(mlen - done - 16, 16u) : 0;
                        if (__any_sync(kFullMask, n1 != 0)) smem_put16(myscr, g15 + 16, stage, Tm + done + 16, n1);
                        done += 32;
                        const bool more = indep && mlen > done;
                        if (!__any_sync(kFullMask, more)) break;
                        __syncwarp();
                        if (more) {  // next 32 source bytes
                            const uint8_t *g = dp + srcpos + done;
                            const uintptr_t ga = reinterpret_cast<uintptr_t>(g) & ~uintptr_t(15);
                            g15 = (uint32_t)(reinterpret_cast<uintptr_t>(g) - ga);
                            const uint32_t need = g15 + min(mlen - done, 32u);
                            cp_async16(myscr, reinterpret_cast<const void *>(ga));
                            if (need > 16) cp_async16(myscr + 16, reinterpret_cast<const void *>(ga + 16));
                            if (need > 32) cp_async16(myscr + 32, reinterpret_cast<const void *>(ga + 32));
                        }
                        cp_async_wait_all();
                        __syncwarp();
                    }
                }
                __syncwarp();
                // 5. dependent matches, in stream order, warp-cooperative
                unsigned dm = __ballot_sync(kFullMask, dep);
                while (dm) {
                    const int t = __ffs(dm) - 1;
                    dm &= dm - 1;
                    const uint32_t tm = __shfl_sync(kFullMask, m, t);
                    const uint32_t tsrc = __shfl_sync(kFullMask, srcpos, t);
                    const uint32_t tlen = __shfl_sync(kFullMask, mlen, t);
                    const uint32_t toff = __shfl_sync(kFullMask, off, t);
                    if (toff >= tlen) {
                        for (uint32_t i = lane; i < tlen; i += 32) {
                            const uint32_t q = tsrc + i;
                            const uint8_t v = q >= x0 ? stage[shift + (q - x0)] : dp[q];
                            stage[shift + (tm - x0) + i] = v;
                        }
                    } else {  // overlapping: replicate the `toff`-byte pattern
                        for (uint32_t i = lane; i < tlen; i += 32) {
                            const uint32_t q = tsrc + i % toff;
                            const uint8_t v = q >= x0 ? stage[shift + (q - x0)] : dp[q];
                            stage[shift + (tm - x0) + i] = v;
                        }
                    }
                    __syncwarp();
                }
                // 6. flush the image: stage[shift .. shift + (xend - x0)) -> dp[x0 .. xend)
                {
                    const uint32_t total = shift + (xend - x0);
                    uint8_t *gbase = dp + x0 - shift;  // 16-byte aligned
                    const uint32_t nchunk = (total + 15) / 16;
                    for (uint32_t c = lane; c < nchunk; c += 32) {
                        const uint32_t lo = 16 * c, hi = lo + 16;
                        // bytes past `total` in the last chunk are this block's future output
                        // (rewritten by the next batch), so a whole-chunk store is fine unless
                        // it would cross the end of the block
                        if ((lo >= shift || head_ok) && (hi <= total || x0 - shift + hi <= dlen)) {
                            *reinterpret_cast<uint4 *>(gbase + lo) = *reinterpret_cast<const uint4 *>(stage + lo);
                        } else {
                            for (uint32_t i = max(lo, shift); i < min(hi, total); i++) gbase[i] = stage[i];
                        }
                    }
                }
                __syncwarp();
            }
        }
        __syncthreads();
        // Round `round` emitted nothing: nothing is left to parse, and the copiers
        // drained round-1 in this very round.  (produced[] is double buffered, so
        // the parser's next write cannot race with this read.)
        if (produced[wb] == 0) break;
    }

    if (threadIdx.x < nslots) status[first_blk + threadIdx.x] = st->dead[threadIdx.x] ? 1 : 0;
}

}  // namespace mz
