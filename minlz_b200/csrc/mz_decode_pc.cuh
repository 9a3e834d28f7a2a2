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
//   * COPIER warps: lane k owns token k of a batch.  All 32 back-reference
//     gathers of a batch go out together (16-byte cp.async per lane into
//     shared memory), literals come from one coalesced load of the batch's
//     stream span, everything is assembled in a shared-memory image of the
//     batch's output span and leaves with coalesced 16-byte stores.  Tokens
//     whose source lies inside the span being assembled are applied afterwards
//     in stream order, warp-cooperatively (overlapping copies replicate the
//     `offset`-byte pattern).
//   * parser and copiers run bulk-synchronously: round r+1 is parsed while
//     round r is copied (double-buffered descriptors, one __syncthreads per
//     round), so there is no fine-grained inter-warp signalling.
//
// Result contract = the reference's: status 0 and exactly dst_len bytes, or
// status 1 (decodeErrCodeCorrupt); a corrupt block never writes outside its
// own dst range (tokens are validated before their descriptors are emitted).
#pragma once

#include "mz_common.cuh"
#include "mz_decode.cuh"

namespace mz {

constexpr int kDecSlots = 32;      // block slots per CTA (one parser lane each)
constexpr int kDecCopiers = 14;    // copier warps per CTA
constexpr int kDecThreads = (1 + kDecCopiers) * 32;
constexpr int kDecTok = 32;        // tokens per batch
constexpr int kDecShort = 64;      // max literal / match length of a "short" token
constexpr int kDecStage = kDecTok * 2 * kDecShort + 32;       // output image of a batch (+ alignment slack)
constexpr int kDecLitStage = kDecTok * (8 + kDecShort) + 48;  // stream span of a batch
constexpr int kDecScratch = 32 * 48;                          // per-lane 32-byte gather landing zone (48 B stride)
constexpr int kDescStride = 5 * kDecTok + 1;                  // words per (slot, buffer); odd -> conflict free

struct DecSlotState {  // written by the parser, read by copiers after the barrier
    uint32_t count[2][kDecSlots];   // tokens in the batch | 0x100 if it is a single long token
    uint32_t s_first[2][kDecSlots]; // stream span of the batch: [s_first, s_end)
    uint32_t s_end[2][kDecSlots];
};

constexpr size_t kDecSmemBytes = sizeof(uint32_t) * 2 * kDecSlots * kDescStride + sizeof(DecSlotState) +
                                 (size_t)kDecCopiers * (kDecStage + kDecLitStage + kDecScratch) + 64;

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

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

__global__ void __launch_bounds__(kDecThreads, 1)
decode_pc_kernel(int nblk, int slots_per_cta, const uint8_t *__restrict__ src, const uint64_t *__restrict__ sbeg,
                 const uint64_t *__restrict__ send, uint8_t *dst, const uint64_t *__restrict__ dbeg,
                 const uint64_t *__restrict__ dend, int32_t *__restrict__ status) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint32_t *desc = reinterpret_cast<uint32_t *>(smem_raw);  // [2][kDecSlots][kDescStride]
    DecSlotState *st = reinterpret_cast<DecSlotState *>(desc + 2 * kDecSlots * kDescStride);
    uint8_t *copier_mem = reinterpret_cast<uint8_t *>(st + 1);
    copier_mem += (16 - (reinterpret_cast<uintptr_t>(copier_mem) & 15)) & 15;
    __shared__ int produced[2];  // produced[r & 1]: the parser emitted tokens in round r

    const int lane = lane_id();
    const int warp = threadIdx.x >> 5;
    const int first_blk = blockIdx.x * slots_per_cta;
    const int nslots = min(slots_per_cta, nblk - first_blk);

    // ---- parser lane state (warp 0) ----
    const uint8_t *p_sp = nullptr;
    int64_t p_slen = 0, p_dlen = 0, p_s = 0, p_d = 0;
    uint32_t p_off = 1;
    bool p_done = true, p_bad = false;
    if (warp == 0 && lane < nslots) {
        const int b = first_blk + lane;
        p_sp = src + sbeg[b];
        p_slen = (int64_t)(send[b] - sbeg[b]);
        p_dlen = (int64_t)(dend[b] - dbeg[b]);
        p_done = false;
    }
    if (threadIdx.x < 2 * kDecSlots) st->count[threadIdx.x / kDecSlots][threadIdx.x % kDecSlots] = 0;
    __syncthreads();

    for (int round = 0;; round++) {
        const int wb = round & 1;        // buffer the parser fills
        const int rb = wb ^ 1;           // buffer the copiers drain (filled in the previous round)
        if (warp == 0) {
            // ================= PARSER =================
            uint32_t *my = desc + (wb * kDecSlots + lane) * kDescStride;
            uint32_t cnt = 0;
            bool cut = false, longtok = false;
            const uint32_t s_first = (uint32_t)p_s;
            for (int k = 0; k < kDecTok; k++) {
                const bool act = !p_done && !cut;
                if (!__any_sync(kFullMask, act)) break;
                if (act) {
                    if (p_s >= p_slen) {
                        p_done = true;  // end of stream: decode.go:615 checks d == len(dst)
                        if (p_d != p_dlen) p_bad = true;
                    } else {
                        const Token t = parse_token(ldg_window(p_sp, p_s, p_slen));
                        const int64_t s1 = p_s + t.hdr;
                        bool bad = s1 > p_slen;
                        if (!bad && t.lit) bad = (int64_t)t.lit > p_dlen - p_d || (int64_t)t.lit > p_slen - s1;
                        uint32_t off = t.repeat ? p_off : t.off;
                        if (!bad && t.mlen)
                            bad = (int64_t)off > p_d + t.lit || (int64_t)t.mlen > p_dlen - p_d - t.lit;
                        if (bad) {
                            p_bad = true;
                            p_done = true;
                        } else {
                            const bool lng = t.lit > kDecShort || t.mlen > kDecShort;
                            if (lng && cnt > 0) {
                                cut = true;  // a long token travels alone: it starts the next batch
                            } else {
                                my[0 * kDecTok + cnt] = (uint32_t)s1;   // literal source (stream position)
                                my[1 * kDecTok + cnt] = (uint32_t)p_d;  // output position
                                my[2 * kDecTok + cnt] = t.lit;
                                my[3 * kDecTok + cnt] = t.mlen;
                                my[4 * kDecTok + cnt] = off;
                                cnt++;
                                p_s = s1 + t.lit;
                                p_d += (int64_t)t.lit + t.mlen;
                                if (t.mlen) p_off = off;
                                if (lng) {
                                    longtok = true;
                                    cut = true;
                                }
                            }
                        }
                    }
                }
            }
            if (lane < kDecSlots) {
                st->count[wb][lane] = cnt | (longtok ? 0x100u : 0u);
                st->s_first[wb][lane] = s_first;
                st->s_end[wb][lane] = (uint32_t)p_s;
            }
            // an empty batch means every block of this CTA is finished (a lane that is
            // not done always emits at least one token per round)
            const bool some = __any_sync(kFullMask, cnt > 0);
            if (lane == 0) produced[wb] = some ? 1 : 0;
        } else if (round > 0) {
            // ================= COPIERS =================
            const int cw = warp - 1;
            uint8_t *stage = copier_mem + (size_t)cw * (kDecStage + kDecLitStage + kDecScratch);
            uint8_t *lstage = stage + kDecStage;
            uint8_t *scratch = lstage + kDecLitStage;
            for (int slot = cw; slot < nslots; slot += kDecCopiers) {
                const uint32_t cword = st->count[rb][slot];
                const int n = (int)(cword & 0xff);
                if (n == 0) continue;
                const int b = first_blk + slot;
                const uint8_t *sp = src + sbeg[b];
                uint8_t *dp = dst + dbeg[b];
                const uint32_t *dsc = desc + (rb * kDecSlots + slot) * kDescStride;
                uint32_t litpos = 0, dpos = 0, lit = 0, mlen = 0, off = 1;
                if (lane < n) {
                    litpos = dsc[0 * kDecTok + lane];
                    dpos = dsc[1 * kDecTok + lane];
                    lit = dsc[2 * kDecTok + lane];
                    mlen = dsc[3 * kDecTok + lane];
                    off = dsc[4 * kDecTok + lane];
                }
                if (cword & 0x100u) {
                    // ---- one long token: cooperative copies straight to global memory ----
                    litpos = __shfl_sync(kFullMask, litpos, 0);
                    dpos = __shfl_sync(kFullMask, dpos, 0);
                    lit = __shfl_sync(kFullMask, lit, 0);
                    mlen = __shfl_sync(kFullMask, mlen, 0);
                    off = __shfl_sync(kFullMask, off, 0);
                    if (lit) warp_copy(dp + dpos, sp + litpos, lit, lane);
                    __syncwarp();
                    if (mlen) warp_copy_overlap(dp + dpos + lit, off, mlen, lane);
                    __syncwarp();
                    continue;
                }
                // ---- batch of short tokens ----
                const uint32_t x0 = __shfl_sync(kFullMask, dpos, 0);
                const uint32_t xend = __shfl_sync(kFullMask, dpos + lit + mlen, n - 1);
                const uint32_t shift = (uint32_t)(reinterpret_cast<uintptr_t>(dp + x0) & 15);  // stage mirrors dst alignment
                // 1. stream span -> lstage (aligned 16-byte chunks of the absolute address)
                const uint32_t s_first = st->s_first[rb][slot], s_end = st->s_end[rb][slot];
                const uint8_t *span0 = sp + s_first;
                const uintptr_t a0 = reinterpret_cast<uintptr_t>(span0) & ~uintptr_t(15);
                const uint32_t lshift = (uint32_t)(reinterpret_cast<uintptr_t>(span0) - a0);
                const uint32_t span_chunks = (lshift + (s_end - s_first) + 15) / 16;
                for (uint32_t c = lane; c < span_chunks; c += 32)
                    cp_async16(lstage + 16 * c, reinterpret_cast<const void *>(a0 + 16 * (uintptr_t)c));
                // 2. back-reference gathers of tokens whose source is complete (ends before this span)
                const uint32_t m = dpos + lit;          // output position of the match
                const uint32_t srcpos = m - off;        // its source
                const bool has_m = lane < n && mlen > 0;
                const bool indep = has_m && srcpos + mlen <= x0;
                const bool dep = has_m && !indep;
                uint8_t *myscr = scratch + lane * 48;
                uint32_t g15 = 0;
                if (indep) {
                    const uint8_t *g = dp + srcpos;
                    const uintptr_t ga = reinterpret_cast<uintptr_t>(g) & ~uintptr_t(15);
                    g15 = (uint32_t)(reinterpret_cast<uintptr_t>(g) - ga);
                    cp_async16(myscr, reinterpret_cast<const void *>(ga));
                    if (g15 + min(mlen, 16u) > 16) cp_async16(myscr + 16, reinterpret_cast<const void *>(ga + 16));
                }
                cp_async_wait_all();
                __syncwarp();
                // 3. literals: lstage -> stage
                {
                    uint32_t maxlit = lit;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) maxlit = max(maxlit, __shfl_xor_sync(kFullMask, maxlit, o));
                    const uint8_t *ls = lstage + lshift + (litpos - s_first);
                    uint8_t *ds = stage + shift + (dpos - x0);
                    for (uint32_t i = 0; i < maxlit; i++)
                        if (i < lit) ds[i] = ls[i];
                }
                // 4. independent matches: scratch -> stage, 16 bytes per round
                {
                    uint8_t *ds = stage + shift + (m - x0);
                    uint32_t done = 0;
                    for (;;) {
                        const uint32_t left = indep && mlen > done ? mlen - done : 0;
                        const uint32_t take = min(left, 16u);
                        const uint8_t *ss = myscr + g15;
#pragma unroll
                        for (int i = 0; i < 16; i++)
                            if ((uint32_t)i < take) ds[done + i] = ss[i];
                        done += 16;
                        const bool more = indep && mlen > done;
                        if (!__any_sync(kFullMask, more)) break;
                        __syncwarp();
                        if (more) {  // next 16 source bytes
                            const uint8_t *g = dp + srcpos + done;
                            const uintptr_t ga = reinterpret_cast<uintptr_t>(g) & ~uintptr_t(15);
                            g15 = (uint32_t)(reinterpret_cast<uintptr_t>(g) - ga);
                            cp_async16(myscr, reinterpret_cast<const void *>(ga));
                            if (g15 + min(mlen - done, 16u) > 16) cp_async16(myscr + 16, reinterpret_cast<const void *>(ga + 16));
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
                    for (uint32_t i = lane; i < tlen; i += 32) {
                        const uint32_t j = toff >= tlen ? i : i % toff;  // overlapping: replicate the pattern
                        const uint32_t q = tsrc + j;
                        const uint8_t v = q >= x0 ? stage[shift + (q - x0)] : dp[q];
                        stage[shift + (tm - x0) + i] = v;
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
                        if (lo >= shift && hi <= total) {
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
        // Batch `round` is empty: nothing left to parse, and the copiers drained
        // batch round-1 in this very round.  (produced[] is double buffered, so the
        // parser's next write cannot race with this read.)
        if (produced[wb] == 0) break;
    }

    if (warp == 0 && lane < nslots) status[first_blk + lane] = p_bad ? 1 : 0;
}

}  // namespace mz
