// mz_encode_l1.cuh -- MinLZ LevelFastest block encoder (sm_100a).
//
// Replaces encodeBlock (reference asm_none.go:51-59 -> encode_l1.go:39-283
// encodeBlockGo and :285-524 encodeBlockGo64K; asm twin encodeBlockAsm*).
// Byte-identical to the pure-Go path: same hash (hash6/15 bit, hash5/13 bit
// for <= 64 KiB), same probe order, same skip rule, same emit rules.
//
// Why it looks the way it does
// ----------------------------
// The hash walk is one serial dependency chain per block (probe -> verify ->
// extend -> next position), so the kernel is bound by memory LATENCY, not by
// bytes per instruction.  Throughput = chains in flight / round trips per
// step.  Hence:
//   * one block per WARP and every block of the batch in flight at once
//     (28+ warps per SM x 148 SMs >= 4096 blocks); shared-memory tables would
//     cap the chip at ~148 chains.
//   * the match table lives in a global-memory workspace, and every slot is a
//     32-byte (one DRAM sector) record {position, 4 bytes before it, 24 bytes
//     from it}.  A probe then returns the candidate AND the bytes needed to
//     verify, back-extend (<= 4) and forward-extend (<= 24) it in ONE round
//     trip; the reference needs table -> src[candidate] -> extend.  The record
//     is a snapshot of immutable source bytes, so results are unchanged.  This
//     spends HBM capacity (1 MiB per in-flight block) to buy latency.
//   * the lanes of the warp evaluate the next few steps of the walk
//     speculatively in the same round trip: the re-match probe at a match end
//     plus the search steps that follow if it misses.  Results are then
//     resolved in the reference's serial order, with in-flight inserts
//     forwarded between lanes (match.any on the slot index), and only the
//     inserts of steps that really executed are written back.
//   * source bytes near the cursor sit in a 1 KiB per-warp shared-memory ring.
#pragma once

#include "mz_common.cuh"

namespace mz {

// Forward match extension, restating the Go loop
//   for s <= limit { if diff := load64(s)^load64(cand); diff != 0 { s += tz>>3; break }; s += 8; cand += 8 }
// with the lanes comparing consecutive 8-byte words; the first round uses 4
// lanes (most matches end within 32 bytes, and the candidate side is a random
// DRAM sector), later rounds all 32.
__device__ __forceinline__ int extend_forward8(const uint8_t *src, int s, int cand, int limit, int lane) {
    int width = 4;
    for (;;) {
        int pos = s + 8 * lane;
        const bool act = lane < width;
        bool past = pos > limit;
        uint64_t diff = 0;
        if (act && !past) diff = ldg_u64_unaligned(src + pos) ^ ldg_u64_unaligned(src + cand + 8 * lane);
        unsigned stop = __ballot_sync(kFullMask, act && (past || diff != 0));
        if (stop == 0) {
            s += 8 * width;
            cand += 8 * width;
            width = 32;
            continue;
        }
        int f = __ffs(stop) - 1;
        int add = past ? 0 : (__ffsll((long long)diff) - 1) >> 3;
        add = __shfl_sync(kFullMask, add, f);
        return s + 8 * f + add;
    }
}

// Writes the low n bytes of tok at dst (n <= 8), one byte per lane.
__device__ __forceinline__ void put_token(uint8_t *dst, uint64_t tok, int n, int lane) {
    if (lane < n) dst[lane] = (uint8_t)(tok >> (8 * lane));
}

__device__ __forceinline__ void copy_bytes(uint8_t *dst, const uint8_t *src, int n, int lane) {
    for (int i = lane; i < n; i += 32) dst[i] = src[i];
}

// emitLiteral asm_none.go:84-122
__device__ __forceinline__ int emit_literal(uint8_t *dst, const uint8_t *lit, int n, int lane) {
    if (n == 0) return 0;
    int hn;
    uint64_t h = tok_literal_hdr((uint32_t)n, &hn);
    put_token(dst, h, hn, lane);
    copy_bytes(dst + hn, lit, n, lane);
    return hn + n;
}

__device__ __forceinline__ int emit_repeat(uint8_t *dst, int length, int lane) {
    int n;
    uint64_t t = tok_repeat((uint32_t)length, &n);
    put_token(dst, t, n, lane);
    return n;
}

__device__ __forceinline__ int emit_copy(uint8_t *dst, int offset, int length, int lane) {
    int n;
    uint64_t t = tok_copy((uint32_t)offset, (uint32_t)length, &n);
    put_token(dst, t, n, lane);
    return n;
}

// emitCopyLits2 asm_none.go:284-308
__device__ __forceinline__ int emit_copy_lits2(uint8_t *dst, const uint8_t *lits, int nlits, int offset, int length,
                                               int lane) {
    uint64_t rep;
    int rn;
    uint32_t h = tok_copy2_fused((uint32_t)offset, (uint32_t)length, (uint32_t)nlits, &rep, &rn);
    put_token(dst, h, 3, lane);
    if (lane < nlits) dst[3 + lane] = lits[lane];
    put_token(dst + 3 + nlits, rep, rn, lane);
    return 3 + nlits + rn;
}

// emitCopyLits3 asm_none.go:313-323
__device__ __forceinline__ int emit_copy_lits3(uint8_t *dst, const uint8_t *lits, int nlits, int offset, int length,
                                               int lane) {
    int n;
    uint64_t t = tok_copy3((uint32_t)offset, (uint32_t)length, (uint32_t)nlits, &n);
    put_token(dst, t, n, lane);
    if (lane < nlits) dst[n + lane] = lits[lane];
    return n + nlits;
}

// LevelFastest: encode_l1.go:39-283 (large) and :285-524 (<= 64 KiB)
template <bool kSmall>
struct L1Params {
    static constexpr int kTableBits = kSmall ? 13 : 15;
    static constexpr int kSkipLog = kSmall ? 5 : 6;
    static constexpr int kStep = 4;
    static constexpr int kMaxFuseLits = kSmall ? kMaxCopy2Lits : kMaxCopy3Lits;
    static constexpr int kMinMatch = 4;         // candidates are verified on 4 bytes
    static constexpr bool kBackExtend = true;   // :169-172
    __device__ static __forceinline__ int dst_limit(int n) { return n - (n >> 5) - 6; }
    __device__ static __forceinline__ uint32_t hash(uint64_t u) {
        return kSmall ? hash5(u, kTableBits) : hash6(u, kTableBits);
    }
};

// LevelSuperFast: encode_l0.go:32-279 (large) and :281-522 (<= 64 KiB).  Same walk
// with an 8-byte minimum match, hash8, no backward extension, extension from +8.
template <bool kSmall>
struct L0Params {
    static constexpr int kTableBits = kSmall ? 12 : 13;
    static constexpr int kSkipLog = kSmall ? 4 : 5;
    static constexpr int kStep = kSmall ? 4 : 5;
    static constexpr int kMaxFuseLits = kSmall ? kMaxCopy2Lits : kMaxCopy3Lits;
    static constexpr int kMinMatch = 8;
    static constexpr bool kBackExtend = false;  // encode_l0.go:164 `for false && ...`
    __device__ static __forceinline__ int dst_limit(int n) { return kSmall ? n - (n >> 4) - 32 : n - (n >> 3) - 6; }
    __device__ static __forceinline__ uint32_t hash(uint64_t u) { return hash8(u, kTableBits); }
};

// Backward extension, restating
//   for cand > 0 && s > floor && src[cand-1] == src[s-1] { cand--; s-- }
// with one byte pair per lane per round.  Returns the number of steps taken.
__device__ __forceinline__ int extend_backward(const uint8_t *src, int cand, int s, int floor_, int lane) {
    int total = 0;
    for (;;) {
        int room = min(cand, s - floor_);
        bool ok = lane < room && src[cand - 1 - lane] == src[s - 1 - lane];
        unsigned m = __ballot_sync(kFullMask, ok);
        int cnt = m == kFullMask ? 32 : __ffs(~m) - 1;
        total += cnt;
        if (cnt < 32) return total;
        cand -= 32;
        s -= 32;
    }
}

// ---- table slot record (one 32-byte DRAM sector) ----------------------------
// a.x = position, a.y = src[pos-4 .. pos), a.z .. b.w = src[pos .. pos+24)
struct __align__(32) Slot {
    uint4 a, b;
};
constexpr int kSnapFwd = 24;  // forward bytes held in a slot

// One slot = one 32-byte DRAM sector, moved with Blackwell's 256-bit global
// accesses (LDG.E.ENL2.256 / STG.E.ENL2.256): a single request per probe, and a
// full-sector write per insert (no read-for-ownership of a half-written sector).
__device__ __forceinline__ void slot_load(const Slot *p, uint4 &a, uint4 &b) {
    asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
                 : "l"(p)
                 : "memory");
}
__device__ __forceinline__ void slot_store(Slot *p, const uint4 &a, const uint4 &b) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
                 "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
                 : "memory");
}

// ---- per-warp ring of source bytes around the cursor ------------------------
constexpr int kRingBytes = 1024;
constexpr int kRingWords = kRingBytes / 4;
constexpr int kRingChunk = 256;

struct SrcRing {
    uint32_t *ring;  // kRingWords words of shared memory owned by this warp
    const uint8_t *src;
    int n;
    int filled;  // chunks up to here are loaded (positions >= n read as 0)

    // Loads 256-byte chunks until `want_end` is covered.  All lanes call.
    __device__ __forceinline__ void ensure(int want_end, int lane) {
        while (filled < want_end) {
            int pos = filled + 8 * lane;
            uint64_t v = 0;
            if (pos + 8 <= n) {
                v = ldg_u64_unaligned(src + pos);
            } else {
                for (int i = 0; i < 8; i++)
                    if (pos + i < n) v |= (uint64_t)src[pos + i] << (8 * i);
            }
            __syncwarp();
            int w = (pos >> 2) & (kRingWords - 1);
            ring[w] = (uint32_t)v;
            ring[w + 1] = (uint32_t)(v >> 32);
            filled += kRingChunk;
            __syncwarp();
        }
    }

    // After a long match the cursor may have left the window: restart behind it.
    __device__ __forceinline__ void seek(int lo) {
        if (lo >= filled) filled = lo & ~(kRingChunk - 1);
    }

    // out[0..8) = src[b .. b+32) (b may be negative at the very start of a
    // block; those bytes are never used).
    __device__ __forceinline__ void fetch32(int b, uint32_t out[8]) const {
        int a = b >> 2;
        unsigned sh = (unsigned)(b & 3) * 8;
        uint32_t w[9];
#pragma unroll
        for (int j = 0; j < 9; j++) w[j] = ring[(a + j) & (kRingWords - 1)];
#pragma unroll
        for (int j = 0; j < 8; j++) out[j] = __funnelshift_r(w[j], w[j + 1], sh);
    }
};

// bytes of common prefix of two 24-byte strings held as 6 words each
__device__ __forceinline__ int prefix24(const uint32_t *x, const uint32_t *y) {
    int r = 24;
#pragma unroll
    for (int j = 5; j >= 0; j--) {
        uint32_t d = x[j] ^ y[j];
        if (d) r = 4 * j + ((__ffs(d) - 1) >> 3);
    }
    return r;
}

constexpr int kEncL1Warps = 4;  // warps per CTA
constexpr int kEncL1SlotsPerWarp = 1 << 15;
constexpr size_t kEncL1WsBytesPerWarp = (size_t)kEncL1SlotsPerWarp * sizeof(Slot);  // 1 MiB

// Speculative search steps evaluated per round trip.
constexpr int kSpecAfterRematch = 1;  // behind a re-match probe
constexpr int kSpecSearch = 2;        // in a pure search batch
constexpr int kMaxLevels = 2;

__device__ __forceinline__ int pick(const int (&t)[kMaxLevels + 1], int i) {
    int r = t[0];
#pragma unroll
    for (int j = 1; j <= kMaxLevels; j++) r = i == j ? t[j] : r;
    return r;
}

// Encodes one block with one warp.  Returns bytes written or 0.
template <class P>
__device__ int encode_l1_block(uint8_t *dst, const uint8_t *src, const int n, Slot *table, uint32_t *ring_mem,
                               const int lane) {
    const int sLimit = n - kInputMargin;
    const int dstLimit = P::dst_limit(n);
    const int fill_limit = (n + 64 + kRingChunk - 1) & ~(kRingChunk - 1);

    SrcRing ring{ring_mem, src, n, 0};
    ring.ensure(min(2 * kRingChunk, fill_limit), lane);

    // Empty slots read as candidate 0 (encode_l1.go:52,86): position 0 and its bytes.
    {
        uint32_t w0[8];
        ring.fetch32(-4, w0);
        const uint4 ia = make_uint4(0, 0, w0[1], w0[2]);
        const uint4 ib = make_uint4(w0[3], w0[4], w0[5], w0[6]);
        const int slots = 1 << P::kTableBits;
        for (int i = lane; i < slots; i += 32) slot_store(table + i, ia, ib);
        __syncwarp();
    }

    int nextEmit = 0;
    int s = 1;
    int repeat = 1;
    int d = 0;
    bool rematch = false;  // the next batch starts with the re-match probe at s (:222-265)

    // Token emission is deferred by one batch: the probes of the next batch only
    // need the match end, so their loads are issued first and the token of the
    // previous match is built and stored while they are in flight.
    int pe_kind = 0;  // 0 none, 1 copy, 2 literals [pe_ne, pe_base) + copy
    int pe_ne = 0, pe_base = 0, pe_repeat = 0, pe_end = 0;
    auto flush_pending = [&]() -> bool {
        if (pe_kind == 0) return true;
        const int length = pe_end - pe_base;
        if (pe_kind == 2 && pe_ne != pe_base) {  // :190-206
            if (pe_base - pe_ne > P::kMaxFuseLits || pe_repeat < kMinCopy2Offset) {
                if (d + (pe_end - pe_ne) > dstLimit) return false;
                d += emit_literal(dst + d, src + pe_ne, pe_base - pe_ne, lane);
                d += emit_copy(dst + d, pe_repeat, length, lane);
            } else if (pe_repeat <= kMaxCopy2Offset) {
                d += emit_copy_lits2(dst + d, src + pe_ne, pe_base - pe_ne, pe_repeat, length, lane);
            } else {
                d += emit_copy_lits3(dst + d, src + pe_ne, pe_base - pe_ne, pe_repeat, length, lane);
            }
        } else {
            d += emit_copy(dst + d, pe_repeat, length, lane);
        }
        pe_kind = 0;
        return true;
    };

    // loop-invariant lane roles
    // lane 0: insert-only (s-2)      lane 1: re-match probe (s)
    // lanes 2+4j .. 5+4j: level j -> hash0(t), hash1(t+1), hash2(t+2), repeat probe at t+1
    const bool lvl_lane = lane >= 2 && lane < 2 + 4 * kMaxLevels;
    const int my_lvl = (lane - 2) >> 2;
    const int sub = lvl_lane ? (lane - 2) & 3 : 0;
    const int padd = sub == 3 ? 1 : sub;
    unsigned vis_base = 0;
    if (lvl_lane && sub < 3) {
        vis_base = (1u << (2 + 4 * my_lvl)) - 1;           // the re-match lanes and every earlier level
        if (sub == 2) vis_base |= 3u << (2 + 4 * my_lvl);  // hash2 is read after hash0/hash1 were written
    }
    const unsigned later = ~((2u << lane) - 1u);

    for (;;) {
        // ---------------- plan the batch ----------------
        if (rematch) {
            nextEmit = s;
            if (s >= sLimit) break;
        }
        const int ne = nextEmit;
        ring.seek(s - 8);
        ring.ensure(min(s + 320, fill_limit), lane);
        // level j probes t[j], t[j]+1, t[j]+2; t[j+1] is its nextS (:79)
        int t[kMaxLevels + 1];
        t[0] = rematch ? s + 1 : s;
        int nlev = 0;
        bool hit_end = false;  // the first level not probed starts with nextS > sLimit (:80)
        {
            const int want = rematch ? kSpecAfterRematch : kSpecSearch;
            bool open = true;
#pragma unroll
            for (int j = 0; j < kMaxLevels; j++) {
                int nxt = t[j] + ((t[j] - ne) >> P::kSkipLog) + P::kStep;
                t[j + 1] = nxt;
                if (open && j < want) {
                    if (nxt > sLimit) {
                        hit_end = true;
                        open = false;
                    } else if (t[j] + 2 + 28 <= ring.filled) {
                        nlev = j + 1;
                    } else {
                        open = false;
                    }
                } else {
                    open = false;
                }
            }
        }
        if (!rematch && nlev == 0) {
            if (hit_end) break;
            return 0;  // unreachable: ensure() always covers one level
        }

        // ---------------- this batch's lane roles ----------------
        int lvl = -1, p = s;
        if (lane < 2) {
            lvl = rematch ? -2 : -1;
            p = lane == 0 ? s - 2 : s;
        } else if (lvl_lane && my_lvl < nlev) {
            lvl = my_lvl;
            p = pick(t, my_lvl) + padd;
        }
        const bool active = lvl != -1;
        const bool isrep = lvl >= 0 && sub == 3;
        const bool inserts = active && !isrep;
        const bool reads = inserts && lane != 0;

        uint32_t W[8];  // src[p-4 .. p+28)
        ring.fetch32(p - 4, W);
        const uint32_t h = P::hash((uint64_t)W[2] << 32 | W[1]);

        // one 32-byte sector per probing lane; the repeat probe reads the source
        uint4 ea = make_uint4(0, 0, 0, 0), eb = ea;
        if (reads) slot_load(table + h, ea, eb);
        uint32_t rep4 = 0;
        if (isrep) rep4 = ldg_u32_unaligned(src + p - repeat);

        // the loads are in flight: emit the previous match now
        if (!flush_pending()) return 0;
        if (rematch && d > dstLimit) return 0;  // :229

        // in-flight forwarding: the latest earlier insert (serial order = lane order) on my slot
        const unsigned ins_mask = __ballot_sync(kFullMask, inserts);
        const unsigned same = __match_any_sync(kFullMask, inserts ? h : 0x80000000u + lane) & ins_mask;
        const unsigned vis = reads ? vis_base & same : 0;
        int cand = (int)ea.x;
        uint32_t cb = ea.y;
        uint32_t cd[6] = {ea.z, ea.w, eb.x, eb.y, eb.z, eb.w};
        const bool fwd = vis != 0;
        if (__any_sync(kFullMask, fwd)) {  // rare: short-period data
            const int from = fwd ? 31 - __clz(vis) : lane;
            const int cpos = __shfl_sync(kFullMask, p, from);
            if (fwd) {
                cand = cpos;
                uint32_t V[8];
                ring.fetch32(cand - 4, V);
                cb = V[0];
#pragma unroll
                for (int j = 0; j < 6; j++) cd[j] = V[j + 1];
            }
        }

        // ---------------- evaluate ----------------
        bool ok = false;
        int fbytes = 0, bbytes = 0;
        if (reads) {
            const int anchor = lvl >= 0 ? pick(t, lvl) : s;
            const bool in_range = lvl >= 0 ? cand >= anchor - kMaxCopy3Offset : s - cand <= kMaxCopy3Offset;
            if (in_range && cd[0] == W[1] && (P::kMinMatch == 4 || cd[1] == W[2])) {
                ok = true;
                fbytes = prefix24(cd, W + 1);
                const uint32_t x = cb ^ W[0];
                bbytes = x ? __clz(x) >> 3 : 4;
            }
        } else if (isrep) {
            ok = rep4 == W[1];
        }
        const unsigned okm = __ballot_sync(kFullMask, ok);

        // serial priority: re-match probe; then per level repeat, hash0, hash1, hash2
        int win = -1;
        int win_lvl = nlev;  // nlev: every probed level missed
        if (rematch && (okm & 2u)) {
            win = 1;
            win_lvl = -2;
        } else {
#pragma unroll
            for (int j = kMaxLevels - 1; j >= 0; j--) {
                const unsigned g = (okm >> (2 + 4 * j)) & 15u;
                if (j < nlev && g) {
                    win = 2 + 4 * j + ((g & 8u) ? 3 : (g & 1u) ? 0 : (g & 2u) ? 1 : 2);
                    win_lvl = j;
                }
            }
        }
        const int wsub = win >= 2 ? (win - 2) & 3 : 0;

        // ---------------- write back the inserts that really happened ----------------
        {
            bool w = false;
            if (inserts) {
                if (lvl == -2) w = true;                       // :238-239 run before the test
                else if (win_lvl == -2) w = false;
                else if (lvl < win_lvl) w = true;
                else if (lvl == win_lvl) w = sub < 2 || wsub == 1 || wsub == 2;  // hash2: :152/:157 only
            }
            const unsigned wm = __ballot_sync(kFullMask, w);
            if (w && (same & wm & later) == 0) {  // a later insert on the same slot wins
                slot_store(table + h, make_uint4((uint32_t)p, W[0], W[1], W[2]), make_uint4(W[3], W[4], W[5], W[6]));
            }
            __syncwarp();
        }

        if (win < 0) {  // every probed step missed: continue the search behind them
            s = pick(t, nlev);
            rematch = false;
            if (hit_end) break;
            continue;
        }

        // ---------------- process the hit ----------------
        const int wcand = __shfl_sync(kFullMask, cand, win);
        const int wf = __shfl_sync(kFullMask, fbytes, win);
        const int wb = __shfl_sync(kFullMask, bbytes, win);

        if (win_lvl >= 0 && wsub == 3) {  // repeat at t+1 (encode_l1.go:94-145)
            const int tt = pick(t, win_lvl);
            int base = tt + 1;
            base -= extend_backward(src, base - repeat, base, ne, lane);
            if (d + (base - ne) > dstLimit) return 0;
            d += emit_literal(dst + d, src + ne, base - ne, lane);
            s = extend_forward8(src, tt + 5, tt + 5 - repeat, sLimit, lane);
            d += emit_repeat(dst + d, s - base, lane);
            nextEmit = s;
            if (s >= sLimit) break;
            rematch = false;
            continue;
        }

        int base, known;  // match start, bytes known equal from it
        if (win_lvl == -2) {
            base = s;
            repeat = s - wcand;
            known = wf;
        } else {
            const int ps = pick(t, win_lvl) + wsub;
            const int room = P::kBackExtend ? min(wcand, ps - ne) : 0;  // :169-172
            int back = min(wb, room);
            if (back == 4 && room > 4) back += extend_backward(src, wcand - 4, ps - 4, ne, lane);
            base = ps - back;
            repeat = ps - wcand;
            known = back + wf;
        }
        {
            // Go: s = base + min match, then 8-byte chunks while s <= n-8 (:181-188)
            int q_stop = base + P::kMinMatch;
            if (q_stop <= n - 8) q_stop += (((n - 8 - q_stop) >> 3) + 1) << 3;
            if (wf < kSnapFwd) {
                s = min(base + known, q_stop);
            } else {
                const int sc = base + P::kMinMatch + 8 * ((known - P::kMinMatch) >> 3);
                s = min(extend_forward8(src, sc, sc - repeat, n - 8, lane), q_stop);
            }
        }
        pe_kind = win_lvl == -2 ? 1 : 2;
        pe_ne = ne;
        pe_base = base;
        pe_repeat = repeat;
        pe_end = s;
        rematch = true;
    }

    if (!flush_pending()) return 0;
    // emitRemainder (encode_l1.go:268-282)
    if (nextEmit < n) {
        if (d + n - nextEmit > dstLimit) return 0;
        d += emit_literal(dst + d, src + nextEmit, n - nextEmit, lane);
    }
    return d;
}

// Persistent kernel: every warp pulls block indices from *counter and owns the
// workspace slice `tables + global_warp * 1 MiB`.
template <bool kSuperFast>
__global__ void __launch_bounds__(kEncL1Warps * 32)
encode_l1_kernel(int nblk, const uint8_t *__restrict__ src, const uint64_t *__restrict__ sbeg,
                 const uint64_t *__restrict__ send, uint8_t *dst, const uint64_t *__restrict__ dbeg,
                 uint32_t *__restrict__ out_len, int *counter, Slot *tables) {
    __shared__ uint32_t rings[kEncL1Warps][kRingWords];
    const int lane = lane_id();
    const int warp = threadIdx.x >> 5;
    const int gwarp = blockIdx.x * kEncL1Warps + warp;
    Slot *table = tables + (size_t)gwarp * kEncL1SlotsPerWarp;
    for (;;) {
        int blk = 0;
        if (lane == 0) blk = atomicAdd(counter, 1);
        blk = __shfl_sync(kFullMask, blk, 0);
        if (blk >= nblk) return;
        const uint8_t *sp = src + sbeg[blk];
        const int64_t n64 = (int64_t)(send[blk] - sbeg[blk]);
        uint8_t *dp = dst + dbeg[blk];
        int res = 0;
        if (n64 >= kMinNonLiteralBlockSize && n64 <= kMaxBlockSize) {
            const int n = (int)n64;
            if (kSuperFast)
                res = n <= 65536 ? encode_l1_block<L0Params<true>>(dp, sp, n, table, rings[warp], lane)
                                 : encode_l1_block<L0Params<false>>(dp, sp, n, table, rings[warp], lane);
            else
                res = n <= 65536 ? encode_l1_block<L1Params<true>>(dp, sp, n, table, rings[warp], lane)
                                 : encode_l1_block<L1Params<false>>(dp, sp, n, table, rings[warp], lane);
        }
        if (lane == 0) out_len[blk] = (uint32_t)res;
        __syncwarp();
    }
}

}  // namespace mz
