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
constexpr int kRingMirror = 16;  // the first 64 bytes are stored again behind the end: windows never wrap

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
            if (w < kRingMirror) {
                ring[kRingWords + w] = (uint32_t)v;
                ring[kRingWords + w + 1] = (uint32_t)(v >> 32);
            }
            filled += kRingChunk;
            __syncwarp();
        }
    }

    // After a long match the cursor may have left the window: restart behind it.
    __device__ __forceinline__ void seek(int lo) {
        if (lo >= filled) filled = lo & ~(kRingChunk - 1);
    }

    // out[0..7) = src[b .. b+28) (b may be negative at the very start of a
    // block; those bytes are never used).
    __device__ __forceinline__ void fetch28(int b, uint32_t out[7]) const {
        const uint32_t *q = ring + ((b >> 2) & (kRingWords - 1));
        const unsigned sh = (unsigned)(b & 3) * 8;
        uint32_t w[8];
#pragma unroll
        for (int j = 0; j < 8; j++) w[j] = q[j];
#pragma unroll
        for (int j = 0; j < 7; j++) out[j] = __funnelshift_r(w[j], w[j + 1], sh);
    }
};

constexpr int kEncL1Warps = 4;  // warps per CTA
#ifndef MZ_ENC_L1_MIN_CTAS
#define MZ_ENC_L1_MIN_CTAS 8  // 32 warps per SM: 148 x 32 >= 4096 blocks in flight
#endif
constexpr int kEncL1SlotsPerWarp = 1 << 15;
constexpr size_t kEncL1WsBytesPerWarp = (size_t)kEncL1SlotsPerWarp * sizeof(Slot);  // 1 MiB

// 24-bit mask, bit k set when byte k of the two 24-byte strings (6 words each)
// differs.  Per word: "has non-zero byte" flags at bits 7/15/23/31, gathered
// into a nibble by one multiply (the partial products do not collide).
__device__ __forceinline__ uint32_t nzmask24(const uint32_t *x, const uint32_t *y) {
    uint32_t acc = 0;
#pragma unroll
    for (int j = 5; j >= 0; j--) {
        const uint32_t d = x[j] ^ y[j];
        const uint32_t t = (d | ((d & 0x7f7f7f7fu) + 0x7f7f7f7fu)) & 0x80808080u;
        acc = acc * 16u + ((t * 0x204081u) >> 28);
    }
    return acc;
}

// Result of one table probe as the serial algorithm would see it.
struct Probe {
    bool hit;     // candidate in range and equal on the minimum match length
    bool fw;      // candidate is an insert of this same batch (cand/nz/bb below are valid)
    int cand;
    uint32_t nz;  // nzmask24(src[cand..], src[pos..])
    int bb;       // equal bytes going back from cand-1 / pos-1, <= 4
};

// Encodes one block with one warp.  Returns bytes written or 0.
//
// One iteration of the outer loop is one BATCH = one DRAM round trip: lane L
// owns position wbase + L, loads its 32 source bytes from the ring, hashes
// them and fetches its table slot (position + snapshot of the bytes around it).
// After the round trip every lane knows whether ITS position would verify
// against the pre-batch table, how far the match runs (<= 24 bytes forward,
// <= 4 back) and whether it passes the repeat check.  The warp then replays the
// reference's control flow (search steps, back-to-back re-match probes, repeat
// checks) over these 32 answers with warp-uniform bit tests, consuming as many
// steps as fall inside the window, and finally writes back the table inserts
// the replay performed.  Inserts of the same batch that alias a later probe's
// slot are forwarded from the ring (`dup` / Probe::fw).
template <class P>
__device__ int encode_l1_block(uint8_t *dst, const uint8_t *src, const int n, Slot *table, uint32_t *ring_mem,
                               const int lane) {
    const int sLimit = n - kInputMargin;
    const int dstLimit = P::dst_limit(n);
    const int fill_limit = (n + 64 + kRingChunk - 1) & ~(kRingChunk - 1);
    constexpr uint32_t kMinMask = P::kMinMatch == 8 ? 0xffu : 0xfu;

    SrcRing ring{ring_mem, src, n, 0};
    ring.ensure(min(2 * kRingChunk, fill_limit), lane);

    // Empty slots read as candidate 0 (encode_l1.go:52,86): position 0 and its bytes.
    {
        uint32_t w0[7];
        ring.fetch28(-4, w0);
        const uint4 ia = make_uint4(0, 0, w0[1], w0[2]);
        const uint4 ib = make_uint4(w0[3], w0[4], w0[5], w0[6]);
        const int slots = 1 << P::kTableBits;
        for (int i = lane; i < slots; i += 32) slot_store(table + i, ia, ib);
        __syncwarp();
    }

    int nextEmit = 0;
    int s = 1;             // cursor: the re-match position, or the search position t
    int repeat = 1;
    int d = 0;
    bool rematch = false;  // the cursor is in the re-match loop (:222-265)

    // Token emission is deferred by one match: the last match of a batch is
    // emitted while the next batch's loads are in flight.
    int pe_kind = 0;  // 0 none, 1 copy, 2 literals [pe_ne, pe_base) + copy
    int pe_ne = 0, pe_base = 0, pe_repeat = 0, pe_end = 0;
    auto flush_pending = [&]() -> bool {
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

    const unsigned below = (1u << lane) - 1u;
    const unsigned above = ~((2u << lane) - 1u);
    bool done = false;  // reached emitRemainder

    while (!done) {
        // ---------------- one round trip for the whole window ----------------
        const int wbase = rematch ? s - 2 : s;
        // When the skip distance is long (incompressible data) only the first search
        // step can fall in the window: do not fetch slots nobody will consume.
        const int K = (!rematch && ((s - nextEmit) >> P::kSkipLog) >= 24) ? 8 : 32;
        ring.seek(wbase - 8);
        ring.ensure(min(wbase + 320, fill_limit), lane);
        const int p = wbase + lane;
        const bool active = lane < K;

        uint32_t W[7];  // src[p-4 .. p+24)
        ring.fetch28(p - 4, W);
        const uint32_t h = P::hash((uint64_t)W[2] << 32 | W[1]);
        uint4 ea = make_uint4(0, 0, 0, 0), eb = ea;
        if (active) slot_load(table + h, ea, eb);
        const bool rep_lane = active && p >= repeat;
        uint32_t rep4 = 0;
        if (rep_lane) rep4 = ldg_u32_unaligned(src + p - repeat);

        // the loads are in flight: emit the last match of the previous batch
        if (pe_kind && !flush_pending()) return 0;

        const unsigned same = __match_any_sync(kFullMask, h);
        const unsigned dup = __ballot_sync(kFullMask, (same & below) != 0);

        const int cand = (int)ea.x;
        uint32_t nz;
        {
            const uint32_t cd[6] = {ea.z, ea.w, eb.x, eb.y, eb.z, eb.w};
            nz = nzmask24(cd, W + 1);
        }
        const uint32_t xb = ea.y ^ W[0];
        const int bb = xb ? __clz(xb) >> 3 : 4;
        const bool eqm = active && (nz & kMinMask) == 0;
        const int dist = p - cand;
        // search probe j of a step sees minSrcPos = t - maxCopy3Offset (:83): dist <= max + j
        const unsigned E0 = __ballot_sync(kFullMask, eqm && dist <= kMaxCopy3Offset);
        unsigned E1 = E0, E2 = E0;
        if (n > kMaxCopy3Offset) {
            E1 = __ballot_sync(kFullMask, eqm && dist <= kMaxCopy3Offset + 1);
            E2 = __ballot_sync(kFullMask, eqm && dist <= kMaxCopy3Offset + 2);
        }
        const unsigned Brep = __ballot_sync(kFullMask, rep_lane && rep4 == W[1]);

        // ---------------- replay the serial walk over the window ----------------
        unsigned ins = 0;     // lanes whose position was inserted, in serial order = lane order
        bool rep_snap = false;  // repeat checks come from the last match's snapshot (Rnz, Rps)
        uint32_t Rnz = 0;
        int Rps = 0;

        auto probe = [&](int L, unsigned E) -> Probe {
            Probe r;
            r.hit = (E >> L) & 1u;
            r.fw = false;
            r.cand = 0;
            r.nz = 0;
            r.bb = 0;
            if ((dup >> L) & 1u) {  // rare: an earlier lane of the window hashes to the same slot
                const unsigned sm = __shfl_sync(kFullMask, same, L);
                const unsigned f = sm & ins & ((1u << L) - 1u);
                if (f) {  // the latest such insert is what the table holds by now
                    r.fw = true;
                    r.cand = wbase + 31 - __clz(f);
                    uint32_t A[7], B[7];
                    ring.fetch28(wbase + L - 4, A);
                    ring.fetch28(r.cand - 4, B);
                    r.nz = nzmask24(B + 1, A + 1);
                    const uint32_t x = A[0] ^ B[0];
                    r.bb = x ? __clz(x) >> 3 : 4;
                    r.hit = (r.nz & kMinMask) == 0;
                }
            }
            return r;
        };

        for (;;) {
            // termination and window checks first: a batch that ends here keeps its
            // last match pending for the next batch's load shadow
            const int L = s - wbase;
            const int nextS = s + ((s - nextEmit) >> P::kSkipLog) + P::kStep;  // :79 (search mode)
            if (rematch) {
                nextEmit = s;
                if (s >= sLimit) {
                    done = true;
                    break;
                }
                if (L >= K) break;
            } else {
                if (nextS > sLimit) {
                    done = true;
                    break;
                }
                if (L + 2 >= K) break;
                if (rep_snap && s + 1 - Rps > kSnapFwd - 4) break;  // repeat check not covered by the snapshot
            }
            if (pe_kind && !flush_pending()) return 0;
            int mps;  // position of the verified probe
            Probe m;
            bool from_rematch;
            if (rematch) {
                if (d > dstLimit) return 0;  // :229
                m = probe(L, E0);            // read before this step's inserts (:236-239)
                ins |= 5u << (L - 2);
                if (!m.hit) {
                    rematch = false;
                    s++;
                    continue;
                }
                mps = s;
                from_rematch = true;
            } else {
                const int t = s;
                // repeat check at t+1 (:94)
                const bool rhit = rep_snap ? ((Rnz >> (t + 1 - Rps)) & 0xfu) == 0 : (Brep >> (L + 1)) & 1u;
                const Probe r0 = probe(L, E0);
                const Probe r1 = probe(L + 1, E1);
                ins |= 3u << L;
                if (rhit) {
                    int base = t + 1;
                    base -= extend_backward(src, base - repeat, base, nextEmit, lane);
                    if (d + (base - nextEmit) > dstLimit) return 0;
                    d += emit_literal(dst + d, src + nextEmit, base - nextEmit, lane);
                    s = extend_forward8(src, t + 5, t + 5 - repeat, sLimit, lane);
                    d += emit_repeat(dst + d, s - base, lane);
                    nextEmit = s;
                    if (s >= sLimit) {
                        done = true;
                        break;
                    }
                    continue;
                }
                if (r0.hit) {
                    m = r0;
                    mps = t;
                } else {
                    ins |= 4u << L;  // :152 / :157
                    if (r1.hit) {
                        m = r1;
                        mps = t + 1;
                    } else {
                        m = probe(L + 2, E2);  // read after the inserts of t and t+1 (:150)
                        if (!m.hit) {
                            s = nextS;
                            continue;
                        }
                        mps = t + 2;
                    }
                }
                from_rematch = false;
            }

            // ---------------- a verified candidate at mps ----------------
            if (!m.fw) {
                const int L = mps - wbase;
                m.cand = __shfl_sync(kFullMask, cand, L);
                m.nz = __shfl_sync(kFullMask, nz, L);
                m.bb = __shfl_sync(kFullMask, bb, L);
            }
            const int f = m.nz ? __ffs(m.nz) - 1 : kSnapFwd;
            int base = mps, known = f;
            if (!from_rematch && P::kBackExtend) {  // :169-172
                const int room = min(m.cand, mps - nextEmit);
                int back = min(m.bb, room);
                if (back == 4 && room > 4) back += extend_backward(src, m.cand - 4, mps - 4, nextEmit, lane);
                base = mps - back;
                known = back + f;
            }
            repeat = mps - m.cand;
            // Go: s = base + min match, then 8-byte chunks while s <= n-8 (:181-188)
            int q_stop = base + P::kMinMatch;
            if (q_stop <= n - 8) q_stop += (((n - 8 - q_stop) >> 3) + 1) << 3;
            int e;
            if (f < kSnapFwd) {
                e = min(base + known, q_stop);
            } else {
                const int sc = base + P::kMinMatch + 8 * ((known - P::kMinMatch) >> 3);
                e = min(extend_forward8(src, sc, sc - repeat, n - 8, lane), q_stop);
            }
            pe_kind = from_rematch ? 1 : 2;
            pe_ne = nextEmit;
            pe_base = base;
            pe_repeat = repeat;
            pe_end = e;
            rep_snap = true;
            Rnz = m.nz;
            Rps = mps;
            s = e;
            rematch = true;
        }
        if (done) break;

        // ---------------- write back the inserts the replay performed ----------------
        if (((ins >> lane) & 1u) && (same & ins & above) == 0) {  // a later insert on the same slot wins
            slot_store(table + h, make_uint4((uint32_t)p, W[0], W[1], W[2]), make_uint4(W[3], W[4], W[5], W[6]));
        }
        __syncwarp();
    }

    if (pe_kind && !flush_pending()) return 0;
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
__global__ void __launch_bounds__(kEncL1Warps * 32, MZ_ENC_L1_MIN_CTAS)
encode_l1_kernel(int nblk, const uint8_t *__restrict__ src, const uint64_t *__restrict__ sbeg,
                 const uint64_t *__restrict__ send, uint8_t *dst, const uint64_t *__restrict__ dbeg,
                 uint32_t *__restrict__ out_len, int *counter, Slot *tables) {
    __shared__ uint32_t rings[kEncL1Warps][kRingWords + kRingMirror];
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
