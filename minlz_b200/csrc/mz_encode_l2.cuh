// mz_encode_l2.cuh -- MinLZ LevelBalanced block encoder (sm_100a).
//
// Replaces encodeBlockBetter (reference asm_none.go:68-76 ->
// encode_l2.go:61-338 encodeBlockBetterGo and :343-596 ...Go64K; asm twin
// encodeBetterBlockAsm*).  Byte-identical to the pure-Go path: long table
// 17 bit / hash7 and short table 14 bit / hash4 (15 bit / hash6 and 12 bit for
// blocks <= 64 KiB), same probe order (long-8, repeat, long-4, short + long at
// s+1), same skip rule, same re-indexing of the match interior.
//
// Mapping: one block per warp, all blocks of the batch in flight at once; the
// tables do not fit shared memory (and shared memory would cap the chip at ~148
// chains of a latency-bound walk), so they live in a global-memory workspace slice
// per warp.  The walk (encode_l2_walk, one template for both flavours) probes a
// window of 32 consecutive positions per DRAM round trip and steps through it on
// the cached entries, which every insert of the walk patches; match extension and
// the re-indexing of a match's interior are lane parallel.  The step-at-a-time
// walks of round 1 (encode_l2_block / encode_l2_asm_block: lanes 0-3 issue the
// probes of ONE step per round trip) are kept behind -DMZ_L2_REPLAY=0 for A/B.
//
// Table entries are SNAPSHOTS, as in the L1 kernel: a long-table entry is
// {position, the 8 source bytes at it} (16 B), a short-table entry {position, 4
// bytes} (8 B).  The reference's table -> src[candidate] chain (two dependent DRAM
// round trips per step) becomes one; the bytes are copies of immutable source
// bytes, so every decision is unchanged.  An untouched (zero) entry stands for
// candidate 0, whose bytes are src[0..8) (encode_l2.go:84,121-130).  2.1 MiB of
// workspace per in-flight block.  (Wider entries -- 12 bytes + the byte before the
// position -- were measured and lost: profiles/r02_l2_wide_snapshot_negative.txt.)
#pragma once

#include "mz_common.cuh"
#include "mz_encode_l1.cuh"

namespace mz {

template <bool kSmall>
struct L2Params {
    static constexpr int kLBits = kSmall ? 15 : 17;
    static constexpr int kSBits = kSmall ? 12 : 14;
    __device__ static __forceinline__ uint32_t hashL(uint64_t u) { return kSmall ? hash6(u, kLBits) : hash7(u, kLBits); }
    __device__ static __forceinline__ uint32_t hashS(uint64_t u) { return hash4(u, kSBits); }
};

// One warp per CTA, as in the LevelFastest kernel (mz_encode_l1.cuh): constant shared-memory addresses;
// 319 -> 298 ms (amd64 flavour) / 301 ms (Go) on 4096 x 1 MiB (profiles/r02_one_warp_per_cta_l1_l2.txt).
#ifndef MZ_ENC_L2_WARPS
#define MZ_ENC_L2_WARPS 1
#endif
constexpr int kEncL2Warps = MZ_ENC_L2_WARPS;  // warps per CTA
#ifndef MZ_ENC_L2_MIN_CTAS
#define MZ_ENC_L2_MIN_CTAS (28 / MZ_ENC_L2_WARPS)  // 28 blocks in flight per SM, <= 72 registers
#endif
constexpr size_t kEncL2WsBytesPerWarp = ((size_t)(1 << 17) * 16) + ((size_t)(1 << 14) * 8);  // 2 MiB + 128 KiB

// long-table entry {pos, 8 bytes at pos}; short-table entry {pos, 4 bytes at pos}
__device__ __forceinline__ void l2_put_long(uint4 *t, uint32_t h, int pos, uint64_t bytes) {
    t[h] = make_uint4((uint32_t)pos, (uint32_t)bytes, (uint32_t)(bytes >> 32), 0u);
}
// Optional 1-bit tag per short-table slot in shared memory (2 KiB per block), as in the L1 kernel:
// a window probe whose tag differs from the tag of its own 4 bytes cannot verify and is not fetched.
// (The long table would need 16 KiB per block and does not fit.)
#ifndef MZ_L2_STAGS
#define MZ_L2_STAGS 0
#endif
#if MZ_L2_STAGS
__device__ __forceinline__ uint32_t *l2_stags() {
    __shared__ uint32_t tags[kEncL2Warps][(1 << 14) / 32];
    return tags[kEncL2Warps == 1 ? 0 : threadIdx.x >> 5];
}
__device__ __forceinline__ uint32_t l2_tag1(uint32_t v) { return (v * 2654435761u) >> 31; }
#endif
__device__ __forceinline__ void l2_put_short(uint2 *t, uint32_t h, int pos, uint64_t bytes) {
    t[h] = make_uint2((uint32_t)pos, (uint32_t)bytes);
#if MZ_L2_STAGS
    uint32_t *tg = l2_stags();
    const uint32_t bit = 1u << (h & 31);
    if (((tg[h >> 5] & bit) != 0) != (l2_tag1((uint32_t)bytes) != 0)) atomicXor(&tg[h >> 5], bit);
#endif
}

// 8 bytes at pos, zero filled past n; never touches a word with no byte < n.
__device__ __forceinline__ uint64_t ld64_clamped(const uint8_t *src, int pos, int n) {
    if (pos + 8 <= n) return ldg_u64_unaligned(src + pos);
    uint64_t v = 0;
    for (int i = 0; i < 8 && pos + i < n; i++) v |= (uint64_t)src[pos + i] << (8 * i);
    return v;
}

// Forward extension with byte tail (encode_l2.go:160-175, :239-254): the
// common prefix of src[s..n) and src[cand..], 8 bytes per lane per round.
__device__ __forceinline__ int extend_to_end(const uint8_t *src, int n, int s, int cand, int lane) {
    int width = 4;
    for (;;) {
        const int pos = s + 8 * lane;
        const bool act = lane < width;
        uint64_t diff = 0;
        bool stop = false;
        if (act) {
            if (pos >= n) {
                stop = true;  // nothing left: ends exactly here
            } else {
                diff = ld64_clamped(src, pos, n) ^ ld64_clamped(src, cand + 8 * lane, n);
                const int left = n - pos;
                if (left < 8) diff |= 1ull << (8 * left);  // the block ends inside this word
                stop = diff != 0;
            }
        }
        const unsigned m = __ballot_sync(kFullMask, stop);
        if (m == 0) {
            s += 8 * width;
            cand += 8 * width;
            width = 32;
            continue;
        }
        const int f = __ffs(m) - 1;
        int add = diff ? (__ffsll((long long)diff) - 1) >> 3 : 0;
        add = __shfl_sync(kFullMask, add, f);
        return s + 8 * f + add;
    }
}

// ---- table inserts of a match, one per lane ------------------------------------
// After every match the reference re-indexes its interior: the two ends in both tables and every
// other position of the middle in the long table (encode_l2.go:183-195,303-326; gen.go:1921-1978)
// -- 7.7 inserts per search step on the benchmark's blocks, and in a walk whose speed is the
// latency of ONE warp's instruction chain that was half of all instructions.  Here lane r performs
// the r-th insert of the reference's serial order (32 per pass): position, 8 source bytes, hash,
// store.  Nothing reads the tables in between, so only the final state matters: when two inserts
// of a pass hit the same slot of the same table the later one in serial order -- the higher lane --
// is the one that stores ("later write wins"), found with one match.any on (table, slot).
enum { kL2OrderGo = 0, kL2OrderAsm = 1, kL2OrderRepeat = 2 };
template <int kOrder, class H>
__device__ __forceinline__ void l2_index_match(const H &P, uint4 *lTable, uint2 *sTable, const uint8_t *src, const int base,
                                               const int e, const int lane) {
    int total;
    int mid = 0;
    if (kOrder == kL2OrderRepeat) {
        // while (index0 < index1) { L[index0] S[index0+1] L[index1] S[index1+1]; index0 += 2; index1 -= 2 }
        const int span = e - base - 3;  // index1 - index0 at the start
        total = span > 0 ? 4 * ((span + 3) >> 2) : 0;
    } else {
        // L[base+1] S[base+2] L[e-2] S[e-1] (Asm: L L S S), then index0 = base+2, index1 = e-3,
        // index2 = (index0 + index1 + 1) >> 1; while (index2 < index1) { L[index0] L[index2]; += 2 }
        mid = (base + 2 + e - 3 + 1) >> 1;
        const int left = e - 3 - mid;
        total = 4 + (left > 0 ? 2 * ((left + 1) >> 1) : 0);
    }
    for (int r0 = 0; r0 < total; r0 += 32) {
        const int r = r0 + lane;
        const bool act = r < total;
        int pos = 0;
        bool is_short = false;
        if (kOrder == kL2OrderRepeat) {
            const int j = r >> 2, t = r & 3;
            pos = (t & 2) ? e - 2 - 2 * j : base + 1 + 2 * j;
            is_short = t & 1;
            pos += is_short ? 1 : 0;
        } else if (r < 4) {
            const bool second = kOrder == kL2OrderAsm ? (r & 1) : (r & 2);  // which end
            is_short = kOrder == kL2OrderAsm ? (r & 2) : (r & 1);
            pos = (second ? e - 2 : base + 1) + (is_short ? 1 : 0);
        } else {
            const int k = (r - 4) >> 1;
            pos = ((r & 1) ? mid : base + 2) + 2 * k;
        }
        uint64_t cv = 0;
        uint32_t key = 0x20000000u + (uint32_t)lane;  // idle lanes: keys of their own
        if (act) {
            cv = ldg_u64_unaligned(src + pos);  // pos + 8 <= n: the caller indexes only behind e < sLimit
            key = is_short ? (0x80000000u | P.hashS(cv)) : P.hashL(cv);
        }
        const unsigned grp = __match_any_sync(kFullMask, key);
        if (act && (grp >> lane) == 1u) {  // no later insert of this pass on my slot
            if (is_short) l2_put_short(sTable, key & 0x7fffffffu, pos, cv);
            else l2_put_long(lTable, key, pos, cv);
        }
        __syncwarp();
    }
}

// ---- token records ------------------------------------------------------------
// As in the L1 kernel the walk only RECORDS its matches (base, end, offset | kind << 24) in a
// per-warp ring in shared memory; 32 at a time they become tokens in emit_group (lane = record,
// prefix-sum positions, the reference's bail-out tests per lane with the serial values of d).
// A match the walk drops (far 4-byte match) leaves no record: its dst test would have failed
// only if the next record's or the remainder's test fails too (d unchanged, literals only grow).
struct L2EmitGo {
    static constexpr bool kAsm = false, kBalanced = true;
    __device__ __forceinline__ int max_fuse_lits2() const { return kMaxCopy2Lits; }  // encode_l2.go:268
    __device__ __forceinline__ int max_fuse_lits3() const { return kMaxCopy3Lits; }  // :279
    __device__ __forceinline__ int lit_overhead() const { return 0; }
    __device__ __forceinline__ bool lit_quirk() const { return false; }
};
struct L2EmitAsm {
    static constexpr bool kAsm = true, kBalanced = true;
    int ovh;
    bool quirk;
    __device__ __forceinline__ int max_fuse_lits2() const { return 4; }  // gen.go:1822
    __device__ __forceinline__ int max_fuse_lits3() const { return 3; }  // gen.go:1850
    __device__ __forceinline__ int lit_overhead() const { return ovh; }
    __device__ __forceinline__ bool lit_quirk() const { return quirk; }
};
struct L2Records {
    uint32_t *recs;  // [3][kRecRing] in shared memory
    int head, pending;
    __device__ __forceinline__ void push(int base, int rk, int end, int lane) {
        if (lane == 0) {
            const int w = (head + pending) & (kRecRing - 1);
            recs[w] = (uint32_t)base;
            recs[kRecRing + w] = (uint32_t)rk;
            recs[2 * kRecRing + w] = (uint32_t)end;
        }
        pending++;
        __syncwarp();
    }
    // Emits full groups of 32 (everything when `all`).  Returns false when a bail-out test fires.
    template <class E>
    __device__ __forceinline__ bool flush(const E prm, uint8_t *dst, const uint8_t *src, int &d, int &emitted, int lane,
                                          int sLimit, int dstLimit, bool all) {
        while (pending >= 32 || (all && pending > 0)) {
            const int cnt = min(pending, 32);
            const int r = (head + lane) & (kRecRing - 1);
            const int g_base = (int)recs[r], g_rk = (int)recs[kRecRing + r], g_end = (int)recs[2 * kRecRing + r];
            if (!emit_group(prm, dst, src, d, emitted, g_base, g_rk, g_end, cnt, lane, sLimit, dstLimit)) return false;
            head = (head + cnt) & (kRecRing - 1);
            pending -= cnt;
            __syncwarp();
        }
        return true;
    }
};

template <bool kSmall>
__device__ int encode_l2_block(uint8_t *dst, const uint8_t *src, const int n, uint4 *lTable, uint2 *sTable,
                               uint32_t *rec_mem, const int lane) {
    using P = L2Params<kSmall>;
    L2Records q{rec_mem, 0, 0};
    const L2EmitGo ep{};
    int emitted = 0;  // nextEmit as the token writer sees it
    const uint64_t src0 = ldg_u64_unaligned(src);  // the bytes an untouched entry (candidate 0) stands for
    const int sLimit = n - kInputMargin;
    const int dstLimit = n - (n >> 5) - 6;
    int nextEmit = 0;
    int s = 1;
    uint64_t cv = ldg_u64_unaligned(src + s);
    int repeat = 1;
    int d = 0;

    for (;;) {
        int candidateL = 0;
        int nextS = 0;
        bool continue_outer = false;
        for (;;) {
            nextS = s + ((s - nextEmit) >> 7) + 1;  // :114
            if (nextS > sLimit) goto emit_remainder;
            const int minSrcPos = s - kMaxCopy3Offset + 1;  // :118
            // lane 0: long(cv)  lane 1: short(cv)  lane 2: repeat  lane 3: long(cv>>8) (used only behind a short hit)
            uint32_t h = 0;
            int c = 0;
            if (lane == 0) h = P::hashL(cv);
            if (lane == 1) h = P::hashS(cv);
            if (lane == 3) h = P::hashL(cv >> 8);
            const uint32_t hL = __shfl_sync(kFullMask, h, 0);
            uint64_t v = 0;  // the candidate's bytes: valLong / valShort (:127-128), load32 (:209)
            if (lane == 0 || lane == 3) {
                const uint4 e = lTable[h];
                c = (int)e.x;
                v = (uint64_t)e.z << 32 | e.y;
            }
            if (lane == 1) {
                const uint2 e = sTable[h];
                c = (int)e.x;
                v = e.y;
            }
            if (lane == 2) v = ldg_u64_unaligned(src + s - repeat);       // :139
            if (lane != 2 && c == 0) v = src0;
            if (lane == 3 && h == hL) c = s, v = cv;  // lTable[hashL] = s precedes the s+1 probe (:124, :207)
            __syncwarp();
            if (lane == 0) l2_put_long(lTable, h, s, cv);
            if (lane == 1) l2_put_short(sTable, h, s, cv);
            const uint64_t repeatMask = 0xffffffffull << 8;
            bool f8 = false, f4 = false;
            if (lane == 0) {
                f8 = c > minSrcPos && cv == v;                            // :130
                f4 = c >= minSrcPos && (uint32_t)cv == (uint32_t)v;       // :199
            } else if (lane == 1) {
                f4 = c >= minSrcPos && (uint32_t)cv == (uint32_t)v;       // :204
            } else if (lane == 2) {
                f4 = repeat > 0 && (cv & repeatMask) == (v & repeatMask); // :139
            } else if (lane == 3) {
                f4 = c > minSrcPos && (uint32_t)(cv >> 8) == (uint32_t)v; // :209
            }
            const unsigned m8 = __ballot_sync(kFullMask, f8);
            const unsigned m4 = __ballot_sync(kFullMask, f4);

            if (m8 & 1u) {  // long candidate matches 8 bytes
                candidateL = __shfl_sync(kFullMask, c, 0);
                break;
            }
            if (m4 & 4u) {  // repeat at s+1 (:139-196)
                int base = s + 1;
                base -= extend_backward(src, base - repeat, base, nextEmit, lane);
                const int cand = s - repeat + 4 + 1;
                s = extend_to_end(src, n, s + 4 + 1, cand, lane);
                q.push(base, repeat | 3 << 24, s, lane);  // :147 test, literals and the repeat token: emit_group
                if (!q.flush(ep, dst, src, d, emitted, lane, sLimit, dstLimit, false)) return 0;
                nextEmit = s;
                if (s >= sLimit) goto emit_remainder;
                // index in-between (:183-195)
                l2_index_match<kL2OrderRepeat>(P(), lTable, sTable, src, base, s, lane);
                cv = ldg_u64_unaligned(src + s);
                continue;
            }
            if (m4 & 1u) {  // long candidate matches 4 bytes (:199)
                candidateL = __shfl_sync(kFullMask, c, 0);
                break;
            }
            if (m4 & 2u) {  // short candidate (:204-216): try the long table at s+1
                if (lane == 3) l2_put_long(lTable, h, s + 1, ldg_u64_unaligned(src + s + 1));
                __syncwarp();
                if (m4 & 8u) {
                    candidateL = __shfl_sync(kFullMask, c, 3);
                    s++;
                } else {
                    candidateL = __shfl_sync(kFullMask, c, 1);
                }
                break;
            }
            cv = ldg_u64_unaligned(src + nextS);  // :218
            s = nextS;
        }

        {
            const int back = extend_backward(src, candidateL, s, nextEmit, lane);  // :223-226
            candidateL -= back;
            s -= back;
        }
        {   // the :229 test travels with the record
            const int base = s;
            const int offset = base - candidateL;
            s = extend_to_end(src, n, s + 4, candidateL + 4, lane);  // :239-254

            if (offset > 65535 && s - base <= 4 && repeat != offset) {  // :257-264
                s = nextS + 1;
                if (s >= sLimit) goto emit_remainder;
                cv = ldg_u64_unaligned(src + s);
                continue_outer = true;
            }
            if (!continue_outer) {
                q.push(base, offset, s, lane);  // :266-289 and the :297 test: emit_group
                if (!q.flush(ep, dst, src, d, emitted, lane, sLimit, dstLimit, false)) return 0;
                repeat = offset;
                nextEmit = s;
                if (s >= sLimit) goto emit_remainder;  // :293

                // index short & long (:303-326)
                l2_index_match<kL2OrderGo>(P(), lTable, sTable, src, base, s, lane);
                cv = ldg_u64_unaligned(src + s);
            }
        }
    }

emit_remainder:  // :329-337
    if (!q.flush(ep, dst, src, d, emitted, lane, sLimit, dstLimit, true)) return 0;
    if (nextEmit < n) {
        if (d + n - nextEmit > dstLimit) return 0;
        d += emit_literal(dst + d, src + nextEmit, n - nextEmit, lane);
    }
    return d;
}

// ---- Asm flavour (MZCU_FLAVOR_AMD64) -----------------------------------------
// What amd64 runs for LevelBalanced: the functions generated by
// _generate/gen.go:1171-2038 genEncodeBetterBlockAsm, seven size classes
// (gen.go:78-88, encode_amd64.go:201-271).  Same walk as encodeBlockBetterGo; it
// differs in the margins (sLimit = len-17 or len-8, `>=` exits), the bail-out tests
// `d + lits + overhead >= dstLimit` (gen.go:1490-1508,1720-1737,1905-1918,1984-2003),
// the skip cap of 100 in the three large classes (gen.go:1327-1354), the far
// 4-byte-match rejection at offset > 65599 with no sLimit test (gen.go:1786-1801),
// candidates clamped to s-2162685 in the 8 MiB class (gen.go:1396-1419), per-class
// tables / hashes and the 64 KiB literal quirk (gen.go:2193-2200).
struct BetterAsmClass {
    static constexpr bool kGo = false;
    int lBits, sBits, skipLog, lHashBytes, maxSkip, outMargin, inMargin, ovh;
    bool quirk, far3, clamp;
    __device__ __forceinline__ uint32_t hashL(uint64_t u) const { return lHashBytes == 7 ? hash7(u, lBits) : hash6(u, lBits); }
    __device__ __forceinline__ uint32_t hashS(uint64_t u) const { return hash4(u, sBits); }
    __device__ static __forceinline__ BetterAsmClass for_len(int n) {
        BetterAsmClass q;
        q.lHashBytes = 7, q.maxSkip = 100, q.outMargin = 17, q.inMargin = 17, q.ovh = 4;
        q.quirk = false, q.far3 = true, q.clamp = false;
        if (n > (2 << 20)) {
            q.lBits = 17, q.sBits = 14, q.skipLog = 8, q.clamp = true;
        } else if (n > (512 << 10)) {
            q.lBits = 17, q.sBits = 14, q.skipLog = 7;
        } else {
            q.outMargin = 11, q.inMargin = 8;
            if (n > (64 << 10)) {
                q.lBits = 16, q.sBits = 13, q.skipLog = 7;
            } else {
                q.maxSkip = 0, q.lHashBytes = 6, q.far3 = false;
                if (n > (16 << 10)) q.lBits = 15, q.sBits = 12, q.skipLog = 6, q.quirk = true;
                else if (n > (4 << 10)) q.lBits = 14, q.sBits = 11, q.skipLog = 6, q.ovh = 3;
                else if (n > (1 << 10)) q.lBits = 12, q.sBits = 10, q.skipLog = 5, q.ovh = 3;
                else q.lBits = 11, q.sBits = 8, q.skipLog = 4, q.ovh = 3;
            }
        }
        return q;
    }
};

// The class of the benchmark's block sizes (512 KiB+1 .. 2 MiB, encodeBetterBlockAsm2MB:
// gen.go:80) with compile-time parameters; every other class runs on the runtime struct above.
struct BetterAsm2MB {
    static constexpr bool kGo = false;
    static constexpr int lBits = 17, sBits = 14, skipLog = 7, lHashBytes = 7, maxSkip = 100, outMargin = 17,
                         inMargin = 17, ovh = 4;
    static constexpr bool quirk = false, far3 = true, clamp = false;
    __device__ __forceinline__ uint32_t hashL(uint64_t u) const { return hash7(u, 17); }
    __device__ __forceinline__ uint32_t hashS(uint64_t u) const { return hash4(u, 14); }
};

template <class C>
__device__ int encode_l2_asm_block(const C P, uint8_t *dst, const uint8_t *src, const int n,
                                   uint4 *lTable, uint2 *sTable, uint32_t *rec_mem, const int lane) {
    const uint64_t src0 = ldg_u64_unaligned(src);  // the bytes an untouched entry (candidate 0) stands for
    L2Records q{rec_mem, 0, 0};
    const L2EmitAsm ep{P.ovh, P.quirk};
    int emitted = 0;  // nextEmit as the token writer sees it
    const int sLimit = n - P.inMargin;                    // gen.go:1272-1282
    const int dstLimit = n - P.outMargin - (n >> 5);      // gen.go:1284-1297
    int nextEmit = 0;
    int s = 1;
    int repeat = 1;
    int d = 0;

    for (;;) {
        // ---- search_loop (gen.go:1307-1690) ----
        const uint32_t skip = (uint32_t)(s - nextEmit) >> P.skipLog;
        const int nextS = (P.maxSkip == 0 || skip <= (uint32_t)(P.maxSkip - 1)) ? s + (int)skip + 1 : s + P.maxSkip;
        if (nextS >= sLimit) break;
        const uint64_t cv = ldg_u64_unaligned(src + s);
        const int minPos = s - kMaxCopy3Offset + 2;
        // lane 0: long(cv)  lane 1: short(cv)  lane 2: repeat  lane 3: long(cv>>8) (used only behind a short hit)
        uint32_t h = 0;
        int c = 0;
        if (lane == 0) h = P.hashL(cv);
        if (lane == 1) h = P.hashS(cv);
        if (lane == 3) h = P.hashL(cv >> 8);
        const uint32_t hL = __shfl_sync(kFullMask, h, 0);
        uint64_t v = 0;  // the candidate's bytes
        if (lane == 0 || lane == 3) {
            const uint4 e = lTable[h];
            c = (int)e.x;
            v = (uint64_t)e.z << 32 | e.y;
        }
        if (lane == 1) {
            const uint2 e = sTable[h];
            c = (int)e.x;
            v = e.y;
        }
        if (lane == 2) v = ldg_u64_unaligned(src + s - repeat);
        if (lane != 2 && c == 0) v = src0;
        if (lane == 3 && h == hL) c = s, v = cv;  // lTab[hash0] = s is stored before the s+1 probe reads
        __syncwarp();
        if (lane == 0) l2_put_long(lTable, h, s, cv);
        if (lane == 1) l2_put_short(sTable, h, s, cv);
        if (P.clamp && lane != 2 && lane < 4 && c <= minPos) {  // CMOVLLE: compared (and matched) at the clamped position
            c = minPos;
            v = ldg_u64_unaligned(src + c);  // not the entry's position any more: read the source (8 MiB class only)
        }
        bool f8 = false, f4 = false;
        if (lane == 0) {
            f8 = cv == v;
            f4 = (uint32_t)cv == (uint32_t)v;
        } else if (lane == 1) {
            f4 = (uint32_t)cv == (uint32_t)v;
        } else if (lane == 2) {
            f4 = ((cv ^ v) & (0xffffffffull << 8)) == 0;
        } else if (lane == 3) {
            f4 = (uint32_t)(cv >> 8) == (uint32_t)v;
        }
        const unsigned m8 = __ballot_sync(kFullMask, f8);
        const unsigned m4 = __ballot_sync(kFullMask, f4);

        int candidate;
        if (m8 & 1u) {
            candidate = __shfl_sync(kFullMask, c, 0);
        } else if (m4 & 4u) {  // repeat at s+1 (gen.go:1445-1622)
            int base = s + 1;
            base -= extend_backward(src, base - repeat, base, nextEmit, lane);
            s = extend_to_end(src, n, s + 5, s + 5 - repeat, lane);
            q.push(base, repeat | 3 << 24, s, lane);  // gen.go:1490-1508 test, literals, repeat token: emit_group
            if (!q.flush(ep, dst, src, d, emitted, lane, sLimit, dstLimit, false)) return 0;
            nextEmit = s;
            if (s >= sLimit) break;
            l2_index_match<kL2OrderRepeat>(P, lTable, sTable, src, base, s, lane);
            continue;
        } else if (m4 & 1u) {
            candidate = __shfl_sync(kFullMask, c, 0);
        } else if (m4 & 2u) {  // short match: try the long table at s+1 (gen.go:1665-1687)
            if (lane == 3) l2_put_long(lTable, h, s + 1, ldg_u64_unaligned(src + s + 1));
            __syncwarp();
            if (m4 & 8u) {
                candidate = __shfl_sync(kFullMask, c, 3);
                s++;
            } else {
                candidate = __shfl_sync(kFullMask, c, 1);
            }
        } else {
            s = nextS;
            continue;
        }

        // ---- candidate_match (gen.go:1692-1918) ----
        {
            const int back = extend_backward(src, candidate, s, nextEmit, lane);
            candidate -= back;
            s -= back;
        }
        const int base = s;  // the gen.go:1720-1737 test travels with the record
        const int offset = base - candidate;
        s = extend_to_end(src, n, s + 4, candidate + 4, lane);
        if (P.far3 && s - base == 4 && offset > kMaxCopy2Offset && offset != repeat) {  // gen.go:1786-1801
            s = nextS + 1;
            continue;
        }
        repeat = offset;
        q.push(base, offset, s, lane);  // gen.go:1803-1897 and the :1905-1918 test: emit_group
        if (!q.flush(ep, dst, src, d, emitted, lane, sLimit, dstLimit, false)) return 0;
        nextEmit = s;
        if (s >= sLimit) break;
        l2_index_match<kL2OrderAsm>(P, lTable, sTable, src, base, s, lane);  // gen.go:1921-1978
    }

    // emit_remainder (gen.go:1980-2017): the bail test runs even when nothing is left
    if (!q.flush(ep, dst, src, d, emitted, lane, sLimit, dstLimit, true)) return 0;
    if (d + (n - nextEmit) + P.ovh >= dstLimit) return 0;
    d += emit_literal(dst + d, src + nextEmit, n - nextEmit, lane, P.quirk);
    return d;
}

// ---- the probe-window walk ------------------------------------------------------
// A step of the LevelBalanced walk is one DRAM round trip (its table probes) plus ~1 us of
// dependent instructions, 123 k times per 1 MiB block, and the profile of the step-at-a-time walk
// above says where the time goes: 41 % waiting for the probes, 17 % for other loads, the rest on
// the warp's own instruction chain (profiles/r02_ncu_encode_l2_step_walk.txt).  This walk probes a
// WINDOW of 32 consecutive positions per round trip -- lane j holds, for position wb + j, its 8
// source bytes, both hashes and both table entries -- and steps through the window on the cached
// entries: 4.3 steps per round trip on the benchmark's blocks (profiles/r02_l2_walk_sim.txt).
// Exactness: every table insert the walk performs while a window is live is also applied to the
// window ("patch"), so a cached entry always equals what a load would return at that moment:
//   * the step's own inserts (long[s], short[s]) and the deferred long[s+1] patch the later lanes
//     that share the slot (found once per window with match.any on the hashes; rare);
//   * the inserts of a match's interior are positions of the window itself, so their hashes are
//     the lanes' own: every lane knows in closed form whether its position was inserted and with
//     which serial rank, and a later lane takes the highest-ranked inserted lane of its hash group;
//   * inserts outside the window end it (behind it: the window is reloaded; beyond it: the walk
//     has left it anyway).
// The repeat check needs 4 bytes at p + 1 - repeat; they are loaded per window and again after
// every match (the repeat offset changed), off the critical path when the next step hits on 8 bytes.
#ifndef MZ_L2_REPLAY
#define MZ_L2_REPLAY 1
#endif

template <bool kSmall>
struct L2GoClass {
    static constexpr bool kGo = true;
    static constexpr int skipLog = 7, maxSkip = 0, ovh = 0, inMargin = kInputMargin, outMargin = 6;
    static constexpr bool quirk = false, far3 = true, clamp = false;
    __device__ __forceinline__ uint32_t hashL(uint64_t u) const { return kSmall ? hash6(u, 15) : hash7(u, 17); }
    __device__ __forceinline__ uint32_t hashS(uint64_t u) const { return hash4(u, kSmall ? 12 : 14); }
};

// membership and serial rank of position p among the inserts of one match (see l2_index_match)
template <int kOrder>
__device__ __forceinline__ void l2_insert_rank(int p, int base, int e, int *rankL, int *rankS) {
    int rl = -1, rs = -1;
    if (kOrder == kL2OrderRepeat) {
        const int span = e - base - 3;
        const int J = span > 0 ? (span + 3) >> 2 : 0;
        const int a = p - (base + 1), b = (e - 2) - p;
        if (a >= 0 && !(a & 1) && (a >> 1) < J) rl = 4 * (a >> 1);
        if (b >= 0 && !(b & 1) && (b >> 1) < J) rl = max(rl, 4 * (b >> 1) + 2);
        const int a1 = a - 1, b1 = b + 1;  // short inserts sit one position behind a long one
        if (a1 >= 0 && !(a1 & 1) && (a1 >> 1) < J) rs = 4 * (a1 >> 1) + 1;
        if (b1 >= 0 && !(b1 & 1) && (b1 >> 1) < J) rs = max(rs, 4 * (b1 >> 1) + 3);
    } else {
        const int mid = (base + 2 + e - 3 + 1) >> 1;
        const int left = e - 3 - mid;
        const int K = left > 0 ? (left + 1) >> 1 : 0;
        if (p == base + 1) rl = 0;
        if (p == e - 2) rl = max(rl, kOrder == kL2OrderAsm ? 1 : 2);
        const int a = p - (base + 2), b = p - mid;
        if (a >= 0 && !(a & 1) && (a >> 1) < K) rl = max(rl, 4 + 2 * (a >> 1));
        if (b >= 0 && !(b & 1) && (b >> 1) < K) rl = max(rl, 5 + 2 * (b >> 1));
        if (p == base + 2) rs = kOrder == kL2OrderAsm ? 2 : 1;
        if (p == e - 1) rs = max(rs, 3);
    }
    *rankL = rl;
    *rankS = rs;
}

template <class C>
__device__ int encode_l2_walk(const C P, uint8_t *dst, const uint8_t *src, const int n, uint4 *lTable, uint2 *sTable,
                              uint32_t *rec_mem, const int lane) {
    constexpr bool kGo = C::kGo;
    const uint64_t src0 = ldg_u64_unaligned(src);  // the bytes an untouched entry (candidate 0) stands for
    L2Records q{rec_mem, 0, 0};
    const L2EmitGo epg{};
    const L2EmitAsm epa{P.ovh, P.quirk};
    int emitted = 0;  // nextEmit as the token writer sees it
    const int sLimit = kGo ? n - kInputMargin : n - P.inMargin;                      // :65 / gen.go:1272-1282
    const int dstLimit = kGo ? n - (n >> 5) - 6 : n - P.outMargin - (n >> 5);        // :92 / gen.go:1284-1297
    int nextEmit = 0, s = 1, repeat = 1, d = 0;
    auto flush = [&](bool all) -> bool {
        return kGo ? q.flush(epg, dst, src, d, emitted, lane, sLimit, dstLimit, all)
                   : q.flush(epa, dst, src, d, emitted, lane, sLimit, dstLimit, all);
    };

    // ---- the window: lane j <-> position wb + j ----
    int wb = 0;
    bool wvalid = false;
    int repw = 0;          // the repeat offset rep4 was loaded for (0 = none)
    uint64_t cv = 0;       // src[p .. p+8)
    uint32_t hL = 0, hS = 0;
    int cL = 0, cS = 0;    // cached table entries of my position's slots
    uint64_t vL = 0;
    uint32_t vS = 0, rep4 = 0;
    unsigned sameL = 0, sameS = 0;
    const unsigned self = 1u << lane;

    for (;;) {
        // ---- next position (encode_l2.go:114-117; gen.go:1307-1354) ----
        const uint32_t skip = (uint32_t)(s - nextEmit) >> P.skipLog;
        const int nextS = (P.maxSkip == 0 || skip <= (uint32_t)(P.maxSkip - 1)) ? s + (int)skip + 1 : s + P.maxSkip;
        if (kGo ? nextS > sLimit : nextS >= sLimit) break;

        // ---- (re)load the window when the step and its s+1 probe are not inside it ----
        int L = s - wb;
        if (!wvalid || L < 0 || L > 30) {
            __syncwarp();  // the stores of earlier inserts are ordered before these loads
            wb = s;
            L = 0;
            wvalid = true;
            const int p = wb + lane;
            cv = ld64_clamped(src, p, n);
            // the next window's source bytes: its hashes (and so its probes) wait for them
            if (lane < 2 && wb + 192 + 128 * lane < n) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + wb + 192 + 128 * lane));
            hL = P.hashL(cv);
            hS = P.hashS(cv);
            const uint4 el = lTable[hL];
            bool fetch_s = true;
#if MZ_L2_STAGS
            // (a far candidate of the clamped class is compared at its clamped position: always fetch there)
            fetch_s = ((l2_stags()[hS >> 5] >> (hS & 31)) & 1u) == l2_tag1((uint32_t)cv) ||
                      (!kGo && P.clamp && p >= kMaxCopy3Offset - 2);
#endif
            uint2 es = make_uint2(0u, ~(uint32_t)cv);  // not fetched: cannot verify
            if (fetch_s) es = sTable[hS];
            repw = repeat;
            rep4 = (p + 5 <= n && p + 1 >= repeat) ? ldg_u32_unaligned(src + p + 1 - repeat) : 0;
            cL = (int)el.x;
            vL = (uint64_t)el.z << 32 | el.y;
            cS = (int)es.x;
            vS = es.y;
            if (cL == 0) vL = src0;
            if (cS == 0 && fetch_s) vS = (uint32_t)src0;
            sameL = __match_any_sync(kFullMask, hL);
            sameS = __match_any_sync(kFullMask, hS);
        }
        const int p = wb + lane;
        if (repw != repeat) {  // the repeat offset changed since the bytes for its check were loaded
            repw = repeat;
            rep4 = (p + 5 <= n && p + 1 >= repeat) ? ldg_u32_unaligned(src + p + 1 - repeat) : 0;
        }

        // ---- every lane answers for its own position (encode_l2.go:118-216; gen.go:1356-1690) ----
        int kL = cL, kS = cS;  // candidates as the compares see them
        uint64_t wL = vL;
        uint32_t wS = vS;
        int kN = cL;           // ... and as the long probe at s+1 of the step at p-1 sees them
        uint64_t wN = vL;
        bool f8, f4l, f4s, fn;
        if (kGo) {
            const int minSrcPos = p - kMaxCopy3Offset + 1;  // :118
            f8 = kL > minSrcPos && cv == wL;                                   // :130
            f4l = kL >= minSrcPos && (uint32_t)cv == (uint32_t)wL;             // :199
            f4s = kS >= minSrcPos && (uint32_t)cv == wS;                       // :204
            fn = kN > minSrcPos - 1 && (uint32_t)cv == (uint32_t)wN;           // :209 (minSrcPos of the step at p-1)
        } else {
            if (P.clamp) {  // CMOVLLE: compared (and matched) at the clamped position (8 MiB class only)
                const int minPos = p - kMaxCopy3Offset + 2;
                if (kL <= minPos) kL = minPos, wL = ldg_u64_unaligned(src + kL);
                if (kS <= minPos) kS = minPos, wS = ldg_u32_unaligned(src + kS);
                if (kN <= minPos - 1) kN = minPos - 1, wN = ldg_u64_unaligned(src + kN);
            }
            f8 = cv == wL;
            f4l = (uint32_t)cv == (uint32_t)wL;
            f4s = (uint32_t)cv == wS;
            fn = (uint32_t)cv == (uint32_t)wN;
        }
        const unsigned m8 = __ballot_sync(kFullMask, f8), m4l = __ballot_sync(kFullMask, f4l),
                       m4s = __ballot_sync(kFullMask, f4s);
        unsigned mn = __ballot_sync(kFullMask, fn);
        const unsigned dupL = __ballot_sync(kFullMask, (sameL & ~self) != 0), dupS = __ballot_sync(kFullMask, (sameS & ~self) != 0);
        const unsigned bitL = 1u << L;

        // ---- the step's inserts: long[s] = s, short[s] = s (before the compares' consequences, :124-125) ----
        if (lane == L) {
            l2_put_long(lTable, hL, s, cv);
            l2_put_short(sTable, hS, s, cv);
        }
        if ((dupL | dupS) & bitL) {  // later lanes of the window share a slot with s: they now see s there
            const uint64_t cvs = __shfl_sync(kFullMask, cv, L);
            const unsigned gl = __shfl_sync(kFullMask, sameL, L), gs = __shfl_sync(kFullMask, sameS, L);
            if (lane > L && (gl & self)) cL = s, vL = cvs;
            if (lane > L && (gs & self)) cS = s, vS = (uint32_t)cvs;
            if (gl & (bitL << 1)) {  // the s+1 probe reads lTable after lTable[hashL] = s (:207)
                bool f = false;
                if (lane == L + 1) {
                    int k2 = cL;
                    uint64_t w2 = vL;
                    if (!kGo && P.clamp && k2 <= p - 1 - kMaxCopy3Offset + 2) k2 = p - 1 - kMaxCopy3Offset + 2, w2 = ldg_u64_unaligned(src + k2);
                    f = (kGo ? k2 > p - 1 - kMaxCopy3Offset + 1 : true) && (uint32_t)cv == (uint32_t)w2;
                    kN = k2;
                    wN = w2;
                }
                mn = (mn & ~(bitL << 1)) | (__ballot_sync(kFullMask, f) & (bitL << 1));
            }
        }

        int kind;  // 0: long candidate of s, 1: short candidate of s, 2: long candidate of s+1
        if (m8 & bitL) {
            kind = 0;
        } else if (__ballot_sync(kFullMask, (!kGo || repeat > 0) && (uint32_t)(cv >> 8) == rep4) & bitL) {
            // repeat at s+1 (:139-196; gen.go:1445-1622).  The check is evaluated only when the 8-byte test
            // failed: its bytes were requested after the previous match and need not have landed before
            int base = s + 1;
            base -= extend_backward(src, base - repeat, base, nextEmit, lane);
            s = extend_to_end(src, n, s + 5, s + 5 - repeat, lane);
            q.push(base, repeat | 3 << 24, s, lane);
            if (!flush(false)) return 0;
            nextEmit = s;
            if (s >= sLimit) break;
            l2_index_match<kL2OrderRepeat>(P, lTable, sTable, src, base, s, lane);
            if (s - wb <= 30) {
                if (base + 1 < wb) {
                    wvalid = false;  // inserts behind the window: their hashes are not at hand
                } else {
                    int rl, rs;
                    l2_insert_rank<kL2OrderRepeat>(p, base, s, &rl, &rs);
                    const unsigned QL = __ballot_sync(kFullMask, rl >= 0), QS = __ballot_sync(kFullMask, rs >= 0);
                    unsigned candl = p >= s ? (QL & sameL) : 0u, cands = p >= s ? (QS & sameS) : 0u;
                    int bestl = -1, bests = -1;
                    while (__any_sync(kFullMask, (candl | cands) != 0)) {
                        const int jl = candl ? __ffs(candl) - 1 : lane, js = cands ? __ffs(cands) - 1 : lane;
                        const int kl = __shfl_sync(kFullMask, rl, jl) << 5 | jl, ks = __shfl_sync(kFullMask, rs, js) << 5 | js;
                        if (candl) bestl = max(bestl, kl), candl &= candl - 1;
                        if (cands) bests = max(bests, ks), cands &= cands - 1;
                    }
                    const uint64_t cvl = __shfl_sync(kFullMask, cv, bestl >= 0 ? bestl & 31 : lane);
                    const uint64_t cvs2 = __shfl_sync(kFullMask, cv, bests >= 0 ? bests & 31 : lane);
                    if (bestl >= 0) cL = wb + (bestl & 31), vL = cvl;
                    if (bests >= 0) cS = wb + (bests & 31), vS = (uint32_t)cvs2;
                }
            }
            continue;
        } else if (m4l & bitL) {
            kind = 0;
        } else if (m4s & bitL) {  // short candidate: try the long table at s+1 (:204-216; gen.go:1665-1687)
            // lTable[hashL(s+1)] = s+1, whatever the s+1 compare says
            if (lane == L + 1) l2_put_long(lTable, hL, s + 1, cv);
            if (dupL & (bitL << 1)) {
                const uint64_t cv1 = __shfl_sync(kFullMask, cv, L + 1);
                const unsigned g1 = __shfl_sync(kFullMask, sameL, L + 1);
                if (lane > L + 1 && (g1 & self)) cL = s + 1, vL = cv1;
            }
            kind = (mn & (bitL << 1)) ? 2 : 1;
        } else {
            s = nextS;  // :218
            continue;
        }

        // ---- a verified candidate (:223-326; gen.go:1692-1978) ----
        const int from = kind == 2 ? L + 1 : L;
        int candidate = __shfl_sync(kFullMask, kind == 1 ? kS : (kind == 2 ? kN : kL), from);
        // what the 8-byte snapshot of a long candidate already says about the match length:
        // 4..7 = it ends there, 8 = at least 8 (a short-table candidate only vouches for 4)
        int flen = 4;
        {
            const uint64_t x = cv ^ (kind == 2 ? wN : wL);
            if (kind != 1) flen = (uint32_t)(x >> 32) ? 4 + ((__ffs((int)(uint32_t)(x >> 32)) - 1) >> 3) : 8;
            flen = __shfl_sync(kFullMask, flen, from);
        }
        if (kind == 2) s++;
        // backward (:223-226) and forward (:239-254) extension in ONE round trip: the byte before the
        // candidate is requested together with the forward bytes; only when it matches (5 %) does the
        // backward loop run
        bool beq = false;
        if (lane == 0 && candidate > 0 && s > nextEmit) beq = src[candidate - 1] == src[s - 1];
        const int pe = (kind != 1 && flen < 8) ? s + flen : extend_to_end(src, n, s + flen, candidate + flen, lane);
        beq = __shfl_sync(kFullMask, (int)beq, 0);
        if (beq) {
            const int back = extend_backward(src, candidate, s, nextEmit, lane);
            candidate -= back;
            s -= back;
        }
        const int base = s;
        const int offset = base - candidate;
        s = pe;
        if (kGo ? (offset > 65535 && s - base <= 4 && repeat != offset)                              // :257-264
                : (P.far3 && s - base == 4 && offset > kMaxCopy2Offset && offset != repeat)) {       // gen.go:1786-1801
            s = nextS + 1;
            if (kGo && s >= sLimit) break;
            continue;
        }
        repeat = offset;
        q.push(base, offset, s, lane);
        if (!flush(false)) return 0;
        nextEmit = s;
        if (s >= sLimit) break;
        l2_index_match<kGo ? kL2OrderGo : kL2OrderAsm>(P, lTable, sTable, src, base, s, lane);
        if (s - wb <= 30) {  // the window lives on: apply the match's inserts to it
            if (base + 1 < wb) {
                wvalid = false;
            } else {
                int rl, rs;
                l2_insert_rank<kGo ? kL2OrderGo : kL2OrderAsm>(p, base, s, &rl, &rs);
                const unsigned QL = __ballot_sync(kFullMask, rl >= 0), QS = __ballot_sync(kFullMask, rs >= 0);
                unsigned candl = p >= s ? (QL & sameL) : 0u, cands = p >= s ? (QS & sameS) : 0u;
                int bestl = -1, bests = -1;
                while (__any_sync(kFullMask, (candl | cands) != 0)) {
                    const int jl = candl ? __ffs(candl) - 1 : lane, js = cands ? __ffs(cands) - 1 : lane;
                    const int kl = __shfl_sync(kFullMask, rl, jl) << 5 | jl, ks = __shfl_sync(kFullMask, rs, js) << 5 | js;
                    if (candl) bestl = max(bestl, kl), candl &= candl - 1;
                    if (cands) bests = max(bests, ks), cands &= cands - 1;
                }
                const uint64_t cvl = __shfl_sync(kFullMask, cv, bestl >= 0 ? bestl & 31 : lane);
                const uint64_t cvs2 = __shfl_sync(kFullMask, cv, bests >= 0 ? bests & 31 : lane);
                if (bestl >= 0) cL = wb + (bestl & 31), vL = cvl;
                if (bests >= 0) cS = wb + (bests & 31), vS = (uint32_t)cvs2;
            }
        }
    }

    // emit_remainder (:329-337; gen.go:1980-2017: there the bail test runs even when nothing is left)
    if (!flush(true)) return 0;
    if (kGo) {
        if (nextEmit < n) {
            if (d + n - nextEmit > dstLimit) return 0;
            d += emit_literal(dst + d, src + nextEmit, n - nextEmit, lane);
        }
    } else {
        if (d + (n - nextEmit) + P.ovh >= dstLimit) return 0;
        d += emit_literal(dst + d, src + nextEmit, n - nextEmit, lane, P.quirk);
    }
    return d;
}

__global__ void __launch_bounds__(kEncL2Warps * 32, MZ_ENC_L2_MIN_CTAS)
encode_l2_asm_kernel(int nblk, const uint8_t *__restrict__ src, const uint64_t *__restrict__ sbeg,
                     const uint64_t *__restrict__ send, uint8_t *dst, const uint64_t *__restrict__ dbeg,
                     uint32_t *__restrict__ out_len, int *counter, uint32_t *tables) {
    __shared__ uint32_t rec_rings[kEncL2Warps][3 * kRecRing];
    const int lane = kEncL2Warps == 1 ? (int)threadIdx.x : lane_id();
    const int warp = kEncL2Warps == 1 ? 0 : (int)(threadIdx.x >> 5);
    const int gwarp = blockIdx.x * kEncL2Warps + warp;
    uint4 *lTable = reinterpret_cast<uint4 *>(tables + (size_t)gwarp * (kEncL2WsBytesPerWarp / 4));
    uint2 *sTable = reinterpret_cast<uint2 *>(lTable + (1 << 17));
    for (;;) {
        int blk = 0;
        if (lane == 0) blk = atomicAdd(counter, 1);
        blk = __shfl_sync(kFullMask, blk, 0);
        if (blk >= nblk) return;
        const uint8_t *sp = src + sbeg[blk];
        const int64_t n64 = (int64_t)(send[blk] - sbeg[blk]);
        uint8_t *dp = dst + dbeg[blk];
        int res = 0;
        if (n64 > kMinNonLiteralBlockSize && n64 <= kMaxBlockSize) {  // encode_amd64.go:260
            const int n = (int)n64;
            const BetterAsmClass cls = BetterAsmClass::for_len(n);
            for (int i = lane; i < (1 << cls.lBits); i += 32) lTable[i] = make_uint4(0, 0, 0, 0);
            uint4 *s4 = reinterpret_cast<uint4 *>(sTable);
            for (int i = lane; i < (1 << cls.sBits) / 2; i += 32) s4[i] = make_uint4(0, 0, 0, 0);
#if MZ_L2_STAGS
            {   // untouched short entries stand for candidate 0: every tag starts as the tag of src[0..4)
                const uint32_t t0 = l2_tag1(ldg_u32_unaligned(sp)) ? 0xffffffffu : 0u;
                uint32_t *tg = l2_stags();
                for (int i = lane; i < (1 << 14) / 32; i += 32) tg[i] = t0;
            }
#endif
            __syncwarp();
#if MZ_L2_REPLAY
            res = (n > (512 << 10) && n <= (2 << 20))
                      ? encode_l2_walk(BetterAsm2MB(), dp, sp, n, lTable, sTable, rec_rings[warp], lane)
                      : encode_l2_walk(cls, dp, sp, n, lTable, sTable, rec_rings[warp], lane);
#else
            res = (n > (512 << 10) && n <= (2 << 20))
                      ? encode_l2_asm_block(BetterAsm2MB(), dp, sp, n, lTable, sTable, rec_rings[warp], lane)
                      : encode_l2_asm_block(cls, dp, sp, n, lTable, sTable, rec_rings[warp], lane);
#endif
        }
        if (lane == 0) out_len[blk] = (uint32_t)res;
        __syncwarp();
    }
}

__global__ void __launch_bounds__(kEncL2Warps * 32, MZ_ENC_L2_MIN_CTAS)
encode_l2_kernel(int nblk, const uint8_t *__restrict__ src, const uint64_t *__restrict__ sbeg,
                 const uint64_t *__restrict__ send, uint8_t *dst, const uint64_t *__restrict__ dbeg,
                 uint32_t *__restrict__ out_len, int *counter, uint32_t *tables) {
    __shared__ uint32_t rec_rings[kEncL2Warps][3 * kRecRing];
    const int lane = kEncL2Warps == 1 ? (int)threadIdx.x : lane_id();
    const int warp = kEncL2Warps == 1 ? 0 : (int)(threadIdx.x >> 5);
    const int gwarp = blockIdx.x * kEncL2Warps + warp;
    uint4 *lTable = reinterpret_cast<uint4 *>(tables + (size_t)gwarp * (kEncL2WsBytesPerWarp / 4));
    uint2 *sTable = reinterpret_cast<uint2 *>(lTable + (1 << 17));
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
            const bool small = n <= (64 << 10);
            // zero both tables (the small variant uses the first 2^15 / 2^12 entries)
            const int lents = small ? (1 << 15) : (1 << 17);
            const int sents = small ? (1 << 12) : (1 << 14);
            for (int i = lane; i < lents; i += 32) lTable[i] = make_uint4(0, 0, 0, 0);
            uint4 *s4 = reinterpret_cast<uint4 *>(sTable);
            for (int i = lane; i < sents / 2; i += 32) s4[i] = make_uint4(0, 0, 0, 0);
#if MZ_L2_STAGS
            {   // untouched short entries stand for candidate 0: every tag starts as the tag of src[0..4)
                const uint32_t t0 = l2_tag1(ldg_u32_unaligned(sp)) ? 0xffffffffu : 0u;
                uint32_t *tg = l2_stags();
                for (int i = lane; i < (1 << 14) / 32; i += 32) tg[i] = t0;
            }
#endif
            __syncwarp();
#if MZ_L2_REPLAY
            res = small ? encode_l2_walk(L2GoClass<true>(), dp, sp, n, lTable, sTable, rec_rings[warp], lane)
                        : encode_l2_walk(L2GoClass<false>(), dp, sp, n, lTable, sTable, rec_rings[warp], lane);
#else
            res = small ? encode_l2_block<true>(dp, sp, n, lTable, sTable, rec_rings[warp], lane)
                        : encode_l2_block<false>(dp, sp, n, lTable, sTable, rec_rings[warp], lane);
#endif
        }
        if (lane == 0) out_len[blk] = (uint32_t)res;
        __syncwarp();
    }
}

}  // namespace mz
