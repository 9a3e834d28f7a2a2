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
//     32-byte record {position, 4 bytes before it, 24 bytes from it}.  A probe then returns the candidate AND the bytes needed to
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
//   * a random 32-byte probe costs a whole 128-byte DRAM line on B200
//     (profiles/r01_micro_gather32.txt), so a 1-bit tag per slot in shared memory
//     keeps the lanes whose probe cannot verify from fetching theirs (-35 % DRAM
//     bytes, profiles/r02_tag_filter_variants.txt).
//   * what finally bounds it: every block is one warp and all of
//     them run at once, so the kernel's time is ONE warp's chain -- 29 952 round
//     trips per 1 MiB block, each a loaded DRAM round trip plus ~620 dependent
//     instructions -- not the bytes moved (63 % issue utilisation, 45 % of the
//     DRAM's copy rate).
#pragma once

#include "mz_common.cuh"

namespace mz {

// Forward match extension, restating the Go loop
//   for s <= limit { if diff := load64(s)^load64(cand); diff != 0 { s += tz>>3; break }; s += 8; cand += 8 }
// with the lanes comparing consecutive 8-byte words; the first round uses 4
// lanes (most matches end within 32 bytes, and the candidate side is a random
// DRAM sector), later rounds all 32.
// ---- arrival gate (host-pointer encode, mz_api.cu) ---------------------------
// When the source is still arriving over PCIe the host copies it in slices
// (bytes [k*slice, (k+1)*slice) of EVERY block, then a 4-byte write of k+1 to
// *gate), and the kernel chases the arrival front: before reading source bytes
// ahead of the cursor a warp waits until enough slices have landed.  gate ==
// nullptr: everything is resident.  Returns the number of bytes of each block
// that have arrived.
__device__ __forceinline__ int gate_wait(const int *gate, int slice, int need) {
    for (;;) {
        int k;
        asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(k) : "l"(gate) : "memory");
        const long long have = (long long)k * slice;
        if (have >= need) return have > 0x7fffffffll ? 0x7fffffff : (int)have;
        __nanosleep(256);
    }
}

__device__ __noinline__ int extend_forward8(const uint8_t *src, int s, int cand, int limit, int lane, const int *gate,
                                            int slice) {
    int width = 4;
    for (;;) {
        if (gate) gate_wait(gate, slice, min(s + 8 * width + 16, limit + 8));
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

// Asm flavour: matchLen (gen.go:3190-3288) -- the exact common prefix of src[s..] and
// src[cand..], running to the END of the block (n), 8 bytes per lane per round.
__device__ __noinline__ int extend_forward_exact(const uint8_t *src, int s, int cand, int n, int lane, const int *gate,
                                                 int slice) {
    int width = 4;
    for (;;) {
        if (s >= n) return n;
        if (gate) gate_wait(gate, slice, min(s + 8 * width + 16, n));
        const int pos = s + 8 * lane;
        const bool act = lane < width;
        const int valid = min(max(n - pos, 0), 8);  // bytes of this lane's word inside the block
        uint64_t diff = 0;
        if (act && valid == 8) {
            diff = ldg_u64_unaligned(src + pos) ^ ldg_u64_unaligned(src + cand + 8 * lane);
        } else if (act) {
            for (int i = 0; i < valid; i++)
                diff |= (uint64_t)(uint8_t)(src[pos + i] ^ src[cand + 8 * lane + i]) << (8 * i);
            diff |= 1ull << (8 * valid);  // the byte at n never matches (valid < 8)
        }
        const unsigned stop = __ballot_sync(kFullMask, act && diff != 0);
        if (stop == 0) {
            s += 8 * width;
            cand += 8 * width;
            width = 32;
            continue;
        }
        const int f = __ffs(stop) - 1;
        const int add = __shfl_sync(kFullMask, (__ffsll((long long)diff) - 1) >> 3, f);
        return s + 8 * f + add;
    }
}

// Writes the low n bytes of tok at dst (n <= 8), one byte per lane.
__device__ __forceinline__ void put_token(uint8_t *dst, uint64_t tok, int n, int lane) {
    if (lane < n) dst[lane] = (uint8_t)__byte_perm((uint32_t)tok, (uint32_t)(tok >> 32), lane);
}

__device__ __forceinline__ void copy_bytes(uint8_t *dst, const uint8_t *src, int n, int lane) {
    for (int i = lane; i < n; i += 32) dst[i] = src[i];
}

// emitLiteral asm_none.go:84-122
// quirk: the 64 KiB Asm classes write runs of 286+ bytes with a 3-byte length (gen.go:2193-2200)
__device__ __forceinline__ int emit_literal(uint8_t *dst, const uint8_t *lit, int n, int lane, bool quirk = false) {
    if (n == 0) return 0;
    int hn;
    uint64_t h = tok_literal_hdr((uint32_t)n, &hn);
    if (quirk && hn == 3) {
        hn = 4;
        h = (h & ~0xffull) | (31u << 3 | kTagLiteral);
    }
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

// ---- walk parameters ----------------------------------------------------------
// Two FLAVOURS of the same walk exist in the reference, chosen by build tags:
//   Go    the pure-Go functions (`-tags noasm`, non-amd64/arm64 platforms):
//         encode_l1.go / encode_l0.go, two size classes;
//   Asm   what amd64 runs: the functions generated by _generate/gen.go:257-1155
//         (asm_amd64.s encodeBlockAsm* / encodeFastBlockAsm*), seven size classes
//         (gen.go:57-76, encode_amd64.go:37-189).  Same walk; it differs in the
//         margins (sLimit = len-17, `>=` exits), the bail-out tests against dstLimit
//         (gen.go:380-417), byte-exact match extension to the end of the block
//         (matchLen, gen.go:632-654,859-887), the per-class table/skip/hash/step,
//         a far-candidate clamp in the 8 MiB class (gen.go:466-490) and a
//         3-byte literal length quirk in the 64 KiB class (gen.go:2193-2200).
// The kernel body is written once over a parameter object: compile-time
// constants for the hot classes, runtime fields for the small Asm classes.

// LevelFastest, Go flavour: encode_l1.go:39-283 (large) and :285-524 (<= 64 KiB)
template <bool kSmall>
struct L1Params {
    static constexpr bool kAsm = false;
    static constexpr bool kMayClamp = false;
    static constexpr bool kBalanced = false;
    static constexpr int kMinMatch = 4;         // candidates are verified on 4 bytes
    static constexpr bool kBackExtend = true;   // :169-172
    __device__ __forceinline__ int table_bits() const { return kSmall ? 13 : 15; }
    __device__ __forceinline__ int skip_log() const { return kSmall ? 5 : 6; }
    __device__ __forceinline__ int step() const { return 4; }
    __device__ __forceinline__ int max_fuse_lits() const { return kSmall ? kMaxCopy2Lits : kMaxCopy3Lits; }
    __device__ __forceinline__ int max_fuse_lits2() const { return max_fuse_lits(); }
    __device__ __forceinline__ int max_fuse_lits3() const { return max_fuse_lits(); }
    __device__ __forceinline__ int s_limit(int n) const { return n - kInputMargin; }
    __device__ __forceinline__ int dst_limit(int n) const { return n - (n >> 5) - 6; }
    __device__ __forceinline__ int lit_overhead() const { return 0; }
    __device__ __forceinline__ bool lit_quirk() const { return false; }
    __device__ __forceinline__ uint32_t hash(uint64_t u) const { return kSmall ? hash5(u, 13) : hash6(u, 15); }
};

// LevelFastest, Asm flavour, blocks > 512 KiB (encodeBlockAsm2MB / encodeBlockAsm:
// gen.go:58-59: 15 bits, skipLog 6, hash6, step 4) -- the benchmark's class.
template <bool kClamp>  // kClamp: the 8 MiB class (encodeBlockAsm), whose far candidates are clamped
struct L1AsmBigParams {
    static constexpr bool kAsm = true;
    static constexpr bool kMayClamp = kClamp;
    static constexpr bool kBalanced = false;
    static constexpr int kMinMatch = 4;
    static constexpr bool kBackExtend = true;
    __device__ __forceinline__ int table_bits() const { return 15; }
    __device__ __forceinline__ int skip_log() const { return 6; }
    __device__ __forceinline__ int step() const { return 4; }
    __device__ __forceinline__ int max_fuse_lits() const { return 3; }            // gen.go:907
    __device__ __forceinline__ int max_fuse_lits2() const { return 3; }
    __device__ __forceinline__ int max_fuse_lits3() const { return 3; }
    __device__ __forceinline__ int s_limit(int n) const { return n - 17; }        // gen.go:369
    __device__ __forceinline__ int dst_limit(int n) const { return n - 17 - (n >> 5); }  // gen.go:380-391
    __device__ __forceinline__ int lit_overhead() const { return 4; }             // gen.go:1157-1169
    __device__ __forceinline__ bool lit_quirk() const { return false; }
    __device__ __forceinline__ uint32_t hash(uint64_t u) const { return hash6(u, 15); }
};

// LevelSuperFast, Asm flavour, blocks of 64 KiB+1 .. 2 MiB (encodeFastBlockAsm512K / ...2MB:
// gen.go:70-71: 13 bits, skipLog 5, hash8, step 4; Fast options gen.go:68).
struct L0AsmBigParams {
    static constexpr bool kAsm = true;
    static constexpr bool kMayClamp = false;
    static constexpr bool kBalanced = false;
    static constexpr int kMinMatch = 8;
    static constexpr bool kBackExtend = false;
    __device__ __forceinline__ int table_bits() const { return 13; }
    __device__ __forceinline__ int skip_log() const { return 5; }
    __device__ __forceinline__ int step() const { return 4; }
    __device__ __forceinline__ int max_fuse_lits() const { return 0; }
    __device__ __forceinline__ int max_fuse_lits2() const { return 0; }
    __device__ __forceinline__ int max_fuse_lits3() const { return 0; }
    __device__ __forceinline__ int s_limit(int n) const { return n - 17; }
    __device__ __forceinline__ int dst_limit(int n) const { return n - 17 - (n >> 3); }
    __device__ __forceinline__ int lit_overhead() const { return 4; }
    __device__ __forceinline__ bool lit_quirk() const { return false; }
    __device__ __forceinline__ uint32_t hash(uint64_t u) const { return hash8(u, 13); }
};

// Asm flavour, every other size class of gen.go:57-76 (runtime fields).
// kMatch8 = the Fast (LevelSuperFast) options: 8-byte compares, hash8, no literal
// fusing, no backward extension, dstLimit from len>>3 (gen.go:68).
template <bool kMatch8>
struct AsmClassParams {
    static constexpr bool kAsm = true;
    static constexpr bool kMayClamp = kMatch8;  // only the Fast dispatch sends blocks > 2 MiB here
    static constexpr bool kBalanced = false;
    static constexpr int kMinMatch = kMatch8 ? 8 : 4;
    static constexpr bool kBackExtend = !kMatch8;
    int tb, sl, st, hb, ovh;
    bool quirk;
    __device__ __forceinline__ int table_bits() const { return tb; }
    __device__ __forceinline__ int skip_log() const { return sl; }
    __device__ __forceinline__ int step() const { return st; }
    __device__ __forceinline__ int max_fuse_lits() const { return kMatch8 ? 0 : 3; }
    __device__ __forceinline__ int max_fuse_lits2() const { return max_fuse_lits(); }
    __device__ __forceinline__ int max_fuse_lits3() const { return max_fuse_lits(); }
    __device__ __forceinline__ int s_limit(int n) const { return n - 17; }
    __device__ __forceinline__ int dst_limit(int n) const { return n - 17 - (n >> (kMatch8 ? 3 : 5)); }
    __device__ __forceinline__ int lit_overhead() const { return ovh; }
    __device__ __forceinline__ bool lit_quirk() const { return quirk; }
    __device__ __forceinline__ uint32_t hash(uint64_t u) const {
        if (kMatch8) return hash8(u, tb);
        return hb == 6 ? hash6(u, tb) : hb == 5 ? hash5(u, tb) : hash4(u, tb);
    }
    // encode_amd64.go:119-189 (encodeBlock) / :37-107 (encodeBlockFast) -> gen.go:57-76
    __device__ static __forceinline__ AsmClassParams for_len(int n) {
        AsmClassParams q;
        q.st = kMatch8 ? 4 : 3;
        q.hb = kMatch8 ? 8 : 6;
        q.ovh = 4;
        q.quirk = false;
        if (n > (2 << 20)) {
            q.tb = kMatch8 ? 14 : 15, q.sl = kMatch8 ? 5 : 6, q.st = 4;
        } else if (n > (512 << 10)) {
            q.tb = kMatch8 ? 13 : 15, q.sl = kMatch8 ? 5 : 6, q.st = 4;
        } else if (n > (64 << 10)) {
            q.tb = kMatch8 ? 13 : 14, q.sl = kMatch8 ? 5 : 6, q.st = 4;
        } else if (n > (16 << 10)) {
            q.tb = kMatch8 ? 12 : 13, q.sl = kMatch8 ? 4 : 5, q.quirk = true;   // maxLen == 64 KiB
        } else if (n > (4 << 10)) {
            q.tb = kMatch8 ? 11 : 12, q.sl = kMatch8 ? 4 : 5, q.hb = kMatch8 ? 8 : 5, q.ovh = 3;
        } else if (n > (1 << 10)) {
            q.tb = 10, q.sl = kMatch8 ? 4 : 5, q.hb = kMatch8 ? 8 : 4, q.ovh = 3;
        } else {
            q.tb = 9, q.sl = kMatch8 ? 3 : 4, q.hb = kMatch8 ? 8 : 4, q.ovh = 3;
        }
        return q;
    }
};

// LevelSuperFast: encode_l0.go:32-279 (large) and :281-522 (<= 64 KiB).  Same walk
// with an 8-byte minimum match, hash8, no backward extension, extension from +8.
template <bool kSmall>
struct L0Params {
    static constexpr bool kAsm = false;
    static constexpr bool kMayClamp = false;
    static constexpr bool kBalanced = false;
    static constexpr int kMinMatch = 8;
    static constexpr bool kBackExtend = false;  // encode_l0.go:164 `for false && ...`
    __device__ __forceinline__ int table_bits() const { return kSmall ? 12 : 13; }
    __device__ __forceinline__ int skip_log() const { return kSmall ? 4 : 5; }
    __device__ __forceinline__ int step() const { return kSmall ? 4 : 5; }
    __device__ __forceinline__ int max_fuse_lits() const { return kSmall ? kMaxCopy2Lits : kMaxCopy3Lits; }
    __device__ __forceinline__ int max_fuse_lits2() const { return max_fuse_lits(); }
    __device__ __forceinline__ int max_fuse_lits3() const { return max_fuse_lits(); }
    __device__ __forceinline__ int s_limit(int n) const { return n - kInputMargin; }
    __device__ __forceinline__ int dst_limit(int n) const { return kSmall ? n - (n >> 4) - 32 : n - (n >> 3) - 6; }
    __device__ __forceinline__ int lit_overhead() const { return 0; }
    __device__ __forceinline__ bool lit_quirk() const { return false; }
    __device__ __forceinline__ uint32_t hash(uint64_t u) const { return hash8(u, kSmall ? 12 : 13); }
};

// Backward extension, restating
//   for cand > 0 && s > floor && src[cand-1] == src[s-1] { cand--; s-- }
// with one byte pair per lane per round.  Returns the number of steps taken.
__device__ __noinline__ int extend_backward(const uint8_t *src, int cand, int s, int floor_, int lane) {
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
constexpr int kEncTagSlots = 1 << 15;  // slots per block (the largest table of any class)
constexpr int kSnapFwd = 24;  // forward bytes held in a slot

// One slot = one 32-byte DRAM sector, moved with Blackwell's 256-bit global
// accesses (LDG.E.ENL2.256 / STG.E.ENL2.256): a single request per probe, and a
// full-sector write per insert (no read-for-ownership of a half-written sector).
__device__ __forceinline__ void slot_load(const Slot *p, uint4 &a, uint4 &b) {
#ifndef MZ_SLOT_LD
// no L1 allocation: a slot is read once per probe and 4.3 GB of tables never fit anyway (-2 %)
#if MZ_ENC_L2_HINTS
#define MZ_SLOT_LD "ld.global.L1::no_allocate.L2::evict_first.v8.b32"
#else
#define MZ_SLOT_LD "ld.global.L1::no_allocate.v8.b32"
#endif
#endif
    asm volatile(MZ_SLOT_LD " {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
                 : "l"(p)
                 : "memory");
}
__device__ __forceinline__ void slot_store(Slot *p, const uint4 &a, const uint4 &b) {
#ifndef MZ_SLOT_ST
#if MZ_ENC_L2_HINTS
#define MZ_SLOT_ST "st.global.L2::evict_first.v8.b32"
#else
#define MZ_SLOT_ST "st.global.v8.b32"
#endif
#endif
    asm volatile(MZ_SLOT_ST " [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
                 "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
                 : "memory");
}

// ---- tag filter in front of the slots ----------------------------------------
// A random 32-byte slot read costs a whole 128-byte DRAM line, and only ~45 % of a
// window's probes can verify at all (profiles/r02_tag_filter.txt).  So every slot has a
// small TAG derived from its 4 verify bytes in a second, dense table.  A probe reads its
// tag first; when it differs from the tag of the lane's own 4 bytes the slot's bytes
// differ too, the probe CANNOT verify, and the DRAM line is never fetched.  A matching
// tag (every real match + 2^-bits of the rest) fetches the slot and verifies on the
// bytes as before, so every decision is unchanged.
// Invariant: tag[h] == tag_of(verify bytes of slot h), kept by one atomic XOR
// (old ^ new) whenever the slot is overwritten; untouched slots hold position 0, so
// the table starts as tag_of(src[0..4)) everywhere.
// Where the tags live decides what the filter is worth (B200, 4096 x 1 MiB, amd64 flavour):
//   no filter                              111.7 ms   420 GB read
//   4-bit tags in global memory (64 MiB)   132.2 ms   390 GB read  -- the tags do NOT stay in the
//       126 MB L2 against 450 GB of streaming slot lines, with or without evict_last hints or a
//       persisting set-aside: every tag read becomes one more DRAM line and one more dependent trip
//   2-bit tags in global memory (32 MiB)   123.9 ms   295 GB read
//   1-bit tags in SHARED memory (4 KiB per block, 28 blocks per SM)
//                                          107.0 ms   274 GB read  <- the default: no extra traffic,
//       no dependent L2 round trip; 2 bits per slot would need 224 KB per SM and do not fit.
// MZ_ENC_TAGS: 0 = no filter, 1 = tags in global memory / L2 (MZ_ENC_TAG_BITS wide), 2 = 1-bit tags in
// shared memory (4 KiB per block: no extra memory traffic and no dependent L2 round trip, half the power)
#ifndef MZ_ENC_TAGS
#define MZ_ENC_TAGS 2
#endif
#ifndef MZ_ENC_TAG_BITS
#define MZ_ENC_TAG_BITS 4
#endif
#if MZ_ENC_TAGS == 2
#undef MZ_ENC_TAG_BITS
#define MZ_ENC_TAG_BITS 1
#endif
#ifndef MZ_ENC_L2_HINTS
#define MZ_ENC_L2_HINTS 0  // tags: L2 evict_last; slots: L2 evict_first.  Measured: no help for the tags, -5 % for the slots
#endif
constexpr int kTagBits = MZ_ENC_TAG_BITS;
constexpr int kTagsPerWord = 32 / kTagBits;
constexpr uint32_t kTagMask = (1u << kTagBits) - 1u;
constexpr int kTagWordsPerWarp = kEncTagSlots / kTagsPerWord;
__device__ __forceinline__ uint32_t tag_of(uint32_t v) { return (v * 2654435761u) >> (32 - kTagBits); }
__device__ __forceinline__ uint32_t tag_word_index(uint32_t h) { return h / kTagsPerWord; }
__device__ __forceinline__ uint32_t tag_shift(uint32_t h) { return (h % kTagsPerWord) * kTagBits; }
__device__ __forceinline__ uint32_t tag_load(const uint32_t *p) {
    uint32_t v;
#if MZ_ENC_TAGS == 2
    v = *reinterpret_cast<const volatile uint32_t *>(p);
#elif MZ_ENC_L2_HINTS
    // L1 is bypassed (.cg): the tags are changed by atomics at L2
    asm volatile("{\n .reg .b64 pol;\n createpolicy.fractional.L2::evict_last.b64 pol, 1.0;\n"
                 " ld.global.cg.L2::cache_hint.u32 %0, [%1], pol;\n}"
                 : "=r"(v) : "l"(p) : "memory");
#else
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
#endif
    return v;
}
__device__ __forceinline__ void tag_xor(uint32_t *p, uint32_t v) {
#if MZ_ENC_TAGS == 2
    atomicXor(p, v);
#elif MZ_ENC_L2_HINTS
    asm volatile("{\n .reg .b64 pol;\n createpolicy.fractional.L2::evict_last.b64 pol, 1.0;\n"
                 " red.global.xor.L2::cache_hint.b32 [%0], %1, pol;\n}" ::"l"(p), "r"(v) : "memory");
#else
    asm volatile("red.global.xor.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#endif
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
    const int *gate;  // arrival gate (see gate_wait), nullptr when the source is resident
    int slice;
    int avail;  // bytes of this block known to have arrived

    // Loads 256-byte chunks until `want_end` is covered.  All lanes call.
    __device__ __forceinline__ void ensure(int want_end, int lane) {
        while (filled < want_end) {
            if (gate) {
                const int need = min(filled + kRingChunk + 16, n);
                if (need > avail) avail = gate_wait(gate, slice, need);
            }
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

// Optional walk counters for profiling builds (-DMZ_ENC_STATS, see profiles/enc_stats.py);
// compiled out of the product library.
#ifdef MZ_ENC_STATS
__device__ unsigned long long g_enc_stats[16];
#define ENC_STAT(i)                                         \
    do {                                                    \
        if (lane == 0) atomicAdd(&g_enc_stats[i], 1ull);    \
    } while (0)
#else
#define ENC_STAT(i) ((void)0)
#endif

// One warp per CTA: a block is encoded by one warp and the warps share nothing, so the CTA is only a
// container -- and with a single warp in it every shared-memory address below (ring, records, tags) is a
// compile-time constant.  With four warps per CTA the kernel carried the per-warp pointers in registers
// it did not have: 172 bytes of spills and ~36 instructions per batch recomputing shared-memory
// addresses (S2R, in front of the probe loads).  Same 72-register budget, no spills:
// 106.9 -> 92.5 ms (amd64 flavour), 108.8 -> 94.7 ms (Go) on 4096 x 1 MiB
// (profiles/r02_one_warp_per_cta_l1_l2.txt, r02_l1_one_warp_per_cta.txt; two warps per CTA change nothing).
#ifndef MZ_ENC_L1_WARPS
#define MZ_ENC_L1_WARPS 1
#endif
constexpr int kEncL1Warps = MZ_ENC_L1_WARPS;  // warps per CTA
#ifndef MZ_ENC_L1_MIN_CTAS
#define MZ_ENC_L1_MIN_CTAS (28 / MZ_ENC_L1_WARPS)  // 28 warps per SM x 148 SMs >= 4096 blocks in flight, 72 registers
#endif
constexpr int kEncL1SlotsPerWarp = 1 << 15;
constexpr size_t kEncL1WsBytesPerWarp = (size_t)kEncL1SlotsPerWarp * sizeof(Slot) + kTagWordsPerWarp * 4;  // 1 MiB of slots + the tags

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

// Cold path of a probe whose slot was written earlier in the same batch: the
// table then holds that insert, whose bytes are still in the ring.  Returns
// nzmask24(src[cand..], src[pos..]) | back-equal byte count (<= 4) << 24.
__device__ __noinline__ uint32_t probe_forwarded(const uint32_t *ring_mem, int pos, int cand) {
    SrcRing ring{const_cast<uint32_t *>(ring_mem), nullptr, 0, 0, nullptr, 0, 0};
    uint32_t A[7], B[7];
    ring.fetch28(pos - 4, A);
    ring.fetch28(cand - 4, B);
    const uint32_t x = A[0] ^ B[0];
    const uint32_t bb = x ? __clz(x) >> 3 : 4;
    return nzmask24(B + 1, A + 1) | bb << 24;
}

// ---- token writer ---------------------------------------------------------------
// The walk appends its matches (base, end, offset | kind << 24) to a small per-warp
// record ring in shared memory; whenever 32 are pending they are turned into tokens
// TOGETHER: lane = record.  Every lane builds its token bytes (literal header, copy /
// repeat / fused token, trailing repeat), one warp prefix sum gives the output
// positions, the reference's bail-out tests are evaluated per lane with the same
// values of d the serial code would see, and the bytes leave with a fixed number of
// predicated stores.  ~9 warp instructions per token instead of ~95 for a serial
// token-at-a-time writer, and the work moves out of the walk's dependency chain.
constexpr int kRecRing = 64;  // records per warp (>= 31 pending + 9 new per batch)


// Emits `cnt` (1..32) records held one per lane.  Returns false when a bail-out test fires.
template <class P>
__device__ __forceinline__ bool emit_group(const P prm, uint8_t *dst, const uint8_t *src, int &d, int &emitted,
                                           const int base, const int rk, const int end, const int cnt, const int lane,
                                           const int sLimit, const int dstLimit) {
    const bool valid = lane < cnt;
    const int prev_end = __shfl_up_sync(kFullMask, end, 1);
    const int ne = lane == 0 ? emitted : prev_end;  // literals start where the previous record ended
    const int kind = rk >> 24, rep = rk & 0xffffff;  // kind 3: literals + repeat, 0: (literals +) copy
    const int litLen = valid ? base - ne : 0;
    const int length = end - base;
    // how the literals travel: L1 / L0 encode_l1.go:190-206 (one limit for both fused forms),
    // L2 encode_l2.go:266-289 / gen.go:1803-1866 (<= 4 with a copy2, <= 3 with a copy3)
    bool sep = false, fused2 = false, fused3 = false;
    if (valid && litLen > 0) {
        if (kind || rep < kMinCopy2Offset) sep = true;
        else if (rep <= kMaxCopy2Offset) (litLen <= prm.max_fuse_lits2() ? fused2 : sep) = true;
        else (litLen <= prm.max_fuse_lits3() ? fused3 : sep) = true;
    }
    uint64_t lh = 0, tok = 0, post = 0;
    int n_lh = 0, n_tok = 0, n_post = 0;
    if (sep) {
        lh = tok_literal_hdr((uint32_t)litLen, &n_lh);
        if (prm.lit_quirk() && n_lh == 3) {  // gen.go:2193-2200
            n_lh = 4;
            lh = (lh & ~0xffull) | (31u << 3 | kTagLiteral);
        }
    }
    if (valid) {
        if (kind) {
            tok = tok_repeat((uint32_t)length, &n_tok);
        } else if (fused2) {
            tok = tok_copy2_fused((uint32_t)rep, (uint32_t)length, (uint32_t)litLen, &post, &n_post);
            n_tok = 3;
        } else if (fused3) {
            tok = tok_copy3((uint32_t)rep, (uint32_t)length, (uint32_t)litLen, &n_tok);
        } else {
            tok = tok_copy((uint32_t)rep, (uint32_t)length, &n_tok);
        }
    }
    const int total = n_lh + n_tok + n_post + litLen;
    int incl = total;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(kFullMask, incl, o);
        if (lane >= o) incl += t;
    }
    const int d0 = d + incl - total;  // d when the serial code reaches this record

    bool fail = false;
    if (valid) {
        const int ovh = prm.lit_overhead();
        if (kind) {  // :103 / encode_l2.go:147; Asm: gen.go:614-624,1490-1508 checkDst(litLen)
            fail = P::kAsm ? d0 + litLen + ovh >= dstLimit : d0 + litLen > dstLimit;
        } else {
            if (P::kBalanced) {  // before every match: encode_l2.go:229; Asm: gen.go:1720-1737
                if (P::kAsm ? d0 + litLen + ovh >= dstLimit : d0 + litLen > dstLimit) fail = true;
            } else {
                if (P::kAsm && d0 >= dstLimit) fail = true;  // gen.go:828 (and :1039 with the same d)
                if (sep && (P::kAsm ? d0 + litLen + ovh >= dstLimit : d0 + (end - ne) > dstLimit)) fail = true;  // :194
            }
            // after the copy (:229 / encode_l2.go:293-297; Asm: gen.go:955-975,1899-1918)
            if (end < sLimit && (P::kAsm ? d0 + total >= dstLimit : d0 + total > dstLimit)) fail = true;
        }
    }
    if (__any_sync(kFullMask, fail)) return false;

    uint8_t *o = dst + d0;
    const int pos_tok = sep ? n_lh + litLen : 0;
    const int pos_post = n_tok + litLen;  // fused2 only
#pragma unroll
    for (int k = 0; k < 4; k++)
        if (k < n_lh) o[k] = (uint8_t)(lh >> (8 * k));
#pragma unroll
    for (int k = 0; k < 7; k++)
        if (k < n_tok) o[pos_tok + k] = (uint8_t)(tok >> (8 * k));
#pragma unroll
    for (int k = 0; k < 4; k++)
        if (k < n_post) o[pos_post + k] = (uint8_t)(post >> (8 * k));
    if (fused2 || fused3) {  // <= 4 literals right behind the copy header
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (k < litLen) o[n_tok + k] = src[ne + k];
    }
    // separate literal runs are copied by the whole warp, one run after the other
    unsigned m = __ballot_sync(kFullMask, sep);
    while (m) {
        const int j = __ffs(m) - 1;
        m &= m - 1;
        const int s_ne = __shfl_sync(kFullMask, ne, j);
        const int s_len = __shfl_sync(kFullMask, litLen, j);
        const int s_pos = __shfl_sync(kFullMask, d0 + n_lh, j);
        copy_bytes(dst + s_pos, src + s_ne, s_len, lane);
    }
    d += __shfl_sync(kFullMask, incl, 31);
    emitted = __shfl_sync(kFullMask, end, cnt - 1);
    return true;
}

// Cold path of the Asm flavour's 8 MiB class: a far candidate is clamped to
// s - 2162685 (gen.go:466-490) and matched THERE; the slot snapshot is of another
// position, so compare straight from the source.  Same result layout as above.
__device__ __noinline__ uint32_t probe_direct(const uint8_t *src, int n, int pos, int cand) {
    uint32_t nz = 0;
    for (int k = kSnapFwd - 1; k >= 0; k--) {
        const uint8_t a = pos + k < n ? src[pos + k] : 0, b = cand + k < n ? src[cand + k] : 0;
        nz = nz << 1 | (a != b);
    }
    uint32_t bb = 0;
    while (bb < 4 && cand - 1 - (int)bb >= 0 && src[cand - 1 - bb] == src[pos - 1 - bb]) bb++;
    return nz | bb << 24;
}

// Encodes one block with one warp.  Returns bytes written or 0.
//
// One iteration of the outer loop is one BATCH = one DRAM round trip: lane L
// owns position wbase + L, loads its 28 source bytes from the ring, hashes
// them and fetches its table slot (position + snapshot of the bytes around it).
// After the round trip every lane knows whether ITS position would verify
// against the pre-batch table, how far the match runs (<= 24 bytes forward,
// <= 4 back) and whether it passes the repeat check.  The warp then replays the
// reference's control flow (search steps, back-to-back re-match probes, repeat
// checks) over these 32 answers with warp-uniform bit tests, consuming as many
// steps as fall inside the window (measured: 4.1 per round trip).  A match
// belongs to the lane of the position it was probed at; the common re-match hit
// is a bit test plus one shuffle because every lane has prepared "a hit at my
// position" beforehand.  At the end of the batch the owners append their
// records to a per-warp ring in shared memory, and whenever 32 records are
// pending they become tokens together (emit_group), in the shadow of the next
// batch's loads.  Finally the table inserts the replay performed are written
// back.  Inserts of the same batch that alias a later probe's slot are
// forwarded from the ring (`dup`, probe_forwarded).  The bail-out tests of the
// reference (dstLimit) are evaluated by the token writer with the same values
// of d, up to 63 records late; the result (0 = incompressible) is the same.
template <class P>
__device__ int encode_l1_block(const P prm, uint8_t *dst, const uint8_t *src, const int n, Slot *table, uint32_t *tags,
                               uint32_t *ring_mem, const int lane, const int *gate, const int slice) {
    const int sLimit = prm.s_limit(n);
    const int dstLimit = prm.dst_limit(n);
    // Asm flavour, 8 MiB class: far candidates are clamped (search) / rejected one byte earlier (re-match)
    constexpr int kClampDist = kMaxCopy3Offset - 2;  // gen.go:467-469: minPos = s - maxOffset + 2
    const bool clamp_far = P::kAsm && P::kMayClamp && n > (2 << 20);
    const int fill_limit = (n + 64 + kRingChunk - 1) & ~(kRingChunk - 1);
    constexpr uint32_t kMinMask = P::kMinMatch == 8 ? 0xffu : 0xfu;

    SrcRing ring{ring_mem, src, n, 0, gate, slice, 0};
    ring.ensure(min(2 * kRingChunk, fill_limit), lane);

    // Empty slots read as candidate 0 (encode_l1.go:52,86): position 0 and its bytes.
    {
        uint32_t w0[7];
        ring.fetch28(-4, w0);
        const uint4 ia = make_uint4(0, 0, w0[1], w0[2]);
        const uint4 ib = make_uint4(w0[3], w0[4], w0[5], w0[6]);
        const int slots = 1 << prm.table_bits();
        for (int i = lane; i < slots; i += 32) slot_store(table + i, ia, ib);
#if MZ_ENC_TAGS
        {
            uint32_t t0 = tag_of(w0[1]);
#pragma unroll
            for (int b = kTagBits; b < 32; b <<= 1) t0 |= t0 << b;
            const uint4 tv = make_uint4(t0, t0, t0, t0);
            for (int i = lane; i < slots / kTagsPerWord / 4; i += 32) reinterpret_cast<uint4 *>(tags)[i] = tv;
        }
#endif
        __syncwarp();
    }

    int nextEmit = 0;
    int s = 1;  // cursor: the re-match position, or the search position t
    int repeat = 1;
    int d = 0;
    bool rematch = false;  // the cursor is in the re-match loop (:222-265)
    bool done = false;     // reached emitRemainder

    // A match found in this batch belongs to the lane of the position it was probed at
    // (positions only grow within a batch, so lane order = stream order): the owner lane
    // holds [o_base, o_end) and o_rep = offset | kind << 24; `mmask` marks the owners.
    // (the literals of a match start where the previous one ended: emitted)
    int emitted = 0;  // nextEmit as the token writer sees it
    uint32_t *recs = ring_mem + kRingWords + kRingMirror;  // [3][kRecRing]: base, offset | kind << 24, end
    int r_head = 0, r_pending = 0;

    const unsigned below = (1u << lane) - 1u;
    const unsigned above = ~((2u << lane) - 1u);

    for (;;) {
        // ---------------- one round trip for the whole window ----------------
        const int wbase = rematch ? s - 2 : s;
        ENC_STAT(0);
        // When the skip distance is long (incompressible data) only the first search
        // step can fall in the window: do not fetch slots nobody will consume.
#ifndef MZ_ENC_WINDOW
#define MZ_ENC_WINDOW 32  // positions probed per round trip (each probe costs a 128-byte DRAM line)
#endif
        const int K = (!rematch && ((s - nextEmit) >> prm.skip_log()) >= 24) ? 8 : MZ_ENC_WINDOW;
        const int p = wbase + lane;
        const bool active = lane < K && !done;
        uint32_t W[7];  // src[p-4 .. p+24)
        uint32_t h = 0;
        uint4 ea = make_uint4(0, 0, 0, 0), eb = ea;
        uint32_t rep_lo = 0, rep_hi = 0;  // raw words around src[p - repeat]; shifted together after the shadow work
        unsigned rep_sh = 0;
        uint64_t far8 = 0;  // src[p-kClampDist-2 .. +8): the three clamped candidates of this position
        const bool rep_lane = active && p >= repeat;
        uint32_t tagw = 0, tagx = 0;  // the tag word of this lane's slot; old ^ own nibble (0 = the slot may verify)
        bool fetched = false;
        if (!done) {
            ring.seek(wbase - 8);
            ring.ensure(min(wbase + 320, fill_limit), lane);
            ring.fetch28(p - 4, W);
            h = prm.hash((uint64_t)W[2] << 32 | W[1]);
            // Positions this batch can never PROBE do not fetch their slot (a probe = a DRAM line):
            // in re-match mode the window starts at s-2, and s-2 / s-1 are only ever inserted
            // (:236-239); in search mode the step probes t, t+1, t+2 and whatever follows -- the next
            // step at nextS >= t + step, or the end of a match of >= 4 bytes -- lies at t + step or beyond.
            const bool never = rematch ? lane < 2 : (lane >= 3 && lane < prm.step());
#if MZ_ENC_TAGS
            if (active) tagw = tag_load(tags + tag_word_index(h));
#else
            if (active && !never) slot_load(table + h, ea, eb);
#endif
            if (rep_lane) {
                // (the funnel shift would wait for the load right here: keep the raw words)
                const uintptr_t ra = reinterpret_cast<uintptr_t>(src + p - repeat);
                const uint32_t *rw = reinterpret_cast<const uint32_t *>(ra & ~uintptr_t(3));
                rep_sh = (unsigned)(ra & 3) * 8;
                rep_lo = rw[0];
                if (rep_sh) rep_hi = rw[1];
            }
            if (clamp_far && active && p >= kClampDist) {
                const int lo = p - kClampDist - 2;
                far8 = lo >= 0 ? ldg_u64_unaligned(src + lo) : ldg_u64_unaligned(src) << (8 * -lo);
            }
#if MZ_ENC_TAGS
            // (a far candidate of the 8 MiB Asm class is compared at its CLAMPED position, which the
            // tag says nothing about: always fetch there)
            tagx = ((tagw >> tag_shift(h)) ^ tag_of(W[1])) & kTagMask;
            fetched = active && !never && (tagx == 0 || (clamp_far && p >= kClampDist));
            if (fetched) slot_load(table + h, ea, eb);
#else
            fetched = active && !never;
#endif
        }

        // ---------------- the loads are in flight: write queued tokens ----------------
        while (r_pending >= 32 || (done && r_pending > 0)) {
            const int cnt = min(r_pending, 32);
            const int r = (r_head + lane) & (kRecRing - 1);
            const int g_base = (int)recs[r], g_rk = (int)recs[kRecRing + r], g_end = (int)recs[2 * kRecRing + r];
            if (!emit_group(prm, dst, src, d, emitted, g_base, g_rk, g_end, cnt, lane, sLimit, dstLimit)) return 0;
            r_head = (r_head + cnt) & (kRecRing - 1);
            r_pending -= cnt;
        }
        __syncwarp();
        if (done) break;

        const unsigned same = __match_any_sync(kFullMask, h);
        const unsigned dup = __ballot_sync(kFullMask, (same & below) != 0);

        const int cand = (int)ea.x;
        uint32_t nz;
        {
            const uint32_t cd[6] = {ea.z, ea.w, eb.x, eb.y, eb.z, eb.w};
            nz = nzmask24(cd, W + 1);
        }
        {
            const uint32_t xb = ea.y ^ W[0];
            nz |= (xb ? __clz(xb) >> 3 : 4) << 24;  // back-equal bytes ride in bits 24..26
        }
#if MZ_ENC_TAGS
        if (!fetched) nz = 0xffffffu;  // cannot verify; never `cold`
#endif
        const bool eqm = active && (nz & kMinMask) == 0;
        const int dist = p - cand;
        // search probe j of a step sees minSrcPos = t - maxCopy3Offset (:83): dist <= max + j
        unsigned E0, E1, E2, ER;  // hit of lane L as search probe 0 / 1 / 2 of a step, as re-match probe
        if (clamp_far) {
            // probe j sees minPos = t - kClampDist: a candidate at distance >= kClampDist + j is
            // replaced by position p - j - kClampDist and its 4 / 8 bytes compared there (CMOV)
            const uint64_t cur = (uint64_t)W[2] << 32 | W[1];
            auto hitj = [&](int j) -> bool {
                if (dist < kClampDist + j) return eqm;
                const uint64_t cb = far8 >> (8 * (2 - j));
                // the 8-byte compare of the Fast classes needs 2 more bytes than far8 holds for j = 0, 1
                if (P::kMinMatch == 8)
                    return active && ldg_u64_unaligned(src + p - j - kClampDist) == cur;
                return active && (uint32_t)cb == (uint32_t)cur;
            };
            E0 = __ballot_sync(kFullMask, hitj(0));
            E1 = __ballot_sync(kFullMask, hitj(1));
            E2 = __ballot_sync(kFullMask, hitj(2));
            ER = __ballot_sync(kFullMask, eqm && dist < kMaxCopy3Offset);  // gen.go:1007-1021: candidate > base - maxOffset
        } else {
            E0 = __ballot_sync(kFullMask, eqm && dist <= kMaxCopy3Offset);
            E1 = E0, E2 = E0;
            if (!P::kAsm && n > kMaxCopy3Offset) {
                E1 = __ballot_sync(kFullMask, eqm && dist <= kMaxCopy3Offset + 1);
                E2 = __ballot_sync(kFullMask, eqm && dist <= kMaxCopy3Offset + 2);
            }
            ER = E0;
        }
        const uint32_t rep4 = __funnelshift_r(rep_lo, rep_hi, rep_sh);
        const unsigned Brep = __ballot_sync(kFullMask, rep_lane && rep4 == W[1]);

        // ---------------- replay the serial walk over the window ----------------
        unsigned ins = 0;     // lanes whose position was inserted; serial order = lane order
        int rep_snap = 0;     // repeat checks come from the last match's snapshot (Rnz, Rps)
        uint32_t Rnz = 0;
        int Rps = 0;
        // Every lane prepares the record of "a re-match hit at my position" up front: the match
        // starts at p (no literals, no backward extension in the re-match loop), runs for the
        // f_own equal snapshot bytes and has offset p - cand.  The replay then only needs the
        // end of such a match (one shuffle) to move on.  Lanes whose answer needs more than the
        // snapshot -- a forwarded probe, a match of 24+ bytes, the block's tail -- are `cold`
        // and take the general path, which overwrites the owner lane's record.
        uint32_t o_nz = nz & 0xffffffu;
        const int f_own = o_nz ? __ffs(o_nz) - 1 : kSnapFwd;
        int o_base = p, o_rep = dist, o_end = P::kAsm ? min(p + f_own, n) : p + f_own;
        const unsigned cold = dup | __ballot_sync(kFullMask, f_own == kSnapFwd || (!P::kAsm && p + f_own > n - 8));
        unsigned mmask = 0;   // owner lanes of this batch's matches
        int Rlane = -1;       // owner lane of the last fast re-match (its Rnz / repeat are fetched on demand)
        auto settle = [&]() {  // make Rnz / Rps / repeat current after fast re-matches
            if (Rlane >= 0) {
                Rnz = __shfl_sync(kFullMask, o_nz, Rlane);
                repeat = __shfl_sync(kFullMask, o_rep, Rlane);
                Rps = wbase + Rlane;
                rep_snap = 1;
                Rlane = -1;
            }
        };

        // Probe of lane L against the table as of `ins_at`: hit flag; when the slot was
        // written earlier in this batch the candidate is that insert (*fcand, *fnz).
        auto probe_cold = [&](int L, unsigned ins_at, bool fast_hit, int *fcand, uint32_t *fnz) -> bool {
            const unsigned sm = __shfl_sync(kFullMask, same, L);
            const unsigned f = sm & ins_at & ((1u << L) - 1u);
            *fcand = -1;
            if (f == 0) return fast_hit;
            ENC_STAT(6);
            *fcand = wbase + 31 - __clz(f);
            *fnz = probe_forwarded(ring_mem, wbase + L, *fcand);
            return (*fnz & kMinMask) == 0;
        };
        // End of a match that starts at `base`, is known equal for `known` bytes and whose
        // snapshot comparison found `f` equal bytes from the probe (24 = all of them).
        // Go: s = base + min match, then 8-byte chunks while s <= n-8 (:181-188)
        auto match_end = [&](int base, int known, int f, int offset) -> int {
            int e = base + known;
            if (P::kAsm) {  // matchLen to the end of the block (gen.go:859-887)
                if (f == kSnapFwd && e < n) {
                    ENC_STAT(8);
                    e = extend_forward_exact(src, e, e - offset, n, lane, gate, slice);
                }
                return min(e, n);
            }
            if (f == kSnapFwd || e > n - 8) {
                int q_stop = base + P::kMinMatch;
                if (q_stop <= n - 8) q_stop += (((n - 8 - q_stop) >> 3) + 1) << 3;
                if (f == kSnapFwd) {
                    const int sc = base + P::kMinMatch + 8 * ((known - P::kMinMatch) >> 3);
                    e = extend_forward8(src, sc, sc - offset, n - 8, lane, gate, slice);
                }
                e = min(e, q_stop);
            }
            return e;
        };

        for (;;) {
            if (rematch) {
                // ---- the re-match chain (:222-265): back-to-back copies, no literals ----
                for (;;) {
                    nextEmit = s;
                    if (s >= sLimit) {
                        done = true;
                        break;
                    }
                    const int L = s - wbase;
                    if (L >= K) {
                        ENC_STAT(3);
                        break;
                    }
                    ENC_STAT(1);
                    if (!((cold >> L) & 1u)) {  // fast: everything this step needs is already in lane L
                        ins |= 5u << (L - 2);
                        if (!((ER >> L) & 1u)) {
                            rematch = false;
                            s++;
                            break;
                        }
                        mmask |= 1u << L;
                        Rlane = L;
                        s = __shfl_sync(kFullMask, o_end, L);
                        continue;
                    }
                    settle();
                    int fcand = -1;
                    uint32_t fnz = 0;
                    bool hit = (ER >> L) & 1u;  // read before this step's inserts (:236-239)
                    if ((dup >> L) & 1u) hit = probe_cold(L, ins, hit, &fcand, &fnz);
                    ins |= 5u << (L - 2);
                    if (!hit) {
                        rematch = false;
                        s++;
                        break;
                    }
                    int mcand = fcand;
                    uint32_t mnz = fnz;
                    if (fcand < 0) {
                        mcand = __shfl_sync(kFullMask, cand, L);
                        mnz = __shfl_sync(kFullMask, nz, L);
                    }
                    mnz &= 0xffffffu;
                    const int f = mnz ? __ffs(mnz) - 1 : kSnapFwd;
                    repeat = s - mcand;
                    const int e = match_end(s, f, f, repeat);
                    if (lane == L) o_base = s, o_rep = repeat, o_end = e;
                    mmask |= 1u << L;
                    rep_snap = 1;
                    Rnz = mnz;
                    Rps = s;
                    s = e;
                }
                if (rematch) break;  // the chain left the window or reached the end
            }

            // ---- one search step at t (:70-160) ----
            settle();
            const int t = s;
            const int L = t - wbase;
            const int nextS = t + ((t - nextEmit) >> prm.skip_log()) + prm.step();  // :79
            if (P::kAsm ? nextS >= sLimit : nextS > sLimit) {  // :80; Asm: gen.go:458-461 JAE
                done = true;
                break;
            }
            if (L + 2 >= K) {
                ENC_STAT(4);
                break;
            }
            bool rhit;  // repeat check at t+1 (:94)
            if (rep_snap) {
                const int dl = t + 1 - Rps;
                if (dl > kSnapFwd - 4) {  // not covered by the snapshot: next batch
                    ENC_STAT(5);
                    break;
                }
                rhit = ((Rnz >> dl) & 0xfu) == 0;
            } else {
                rhit = (Brep >> (L + 1)) & 1u;
            }
            ENC_STAT(2);
            if (rhit) {
                ENC_STAT(7);
                ins |= 3u << L;
                int base = t + 1;
                // Go: both levels extend a repeat backwards; Asm: only with checkBack (gen.go:593)
                if (!P::kAsm || P::kBackExtend) base -= extend_backward(src, base - repeat, base, nextEmit, lane);
                s = P::kAsm ? extend_forward_exact(src, t + 5, t + 5 - repeat, n, lane, gate, slice)  // gen.go:632-660
                            : extend_forward8(src, t + 5, t + 5 - repeat, sLimit, lane, gate, slice);
                if (lane == L + 1) o_base = base, o_rep = repeat | 3 << 24, o_end = s;  // probed at t+1
                mmask |= 2u << L;
                nextEmit = s;
                if (s >= sLimit) {
                    done = true;
                    break;
                }
                continue;
            }
            const unsigned x0 = E0 >> L, x1 = E1 >> L, x2 = E2 >> L;
            bool h0 = x0 & 1u, h1 = x1 & 2u, h2 = x2 & 4u;
            int fcand = -1;
            uint32_t fnz = 0;
            if ((dup >> L) & 7u) {  // cold: re-evaluate the three probes in serial order
                int c;
                uint32_t z;
                h0 = probe_cold(L, ins, h0, &c, &z);
                if (h0) {
                    fcand = c;
                    fnz = z;
                } else {
                    h1 = probe_cold(L + 1, ins, h1, &c, &z);
                    if (h1) {
                        fcand = c;
                        fnz = z;
                    } else {
                        h2 = probe_cold(L + 2, ins | 3u << L, h2, &c, &z);  // read after t and t+1 (:150)
                        fcand = c;
                        fnz = z;
                    }
                }
            }
            int mps;  // position of the verified probe
            if (h0) {
                ins |= 3u << L;
                mps = t;
            } else {
                ins |= 7u << L;  // :152 / :157
                if (h1) {
                    mps = t + 1;
                } else if (h2) {
                    mps = t + 2;
                } else {
                    s = nextS;
                    continue;
                }
            }

            // ---- a verified candidate at mps: literals + copy ----
            int mcand = fcand;
            uint32_t mnz = fnz;
            if (fcand < 0) {
                mcand = __shfl_sync(kFullMask, cand, mps - wbase);
                mnz = __shfl_sync(kFullMask, nz, mps - wbase);
            }
            if (clamp_far && fcand < 0 && mps - mcand >= kClampDist + (mps - t)) {  // matched at the clamped position
                mcand = t - kClampDist;
                mnz = probe_direct(src, n, mps, mcand);
            }
            const int mbb = mnz >> 24;
            mnz &= 0xffffffu;
            const int f = mnz ? __ffs(mnz) - 1 : kSnapFwd;
            int base = mps, known = f;
            if (P::kBackExtend) {  // :169-172
                const int room = min(mcand, mps - nextEmit);
                int back = min(mbb, room);
                if (back == 4 && room > 4) {
                    ENC_STAT(9);
                    back += extend_backward(src, mcand - 4, mps - 4, nextEmit, lane);
                }
                base = mps - back;
                known = back + f;
            }
            repeat = mps - mcand;
            const int e = match_end(base, known, f, repeat);
            if (lane == mps - wbase) o_base = base, o_rep = repeat, o_end = e;
            mmask |= 1u << (mps - wbase);
            rep_snap = 1;
            Rnz = mnz;
            Rps = mps;
            s = e;
            rematch = true;
        }

        settle();  // `repeat` feeds the next batch's repeat-check loads
        if (mmask) {  // append this batch's matches to the record ring, in lane (= stream) order
            if ((mmask >> lane) & 1u) {
                const int w = (r_head + r_pending + __popc(mmask & below)) & (kRecRing - 1);
                recs[w] = (uint32_t)o_base;
                recs[kRecRing + w] = (uint32_t)o_rep;
                recs[2 * kRecRing + w] = (uint32_t)o_end;
            }
            r_pending += __popc(mmask);
        }
        // ---------------- write back the inserts the replay performed ----------------
        if (((ins >> lane) & 1u) && (same & ins & above) == 0) {  // a later insert on the same slot wins
            slot_store(table + h, make_uint4((uint32_t)p, W[0], W[1], W[2]), make_uint4(W[3], W[4], W[5], W[6]));
#if MZ_ENC_TAGS
            if (tagx) tag_xor(tags + tag_word_index(h), tagx << tag_shift(h));
#endif
        }
        __syncwarp();
    }

    // emitRemainder (encode_l1.go:268-282)
    if (nextEmit < n) {  // Asm: gen.go:1092-1119
        if (P::kAsm ? d + (n - nextEmit) + prm.lit_overhead() >= dstLimit : d + n - nextEmit > dstLimit) return 0;
        d += emit_literal(dst + d, src + nextEmit, n - nextEmit, lane, prm.lit_quirk());
    }
    return d;
}

// Persistent kernel: every warp pulls block indices from *counter and owns the
// workspace slice `tables + global_warp * 1 MiB`.
template <bool kSuperFast>
__global__ void __launch_bounds__(kEncL1Warps * 32, MZ_ENC_L1_MIN_CTAS)
encode_l1_kernel(int nblk, const uint8_t *__restrict__ src, const uint64_t *__restrict__ sbeg,
                 const uint64_t *__restrict__ send, uint8_t *dst, const uint64_t *__restrict__ dbeg,
                 uint32_t *__restrict__ out_len, int *counter, Slot *tables, const int *gate, int slice) {
    __shared__ uint32_t rings[kEncL1Warps][kRingWords + kRingMirror + 3 * kRecRing];
    const int lane = kEncL1Warps == 1 ? (int)threadIdx.x : lane_id();
    const int warp = kEncL1Warps == 1 ? 0 : (int)(threadIdx.x >> 5);
    const int gwarp = blockIdx.x * kEncL1Warps + warp;
    Slot *table = tables + (size_t)gwarp * kEncL1SlotsPerWarp;
    // the tag tables of all warps lie together behind the slots (dense: they are meant to stay in L2)
#if MZ_ENC_TAGS == 2
    __shared__ uint32_t tag_mem[kEncL1Warps][kTagWordsPerWarp];
    uint32_t *tags = tag_mem[warp];
#else
    uint32_t *tags = reinterpret_cast<uint32_t *>(tables + (size_t)gridDim.x * kEncL1Warps * kEncL1SlotsPerWarp) +
                     (size_t)gwarp * kTagWordsPerWarp;
#endif
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
                res = n <= 65536 ? encode_l1_block(L0Params<true>(), dp, sp, n, table, tags, rings[warp], lane, gate, slice)
                                 : encode_l1_block(L0Params<false>(), dp, sp, n, table, tags, rings[warp], lane, gate, slice);
            else
                res = n <= 65536 ? encode_l1_block(L1Params<true>(), dp, sp, n, table, tags, rings[warp], lane, gate, slice)
                                 : encode_l1_block(L1Params<false>(), dp, sp, n, table, tags, rings[warp], lane, gate, slice);
        }
        if (lane == 0) out_len[blk] = (uint32_t)res;
        __syncwarp();
    }
}

// The same persistent kernel for the Asm flavour (MZCU_FLAVOR_AMD64): size-class
// dispatch of encode_amd64.go:37-189.  A separate kernel so that the Go-flavour
// kernel's register allocation and code size are untouched.
template <bool kSuperFast>
__global__ void __launch_bounds__(kEncL1Warps * 32, MZ_ENC_L1_MIN_CTAS)
encode_l1_asm_kernel(int nblk, const uint8_t *__restrict__ src, const uint64_t *__restrict__ sbeg,
                     const uint64_t *__restrict__ send, uint8_t *dst, const uint64_t *__restrict__ dbeg,
                     uint32_t *__restrict__ out_len, int *counter, Slot *tables, const int *gate, int slice) {
    __shared__ uint32_t rings[kEncL1Warps][kRingWords + kRingMirror + 3 * kRecRing];
    const int lane = kEncL1Warps == 1 ? (int)threadIdx.x : lane_id();
    const int warp = kEncL1Warps == 1 ? 0 : (int)(threadIdx.x >> 5);
    const int gwarp = blockIdx.x * kEncL1Warps + warp;
    Slot *table = tables + (size_t)gwarp * kEncL1SlotsPerWarp;
    // the tag tables of all warps lie together behind the slots (dense: they are meant to stay in L2)
#if MZ_ENC_TAGS == 2
    __shared__ uint32_t tag_mem[kEncL1Warps][kTagWordsPerWarp];
    uint32_t *tags = tag_mem[warp];
#else
    uint32_t *tags = reinterpret_cast<uint32_t *>(tables + (size_t)gridDim.x * kEncL1Warps * kEncL1SlotsPerWarp) +
                     (size_t)gwarp * kTagWordsPerWarp;
#endif
    for (;;) {
        int blk = 0;
        if (lane == 0) blk = atomicAdd(counter, 1);
        blk = __shfl_sync(kFullMask, blk, 0);
        if (blk >= nblk) return;
        const uint8_t *sp = src + sbeg[blk];
        const int64_t n64 = (int64_t)(send[blk] - sbeg[blk]);
        uint8_t *dp = dst + dbeg[blk];
        int res = 0;
        // encode_amd64.go:106,187: the smallest class needs len > 32 (Fast) / > 16
        if (n64 > (kSuperFast ? 32 : kMinNonLiteralBlockSize) && n64 <= kMaxBlockSize) {
            const int n = (int)n64;
            if (!kSuperFast && n > (2 << 20))
                res = encode_l1_block(L1AsmBigParams<true>(), dp, sp, n, table, tags, rings[warp], lane, gate, slice);
            else if (!kSuperFast && n > (512 << 10))
                res = encode_l1_block(L1AsmBigParams<false>(), dp, sp, n, table, tags, rings[warp], lane, gate, slice);
            else if (kSuperFast && n > (64 << 10) && n <= (2 << 20))
                res = encode_l1_block(L0AsmBigParams(), dp, sp, n, table, tags, rings[warp], lane, gate, slice);
            else
                res = encode_l1_block(AsmClassParams<kSuperFast>::for_len(n), dp, sp, n, table, tags, rings[warp], lane,
                                      gate, slice);
        }
        if (lane == 0) out_len[blk] = (uint32_t)res;
        __syncwarp();
    }
}

}  // namespace mz
