// mz_encode_l1.cuh -- MinLZ LevelFastest block encoder (sm_100a).
//
// Replaces encodeBlock (reference asm_none.go:51-59 -> encode_l1.go:39-283
// encodeBlockGo and :285-524 encodeBlockGo64K; asm twin encodeBlockAsm*).
// Byte-identical to the pure-Go path: same hash (hash6/15 bit, hash5/13 bit
// for <= 64 KiB), same probe order, same skip rule, same emit rules.
//
// Mapping: one block per CTA, the match table in shared memory (128 KiB for
// the 15-bit table, so one CTA per SM), the hash walk kept warp-uniform in
// warp 0 with the 32 lanes sharing match extension and literal copies.
#pragma once

#include "mz_common.cuh"

namespace mz {

// Forward match extension, restating the Go loop
//   for s <= limit { if diff := load64(s)^load64(cand); diff != 0 { s += tz>>3; break }; s += 8; cand += 8 }
// with 32 lanes comparing 32 consecutive 8-byte words per round.
__device__ __forceinline__ int extend_forward8(const uint8_t *src, int s, int cand, int limit, int lane) {
    for (;;) {
        int pos = s + 8 * lane;
        bool past = pos > limit;
        uint64_t diff = 0;
        if (!past) diff = ldg_u64_unaligned(src + pos) ^ ldg_u64_unaligned(src + cand + 8 * lane);
        unsigned stop = __ballot_sync(kFullMask, past || diff != 0);
        if (stop == 0) {
            s += 256;
            cand += 256;
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

template <bool kSmall>
struct L1Params {
    static constexpr int kTableBits = kSmall ? 13 : 15;
    static constexpr int kSkipLog = kSmall ? 5 : 6;
    static constexpr int kMaxFuseLits = kSmall ? kMaxCopy2Lits : kMaxCopy3Lits;
    __device__ static __forceinline__ uint32_t hash(uint64_t u) {
        return kSmall ? hash5(u, kTableBits) : hash6(u, kTableBits);
    }
};

// Encodes one block with one warp.  `table` is (1 << kTableBits) zeroed u32
// in shared memory.  Returns bytes written or 0 (not compressible).
template <bool kSmall>
__device__ int encode_l1_block(uint8_t *dst, const uint8_t *src, int n, uint32_t *table, int lane) {
    using P = L1Params<kSmall>;
    const int sLimit = n - kInputMargin;
    const int dstLimit = n - (n >> 5) - 6;
    int nextEmit = 0;
    int s = 1;
    uint64_t cv = ldg_u64_unaligned(src + s);
    int repeat = 1;
    int d = 0;
    int candidate;

    for (;;) {
        candidate = 0;
        for (;;) {
            int nextS = s + ((s - nextEmit) >> P::kSkipLog) + 4;
            if (nextS > sLimit) goto emit_remainder;
            int minSrcPos = s - kMaxCopy3Offset;
            uint32_t hash0 = P::hash(cv);
            uint32_t hash1 = P::hash(cv >> 8);
            candidate = (int)table[hash0];
            int candidate2 = (int)table[hash1];
            __syncwarp();
            if (lane == 0) {
                table[hash0] = (uint32_t)s;
                table[hash1] = (uint32_t)(s + 1);
            }
            __syncwarp();
            uint32_t hash2 = P::hash(cv >> 16);

            if ((uint32_t)(cv >> 8) == ldg_u32_unaligned(src + s - repeat + 1)) {
                int base = s + 1;
                for (int i = base - repeat; base > nextEmit && i > 0 && src[i - 1] == src[base - 1];) {
                    i--;
                    base--;
                }
                if (d + (base - nextEmit) > dstLimit) return 0;
                d += emit_literal(dst + d, src + nextEmit, base - nextEmit, lane);
                int cand = s - repeat + 4 + 1;
                s += 4 + 1;
                s = extend_forward8(src, s, cand, sLimit, lane);
                d += emit_repeat(dst + d, s - base, lane);
                nextEmit = s;
                if (s >= sLimit) goto emit_remainder;
                cv = ldg_u64_unaligned(src + s);
                continue;
            }

            if (candidate >= minSrcPos && (uint32_t)cv == ldg_u32_unaligned(src + candidate)) break;
            candidate = (int)table[hash2];
            __syncwarp();
            if (lane == 0) table[hash2] = (uint32_t)(s + 2);
            __syncwarp();
            if (candidate2 >= minSrcPos && (uint32_t)(cv >> 8) == ldg_u32_unaligned(src + candidate2)) {
                candidate = candidate2;
                s++;
                break;
            }
            if (candidate >= minSrcPos && (uint32_t)(cv >> 16) == ldg_u32_unaligned(src + candidate)) {
                s += 2;
                break;
            }
            cv = ldg_u64_unaligned(src + nextS);
            s = nextS;
        }

        while (candidate > 0 && s > nextEmit && src[candidate - 1] == src[s - 1]) {
            candidate--;
            s--;
        }
        int base = s;
        repeat = base - candidate;
        s += 4;
        candidate += 4;
        s = extend_forward8(src, s, candidate, n - 8, lane);
        int length = s - base;
        if (nextEmit != base) {
            if (base - nextEmit > P::kMaxFuseLits || repeat < kMinCopy2Offset) {
                if (d + (s - nextEmit) > dstLimit) return 0;
                d += emit_literal(dst + d, src + nextEmit, base - nextEmit, lane);
                d += emit_copy(dst + d, repeat, length, lane);
            } else if (repeat <= kMaxCopy2Offset) {
                d += emit_copy_lits2(dst + d, src + nextEmit, base - nextEmit, repeat, length, lane);
            } else {
                d += emit_copy_lits3(dst + d, src + nextEmit, base - nextEmit, repeat, length, lane);
            }
        } else {
            d += emit_copy(dst + d, repeat, length, lane);
        }

        for (;;) {
            nextEmit = s;
            if (s >= sLimit) goto emit_remainder;
            uint64_t x = ldg_u64_unaligned(src + s - 2);
            if (d > dstLimit) return 0;
            uint32_t m2Hash = P::hash(x);
            x >>= 16;
            uint32_t currHash = P::hash(x);
            candidate = (int)table[currHash];
            __syncwarp();
            if (lane == 0) {
                table[m2Hash] = (uint32_t)(s - 2);
                table[currHash] = (uint32_t)s;
            }
            __syncwarp();
            if (s - candidate > kMaxCopy3Offset || (uint32_t)x != ldg_u32_unaligned(src + candidate)) {
                cv = ldg_u64_unaligned(src + s + 1);
                s++;
                break;
            }
            repeat = s - candidate;
            base = s;
            s += 4;
            candidate += 4;
            s = extend_forward8(src, s, candidate, n - 8, lane);
            d += emit_copy(dst + d, repeat, s - base, lane);
        }
    }

emit_remainder:
    if (nextEmit < n) {
        if (d + n - nextEmit > dstLimit) return 0;
        d += emit_literal(dst + d, src + nextEmit, n - nextEmit, lane);
    }
    return d;
}

// Persistent kernel: CTAs pull block indices from *counter.
__global__ void __launch_bounds__(32)
encode_l1_kernel(int nblk, const uint8_t *__restrict__ src, const uint64_t *__restrict__ sbeg,
                 const uint64_t *__restrict__ send, uint8_t *dst, const uint64_t *__restrict__ dbeg,
                 uint32_t *__restrict__ out_len, int *counter) {
    extern __shared__ __align__(16) uint32_t table[];
    const int lane = lane_id();
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
            const bool small = n <= 65536;
            const int words = small ? (1 << 13) : (1 << 15);
            uint4 *t4 = reinterpret_cast<uint4 *>(table);
            for (int i = lane; i < words / 4; i += 32) t4[i] = make_uint4(0, 0, 0, 0);
            __syncwarp();
            res = small ? encode_l1_block<true>(dp, sp, n, table, lane) : encode_l1_block<false>(dp, sp, n, table, lane);
        }
        if (lane == 0) out_len[blk] = (uint32_t)res;
        __syncwarp();
    }
}

}  // namespace mz
