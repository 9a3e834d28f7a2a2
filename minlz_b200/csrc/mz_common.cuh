// mz_common.cuh -- shared device helpers for the MinLZ sm_100a kernels.
//
// Format constants follow the reference (encode.go:30-58, minlz.go:68-75).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace mz {

constexpr int kMaxBlockSize = 8 << 20;
constexpr int kMaxCopy1Offset = 1024;
constexpr int kMinCopy2Offset = 64;
constexpr int kMaxCopy2Offset = 64 + 65535;
constexpr int kCopy2LitMaxLen = 7 + 4;
constexpr int kMaxCopy2Lits = 4;
constexpr int kMaxCopy3Lits = 3;
constexpr int kMinCopy3Offset = 65536;
constexpr int kMaxCopy3Offset = (2 << 20) + 65535;
constexpr int kInputMargin = 8;
constexpr int kMinNonLiteralBlockSize = 16;

constexpr uint32_t kTagLiteral = 0, kTagRepeat = 4, kTagCopy1 = 1, kTagCopy2 = 2, kTagCopy3 = 7, kTagCopy2Fused = 3;

constexpr unsigned kFullMask = 0xffffffffu;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// ---- unaligned little-endian loads from global memory ----------------------
// Built from aligned 32-bit words + funnel shift.  `p` must point into a
// buffer whose containing aligned words are readable (true for any cudaMalloc
// range: allocations are 256 B aligned and padded); the word after the one
// holding the last requested byte is never touched.

__device__ __forceinline__ uint32_t ldg_u32_unaligned(const uint8_t *p) {
    uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t *w = reinterpret_cast<const uint32_t *>(a & ~uintptr_t(3));
    unsigned sh = (unsigned)(a & 3) * 8;
    uint32_t lo = w[0];
    if (sh == 0) return lo;
    uint32_t hi = w[1];
    return __funnelshift_r(lo, hi, sh);
}

__device__ __forceinline__ uint64_t ldg_u64_unaligned(const uint8_t *p) {
    uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t *w = reinterpret_cast<const uint32_t *>(a & ~uintptr_t(3));
    unsigned sh = (unsigned)(a & 3) * 8;
    uint32_t w0 = w[0], w1 = w[1];
    if (sh == 0) return (uint64_t)w1 << 32 | w0;
    uint32_t w2 = w[2];
    return (uint64_t)__funnelshift_r(w1, w2, sh) << 32 | __funnelshift_r(w0, w1, sh);
}

// 8 stream bytes at s (zero filled past slen); never reads a word that holds
// no byte of [0, slen).
__device__ __forceinline__ uint64_t ldg_window(const uint8_t *sp, int64_t s, int64_t slen) {
    if (s + 8 <= slen) return ldg_u64_unaligned(sp + s);
    uint64_t w = 0;
    for (int i = 0; i < 8 && s + i < slen; i++) w |= (uint64_t)sp[s + i] << (8 * i);
    return w;
}

// hashes: encode_l1.go:26-29, encode_l2.go:25-49
__device__ __forceinline__ uint32_t hash4(uint64_t u, int h) { return ((uint32_t)u * 2654435761u) >> (32 - h); }
__device__ __forceinline__ uint32_t hash5(uint64_t u, int h) {
    return (uint32_t)(((u << 24) * 889523592379ull) >> (64 - h));
}
__device__ __forceinline__ uint32_t hash6(uint64_t u, int h) {
    return (uint32_t)(((u << 16) * 227718039650203ull) >> (64 - h));
}
__device__ __forceinline__ uint32_t hash8(uint64_t u, int h) { return (uint32_t)((u * 0xcf1bbcdcb7a56463ull) >> (64 - h)); }
__device__ __forceinline__ uint32_t hash7(uint64_t u, int h) {
    return (uint32_t)(((u << 8) * 58295818150454627ull) >> (64 - h));
}

// ---- token builders ---------------------------------------------------------
// Each returns the token bytes packed little-endian in a uint64 and the byte
// count in *n (<= 7).  They restate the reference emitters byte for byte:
// emitRepeat asm_none.go:125-156, emitCopy :207-278, encodeCopy2
// encode.go:247-282, encodeCopy3 asm_none.go:160-200, emitLiteral header
// asm_none.go:84-122.

__device__ __forceinline__ uint64_t tok_literal_hdr(uint32_t len, int *n) {
    uint32_t v = len - 1;
    if (v < 29) {
        *n = 1;
        return v << 3 | kTagLiteral;
    }
    if (v < (1u << 8) + 29) {
        *n = 2;
        return (29u << 3 | kTagLiteral) | (uint64_t)(v - 29) << 8;
    }
    if (v < (1u << 16) + 29) {
        *n = 3;
        return (30u << 3 | kTagLiteral) | (uint64_t)(v - 29) << 8;
    }
    *n = 4;
    return (31u << 3 | kTagLiteral) | (uint64_t)(v - 29) << 8;
}

__device__ __forceinline__ uint64_t tok_repeat(uint32_t length, int *n) {
    if (length < 30) {
        *n = 1;
        return (length - 1) << 3 | kTagRepeat;
    }
    length -= 30;
    if (length < 256) {
        *n = 2;
        return (29u << 3 | kTagRepeat) | (uint64_t)length << 8;
    }
    if (length < 65536) {
        *n = 3;
        return (30u << 3 | kTagRepeat) | (uint64_t)length << 8;
    }
    *n = 4;
    return (31u << 3 | kTagRepeat) | (uint64_t)length << 8;
}

__device__ __forceinline__ uint64_t tok_copy3(uint32_t offset, uint32_t length, uint32_t lits, int *n) {
    length -= 4;
    uint64_t enc = (uint64_t)((offset - 65536) << 11 | kTagCopy3 | lits << 3);
    if (length <= 60) {
        *n = 4;
        return enc | length << 5;
    }
    length -= 60;
    if (length < 256) {
        *n = 5;
        return enc | 61u << 5 | (uint64_t)length << 32;
    }
    if (length < 65536) {
        *n = 6;
        return enc | 62u << 5 | (uint64_t)length << 32;
    }
    *n = 7;
    return enc | 63u << 5 | (uint64_t)length << 32;
}

__device__ __forceinline__ uint64_t tok_copy2(uint32_t offset, uint32_t length, int *n) {
    length -= 4;
    uint64_t off = (uint64_t)(offset - kMinCopy2Offset) << 8;
    if (length <= 60) {
        *n = 3;
        return off | length << 2 | kTagCopy2;
    }
    length -= 60;
    if (length < 256) {
        *n = 4;
        return off | (61u << 2 | kTagCopy2) | (uint64_t)length << 24;
    }
    if (length < 65536) {
        *n = 5;
        return off | (62u << 2 | kTagCopy2) | (uint64_t)length << 24;
    }
    *n = 6;
    return off | (63u << 2 | kTagCopy2) | (uint64_t)length << 24;
}

// emitCopy: may expand to copy1 + repeat (<= 2 + 4 bytes).
__device__ __forceinline__ uint64_t tok_copy(uint32_t offset, uint32_t length, int *n) {
    if (offset > (uint32_t)kMaxCopy2Offset) return tok_copy3(offset, length, 0, n);
    if (offset <= (uint32_t)kMaxCopy1Offset) {
        uint32_t o = (offset - 1) << 6;
        if (length < 15 + 4) {
            *n = 2;
            return (o | (length - 4) << 2 | kTagCopy1) & 0xffff;
        }
        if (length < 256 + 18) {
            *n = 3;
            return ((o | 15u << 2 | kTagCopy1) & 0xffff) | (uint64_t)(length - 18) << 16;
        }
        int rn;
        uint64_t r = tok_repeat(length - 18, &rn);
        *n = 2 + rn;
        return ((o | 14u << 2 | kTagCopy1) & 0xffff) | r << 16;
    }
    return tok_copy2(offset, length, n);
}

// emitCopyLits2 header (3 bytes); literals follow, then an optional repeat for
// length > 11 (returned through *rep / *rn).  asm_none.go:284-308
__device__ __forceinline__ uint32_t tok_copy2_fused(uint32_t offset, uint32_t length, uint32_t nlits, uint64_t *rep,
                                                    int *rn) {
    uint32_t off = (offset - kMinCopy2Offset) << 8;
    length -= 4;
    const uint32_t maxraw = kCopy2LitMaxLen - 4;
    *rn = 0;
    *rep = 0;
    if (length > maxraw) {
        *rep = tok_repeat(length - maxraw, rn);
        length = maxraw;
    }
    return off | kTagCopy2Fused | length << 5 | (nlits - 1) << 3;
}

}  // namespace mz
