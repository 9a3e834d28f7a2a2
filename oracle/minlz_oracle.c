/*
 * minlz_oracle.c -- CPU restatement of the MinLZ block codec: the pure-Go path
 * and the amd64-assembly flavour of the encoders.
 *
 * TEST INFRASTRUCTURE ONLY: see minlz_oracle.h.  Every function cites the
 * reference file:line it follows (paths relative to the minio/minlz tree).
 * Pinned against the reference's own assembly run through oracle/_ref
 * (tests/test_ref_asm.py) and against the reference's golden vectors
 * (tests/test_oracle_golden.py).
 */
#include "minlz_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#ifdef MZO_STATS
uint64_t mzo_stats[16];
#define STAT(i) (mzo_stats[i]++)
#else
#define STAT(i) ((void)0)
#endif

/* ---- encode.go:30-58 constants ---------------------------------------- */
enum {
    kMaxCopy1Offset = 1024,
    kMinCopy2Offset = 64,
    kMaxCopy2Offset = 64 + 65535,
    kCopy2LitMaxLen = 7 + 4,
    kMaxCopy2Lits = 4,
    kMaxCopy3Lits = 3,
    kMinCopy3Offset = 65536,
    kMaxCopy3Offset = (2 << 20) + 65535,
    kInputMargin = 8,            /* encode.go:216 */
    kMinNonLiteralBlockSize = 16 /* encode.go:220 */
};

/* minlz.go:68-75 tags */
enum { TAG_LITERAL = 0, TAG_REPEAT = 4, TAG_COPY1 = 1, TAG_COPY2 = 2, TAG_COPY3 = 7, TAG_COPY2_FUSED = 3 };

/* ---- unsafe_enabled.go:26-60 little-endian unaligned access ------------ */
static inline uint16_t ld16(const uint8_t *b, int64_t i) { uint16_t v; memcpy(&v, b + i, 2); return v; }
static inline uint32_t ld32(const uint8_t *b, int64_t i) { uint32_t v; memcpy(&v, b + i, 4); return v; }
static inline uint64_t ld64(const uint8_t *b, int64_t i) { uint64_t v; memcpy(&v, b + i, 8); return v; }
static inline void st16(uint8_t *b, int64_t i, uint16_t v) { memcpy(b + i, &v, 2); }
static inline void st32(uint8_t *b, int64_t i, uint32_t v) { memcpy(b + i, &v, 4); }

/* ---- hashes: encode_l1.go:26-29, encode_l2.go:25-49 -------------------- */
static inline uint32_t hash4(uint64_t u, int h) { return ((uint32_t)u * 2654435761u) >> (32 - h); }
static inline uint32_t hash5(uint64_t u, int h) { return (uint32_t)(((u << 24) * 889523592379ull) >> (64 - h)); }
static inline uint32_t hash6(uint64_t u, int h) { return (uint32_t)(((u << 16) * 227718039650203ull) >> (64 - h)); }
static inline uint32_t hash7(uint64_t u, int h) { return (uint32_t)(((u << 8) * 58295818150454627ull) >> (64 - h)); }
static inline uint32_t hash8(uint64_t u, int h) { return (uint32_t)((u * 0xcf1bbcdcb7a56463ull) >> (64 - h)); }

static inline uint32_t hashN(uint64_t u, int h, int nbytes) {
    switch (nbytes) {
    case 4: return hash4(u, h);
    case 5: return hash5(u, h);
    case 6: return hash6(u, h);
    default: return hash7(u, h);
    }
}

/* encode.go:234-244 MaxEncodedLen */
int64_t mzo_max_encoded_len(int64_t src_len) {
    if (src_len < 0 || (uint64_t)src_len > MZO_MAX_BLOCK_SIZE) return -1;
    if (src_len == 0) return 1;
    return src_len + 2;
}

/* ---- emitters ---------------------------------------------------------- */

/* asm_none.go:84-122 emitLiteral */
int mzo_emit_literal(uint8_t *dst, const uint8_t *lit, size_t len) {
    if (len == 0) return 0;
    uint32_t n = (uint32_t)(len - 1);
    int i;
    if (n < 29) {
        dst[0] = (uint8_t)(n << 3) | TAG_LITERAL;
        i = 1;
    } else if (n < (1u << 8) + 29) {
        dst[1] = (uint8_t)(n - 29);
        dst[0] = 29 << 3 | TAG_LITERAL;
        i = 2;
    } else if (n < (1u << 16) + 29) {
        n -= 29;
        dst[2] = (uint8_t)(n >> 8);
        dst[1] = (uint8_t)n;
        dst[0] = 30 << 3 | TAG_LITERAL;
        i = 3;
    } else {
        n -= 29;
        dst[3] = (uint8_t)(n >> 16);
        dst[2] = (uint8_t)(n >> 8);
        dst[1] = (uint8_t)n;
        dst[0] = 31 << 3 | TAG_LITERAL;
        i = 4;
    }
    memcpy(dst + i, lit, len);
    return i + (int)len;
}

/* asm_none.go:125-156 emitRepeat */
int mzo_emit_repeat(uint8_t *dst, int length) {
    if (length < 30) {
        dst[0] = (uint8_t)((length - 1) << 3) | TAG_REPEAT;
        return 1;
    }
    length -= 30;
    if (length < 256) {
        dst[1] = (uint8_t)length;
        dst[0] = 29 << 3 | TAG_REPEAT;
        return 2;
    }
    if (length < 65536) {
        dst[2] = (uint8_t)(length >> 8);
        dst[1] = (uint8_t)length;
        dst[0] = 30 << 3 | TAG_REPEAT;
        return 3;
    }
    dst[3] = (uint8_t)(length >> 16);
    dst[2] = (uint8_t)(length >> 8);
    dst[1] = (uint8_t)length;
    dst[0] = 31 << 3 | TAG_REPEAT;
    return 4;
}

/* asm_none.go:160-200 encodeCopy3 (also the expanded copy in emitCopy :221-258) */
static int encode_copy3(uint8_t *dst, int offset, int length, int lits) {
    length -= 4;
    uint32_t enc = (uint32_t)(offset - 65536) << 11 | TAG_COPY3 | (uint32_t)(lits << 3);
    if (length <= 60) {
        enc |= (uint32_t)(length << 5);
        st32(dst, 0, enc);
        return 4;
    }
    length -= 60;
    if (length < 256) {
        dst[4] = (uint8_t)length;
        enc |= 61 << 5;
        st32(dst, 0, enc);
        return 5;
    }
    if (length < 65536) {
        enc |= 62 << 5;
        dst[5] = (uint8_t)(length >> 8);
        dst[4] = (uint8_t)length;
        st32(dst, 0, enc);
        return 6;
    }
    enc |= 63 << 5;
    dst[6] = (uint8_t)(length >> 16);
    dst[5] = (uint8_t)(length >> 8);
    dst[4] = (uint8_t)length;
    st32(dst, 0, enc);
    return 7;
}

/* encode.go:247-282 encodeCopy2 */
static int encode_copy2(uint8_t *dst, int offset, int length) {
    length -= 4;
    offset -= kMinCopy2Offset;
    st16(dst, 1, (uint16_t)offset);
    if (length <= 60) {
        dst[0] = (uint8_t)(length << 2) | TAG_COPY2;
        return 3;
    }
    length -= 60;
    if (length < 256) {
        dst[3] = (uint8_t)length;
        dst[0] = 61 << 2 | TAG_COPY2;
        return 4;
    }
    if (length < 65536) {
        dst[4] = (uint8_t)(length >> 8);
        dst[3] = (uint8_t)length;
        dst[0] = 62 << 2 | TAG_COPY2;
        return 5;
    }
    dst[5] = (uint8_t)(length >> 16);
    dst[4] = (uint8_t)(length >> 8);
    dst[3] = (uint8_t)length;
    dst[0] = 63 << 2 | TAG_COPY2;
    return 6;
}

/* asm_none.go:207-278 emitCopy */
int mzo_emit_copy(uint8_t *dst, int offset, int length) {
    if (offset > kMaxCopy2Offset) return encode_copy3(dst, offset, length, 0);
    if (offset <= kMaxCopy1Offset) {
        offset--;
        if (length < 15 + 4) {
            st16(dst, 0, (uint16_t)(offset << 6) | (uint16_t)((length - 4) << 2) | TAG_COPY1);
            return 2;
        }
        if (length < 256 + 18) {
            st16(dst, 0, (uint16_t)(offset << 6) | (15 << 2 | TAG_COPY1));
            dst[2] = (uint8_t)(length - 18);
            return 3;
        }
        /* copy1 of 18 bytes, rest as a repeat */
        st16(dst, 0, (uint16_t)(offset << 6) | (14 << 2) | TAG_COPY1);
        return 2 + mzo_emit_repeat(dst + 2, length - 18);
    }
    return encode_copy2(dst, offset, length);
}

/* asm_none.go:284-308 emitCopyLits2 */
int mzo_emit_copy_lits2(uint8_t *dst, const uint8_t *lits, int nlits, int offset, int length) {
    offset -= kMinCopy2Offset;
    length -= 4;
    const int maxraw = kCopy2LitMaxLen - 4;
    st16(dst, 1, (uint16_t)offset);
    if (length > maxraw) {
        dst[0] = TAG_COPY2_FUSED | (uint8_t)(maxraw << 5) | (uint8_t)((nlits - 1) << 3);
        memcpy(dst + 3, lits, (size_t)nlits);
        int n = nlits + 3;
        return n + mzo_emit_repeat(dst + n, length - maxraw);
    }
    dst[0] = TAG_COPY2_FUSED | (uint8_t)(length << 5) | (uint8_t)((nlits - 1) << 3);
    memcpy(dst + 3, lits, (size_t)nlits);
    return nlits + 3;
}

/* asm_none.go:313-323 emitCopyLits3 */
int mzo_emit_copy_lits3(uint8_t *dst, const uint8_t *lits, int nlits, int offset, int length) {
    int n = encode_copy3(dst, offset, length, nlits);
    memcpy(dst + n, lits, (size_t)nlits);
    return n + nlits;
}

/* ---- match extension helpers ------------------------------------------- */

/* The 8-bytes-at-a-time extension loop "for s <= limit" used by L1
 * (encode_l1.go:115-122 with limit=sLimit, :181-188/:251-258 with limit=len-8). */
static inline int extend8(const uint8_t *src, int s, int cand, int limit) {
    while (s <= limit) {
        uint64_t diff = ld64(src, s) ^ ld64(src, cand);
        if (diff != 0) {
            s += __builtin_ctzll(diff) >> 3;
            break;
        }
        s += 8;
        cand += 8;
    }
    return s;
}

/* L2 extension with byte tail (encode_l2.go:160-175, :239-254). */
static inline int extend_tail(const uint8_t *src, int n, int s, int cand) {
    while (s < n) {
        if (n - s < 8) {
            if (src[s] == src[cand]) {
                s++;
                cand++;
                continue;
            }
            break;
        }
        uint64_t diff = ld64(src, s) ^ ld64(src, cand);
        if (diff != 0) {
            s += __builtin_ctzll(diff) >> 3;
            break;
        }
        s += 8;
        cand += 8;
    }
    return s;
}

/* ---- L1: encode_l1.go:39-283 (tableBits 15, hash6, skipLog 6, fuse<=3) and
 *          encode_l1.go:285-524 (tableBits 13, hash5, skipLog 5, fuse<=4).
 * The two Go functions differ only in those parameters, in the width of the
 * table entries (positions always fit) and in the offset guards, which can
 * never fire for len <= 64 KiB; one parameterised body restates both. ------ */
static int64_t encode_l1(uint8_t *dst, const uint8_t *src, int n, int tableBits, int hashBytes,
                         int skipLog, int maxFuseLits) {
    uint32_t *table = (uint32_t *)calloc((size_t)1 << tableBits, sizeof(uint32_t)); /* :52 */
    if (!table) return 0;
    const int sLimit = n - kInputMargin;       /* :57 */
    const int dstLimit = n - (n >> 5) - 6;     /* :60 */
    int nextEmit = 0;                          /* :63 */
    int s = 1;                                 /* :67 */
    uint64_t cv = ld64(src, s);                /* :68 */
    int repeat = 1;                            /* :71 */
    int d = 0;
    int candidate;

    for (;;) {
        candidate = 0;
        for (;;) {
            STAT(0);
            int nextS = s + ((s - nextEmit) >> skipLog) + 4; /* :79 */
            if (nextS > sLimit) goto emit_remainder;          /* :80-82 */
            int minSrcPos = s - kMaxCopy3Offset;              /* :83 */
            uint32_t hash0 = hashN(cv, tableBits, hashBytes);
            uint32_t hash1 = hashN(cv >> 8, tableBits, hashBytes);
            candidate = (int)table[hash0];
            int candidate2 = (int)table[hash1];
            table[hash0] = (uint32_t)s;
            table[hash1] = (uint32_t)(s + 1);
            uint32_t hash2 = hashN(cv >> 16, tableBits, hashBytes); /* :90 */

            /* repeat check at s+1, :94-145 */
            if ((uint32_t)(cv >> 8) == ld32(src, s - repeat + 1)) {
                STAT(1);
                int base = s + 1;
                for (int i = base - repeat; base > nextEmit && i > 0 && src[i - 1] == src[base - 1];) {
                    i--;
                    base--;
                }
                if (d + (base - nextEmit) > dstLimit) { /* :103 */
                    free(table);
                    return 0;
                }
                d += mzo_emit_literal(dst + d, src + nextEmit, (size_t)(base - nextEmit));
                int cand = s - repeat + 4 + 1; /* :113 */
                s += 4 + 1;
                s = extend8(src, s, cand, sLimit); /* :115-122 */
                d += mzo_emit_repeat(dst + d, s - base);
                nextEmit = s;
                if (s >= sLimit) goto emit_remainder; /* :139 */
                cv = ld64(src, s);
                continue;
            }

            if (candidate >= minSrcPos && (uint32_t)cv == ld32(src, candidate)) { STAT(2); break; } /* :147 */
            candidate = (int)table[hash2];                                               /* :150 */
            if (candidate2 >= minSrcPos && (uint32_t)(cv >> 8) == ld32(src, candidate2)) {
                table[hash2] = (uint32_t)(s + 2);
                candidate = candidate2;
                s++;
                STAT(3);
                break;
            }
            table[hash2] = (uint32_t)(s + 2); /* :157 */
            if (candidate >= minSrcPos && (uint32_t)(cv >> 16) == ld32(src, candidate)) {
                s += 2;
                STAT(4);
                break;
            }
            STAT(5);
            cv = ld64(src, nextS); /* :163 */
            s = nextS;
        }

        /* extend backwards, :169-172 */
#ifdef MZO_STATS
        int probe_pos = s;
#endif
        while (candidate > 0 && s > nextEmit && src[candidate - 1] == src[s - 1]) {
            candidate--;
            s--;
        }
#ifdef MZO_STATS
        if (probe_pos - s > 4) STAT(8); /* backward extension beyond a 4-byte snapshot */
#endif
        int base = s;
        repeat = base - candidate; /* :176 */
        s += 4;
        candidate += 4;
        s = extend8(src, s, candidate, n - 8); /* :181-188 */
#ifdef MZO_STATS
        if (s - probe_pos >= 24) STAT(9); /* forward match beyond a 24-byte snapshot */
        mzo_stats[12] += (uint64_t)(s - base);
#endif
        int length = s - base;
        if (nextEmit != base) { /* :190-206 */
            if (base - nextEmit > maxFuseLits || repeat < kMinCopy2Offset) {
                if (d + (s - nextEmit) > dstLimit) {
                    free(table);
                    return 0;
                }
                d += mzo_emit_literal(dst + d, src + nextEmit, (size_t)(base - nextEmit));
                d += mzo_emit_copy(dst + d, repeat, length);
            } else if (repeat <= kMaxCopy2Offset) {
                d += mzo_emit_copy_lits2(dst + d, src + nextEmit, base - nextEmit, repeat, length);
            } else {
                d += mzo_emit_copy_lits3(dst + d, src + nextEmit, base - nextEmit, repeat, length);
            }
        } else {
            d += mzo_emit_copy(dst + d, repeat, length);
        }

        /* immediate re-match loop, :222-265 */
        for (;;) {
            nextEmit = s;
            if (s >= sLimit) goto emit_remainder;
            uint64_t x = ld64(src, s - 2);
            if (d > dstLimit) { /* :229 */
                free(table);
                return 0;
            }
            uint32_t m2Hash = hashN(x, tableBits, hashBytes);
            x >>= 16;
            uint32_t currHash = hashN(x, tableBits, hashBytes);
            candidate = (int)table[currHash];
            table[m2Hash] = (uint32_t)(s - 2);
            table[currHash] = (uint32_t)s;
            STAT(6);
            if (s - candidate > kMaxCopy3Offset || (uint32_t)x != ld32(src, candidate)) { /* :242 */
                STAT(7);
                cv = ld64(src, s + 1);
                s++;
                break;
            }
            repeat = s - candidate;
            base = s;
            s += 4;
            candidate += 4;
            s = extend8(src, s, candidate, n - 8);
#ifdef MZO_STATS
            if (s - base >= 24) STAT(10);
            mzo_stats[13] += (uint64_t)(s - base);
#endif
            d += mzo_emit_copy(dst + d, repeat, s - base);
        }
    }

emit_remainder: /* :268-282 */
    free(table);
    if (nextEmit < n) {
        if (d + n - nextEmit > dstLimit) return 0;
        d += mzo_emit_literal(dst + d, src + nextEmit, (size_t)(n - nextEmit));
    }
    return d;
}

/* asm_none.go:51-59 encodeBlock */
int64_t mzo_encode_block_l1(uint8_t *dst, const uint8_t *src, size_t n) {
    if (n < kMinNonLiteralBlockSize || n > MZO_MAX_BLOCK_SIZE) return 0;
    if (n <= 65536) return encode_l1(dst, src, (int)n, 13, 5, 5, kMaxCopy2Lits);
    return encode_l1(dst, src, (int)n, 15, 6, 6, kMaxCopy3Lits);
}

/* ---- L0 "SuperFast": encode_l0.go:32-279 (tableBits 13, skipLog 5, step 5,
 *      dstLimit len - len>>3 - 6, fuse<=3) and :281-522 (tableBits 12, skipLog 4,
 *      step 4, dstLimit len - len>>4 - 32, fuse<=4).  Same skeleton as L1 with an
 *      8-byte minimum match (hash8, full 8-byte compares), no backward extension
 *      (encode_l0.go:164 `for false && ...`) and extension from +8. ---------- */
static int64_t encode_l0(uint8_t *dst, const uint8_t *src, int n, int tableBits, int skipLog, int step, int dstLimit,
                         int maxFuseLits) {
    uint32_t *table = (uint32_t *)calloc((size_t)1 << tableBits, sizeof(uint32_t));
    if (!table) return 0;
    const int sLimit = n - kInputMargin;
    int nextEmit = 0;
    int s = 1;
    uint64_t cv = ld64(src, s);
    int repeat = 1;
    int d = 0;
    int candidate;

    for (;;) {
        candidate = 0;
        for (;;) {
            int nextS = s + ((s - nextEmit) >> skipLog) + step; /* :60 */
            if (nextS > sLimit) goto emit_remainder;
            int minSrcPos = s - kMaxCopy3Offset;
            uint32_t hash0 = hash8(cv, tableBits);
            uint64_t cv1 = ld64(src, s + 1);
            uint32_t hash1 = hash8(cv1, tableBits);
            candidate = (int)table[hash0];
            int candidate2 = (int)table[hash1];
            table[hash0] = (uint32_t)s;
            table[hash1] = (uint32_t)(s + 1);
            uint64_t cv2 = ld64(src, s + 2);
            uint32_t hash2 = hash8(cv2, tableBits);

            if ((uint32_t)cv1 == ld32(src, s - repeat + 1)) { /* :77 */
                int base = s + 1;
                for (int i = base - repeat; base > nextEmit && i > 0 && src[i - 1] == src[base - 1];) {
                    i--;
                    base--;
                }
                if (d + (base - nextEmit) > dstLimit) {
                    free(table);
                    return 0;
                }
                d += mzo_emit_literal(dst + d, src + nextEmit, (size_t)(base - nextEmit));
                int cand = s - repeat + 4 + 1;
                s += 4 + 1;
                s = extend8(src, s, cand, sLimit);
                d += mzo_emit_repeat(dst + d, s - base);
                nextEmit = s;
                if (s >= sLimit) goto emit_remainder;
                cv = ld64(src, s);
                continue;
            }
            if (candidate >= minSrcPos && cv == ld64(src, candidate)) break; /* :142 */
            candidate = (int)table[hash2];
            if (candidate2 >= minSrcPos && cv1 == ld64(src, candidate2)) {
                table[hash2] = (uint32_t)(s + 2);
                candidate = candidate2;
                s++;
                break;
            }
            table[hash2] = (uint32_t)(s + 2);
            if (candidate >= minSrcPos && cv2 == ld64(src, candidate)) {
                s += 2;
                break;
            }
            cv = ld64(src, nextS);
            s = nextS;
        }
        /* no backward extension (:164) */
        int base = s;
        repeat = base - candidate;
        s += 8; /* :174 */
        candidate += 8;
        s = extend8(src, s, candidate, n - 8);
        int length = s - base;
        if (nextEmit != base) {
            if (base - nextEmit > maxFuseLits || repeat < kMinCopy2Offset) {
                if (d + (s - nextEmit) > dstLimit) {
                    free(table);
                    return 0;
                }
                d += mzo_emit_literal(dst + d, src + nextEmit, (size_t)(base - nextEmit));
                d += mzo_emit_copy(dst + d, repeat, length);
            } else if (repeat <= kMaxCopy2Offset) {
                d += mzo_emit_copy_lits2(dst + d, src + nextEmit, base - nextEmit, repeat, length);
            } else {
                d += mzo_emit_copy_lits3(dst + d, src + nextEmit, base - nextEmit, repeat, length);
            }
        } else {
            d += mzo_emit_copy(dst + d, repeat, length);
        }
        for (;;) { /* :218-261 */
            nextEmit = s;
            if (s >= sLimit) goto emit_remainder;
            uint64_t x = ld64(src, s - 2);
            if (d > dstLimit) {
                free(table);
                return 0;
            }
            uint32_t m2Hash = hash8(x, tableBits);
            x = ld64(src, s);
            uint32_t currHash = hash8(x, tableBits);
            candidate = (int)table[currHash];
            table[m2Hash] = (uint32_t)(s - 2);
            table[currHash] = (uint32_t)s;
            if (s - candidate > kMaxCopy3Offset || x != ld64(src, candidate)) {
                cv = ld64(src, s + 1);
                s++;
                break;
            }
            repeat = s - candidate;
            base = s;
            s += 8;
            candidate += 8;
            s = extend8(src, s, candidate, n - 8);
            d += mzo_emit_copy(dst + d, repeat, s - base);
        }
    }
emit_remainder:
    free(table);
    if (nextEmit < n) {
        if (d + n - nextEmit > dstLimit) return 0;
        d += mzo_emit_literal(dst + d, src + nextEmit, (size_t)(n - nextEmit));
    }
    return d;
}

/* asm_none.go:33-43 encodeBlockFast */
int64_t mzo_encode_block_l0(uint8_t *dst, const uint8_t *src, size_t n) {
    if (n < kMinNonLiteralBlockSize || n > MZO_MAX_BLOCK_SIZE) return 0;
    int nn = (int)n;
    if (n <= 65536) return encode_l0(dst, src, nn, 12, 4, 4, nn - (nn >> 4) - 32, kMaxCopy2Lits);
    return encode_l0(dst, src, nn, 13, 5, 5, nn - (nn >> 3) - 6, kMaxCopy3Lits);
}

/* ---- L1 / L0, AMD64-assembly flavour ------------------------------------
 * What `go build` runs on amd64: the functions generated by
 * _generate/gen.go:257-1155 genEncodeBlockAsm(name, tableBits, skipLog,
 * hashBytes, maxLen) under fastOpts (gen.go:57-76), selected per block length by
 * encode_amd64.go:37-189.  Same walk as the Go functions above; what differs
 * (README.md:375 "will often produce slightly different output"):
 *   sLimit = len-17 and the search ends on nextS >= sLimit      gen.go:52-53,369,459
 *   dstLimit = len - len>>minSizeLog - 17; every bail test is
 *     `dst [+ lits + literalMaxOverhead] >= dstLimit`             gen.go:380-417
 *   matches and repeats extend with matchLen to the END of src   gen.go:632-654,859-887
 *   table bits / skipLog / hash bytes / step per size class      gen.go:57-76
 *   8 MiB variant clamps far candidates to s-2162685 (CMOV)      gen.go:466-490
 *   Fast (L0): 8-byte compares, no literal fusing, no back-extension.
 * PINNED: tests/test_ref_asm.py runs the reference's real assembly
 * (oracle/_ref, see ref_shim.c) and requires byte identity on every corpus. */
typedef struct {
    int tableBits, skipLog, hashBytes, maxLen;
    int match8, fuselits, checkBack, incLoop, minSizeLog;
} asm_opts;

static inline uint32_t hashNx(uint64_t u, int h, int nbytes) { /* gen.go:2072-2128 */
    switch (nbytes) {
    case 4: return (uint32_t)(((u << 32) * 2654435761ull) >> (64 - h));
    case 5: return hash5(u, h);
    case 6: return hash6(u, h);
    case 7: return hash7(u, h);
    default: return hash8(u, h);
    }
}

/* gen.go:3190-3288 matchLen: exact common prefix of src[a..] and src[b..], at most `left` bytes */
static inline int match_len_full(const uint8_t *src, int a, int b, int left) {
    int n = 0;
    while (left - n >= 8) {
        uint64_t diff = ld64(src, a + n) ^ ld64(src, b + n);
        if (diff) return n + (__builtin_ctzll(diff) >> 3);
        n += 8;
    }
    while (n < left && src[a + n] == src[b + n]) n++;
    return n;
}

/* gen.go:2161-2230 inline emitLiteral.  Reference quirk kept for parity: when
 * maxLen is exactly 64 KiB the generator omits the `n < 1<<16` compare
 * (gen.go:2193-2198, `maxLen >= 30+1<<16` is false) but still emits the
 * four-byte form (gen.go:2200, `maxLen >= 1<<16`), so the 64K variants write
 * literal runs of 286+ bytes with a 3-byte length (valid, not canonical). */
static int emit_literal_asm(uint8_t *dst, const uint8_t *lit, size_t len, int maxLen) {
    if (maxLen >= (1 << 16) && maxLen < 30 + (1 << 16) && len >= 1 + 29 + 256) {
        uint32_t v = (uint32_t)len - 30;
        dst[0] = 31 << 3 | TAG_LITERAL;
        dst[1] = (uint8_t)v;
        dst[2] = (uint8_t)(v >> 8);
        dst[3] = (uint8_t)(v >> 16);
        memcpy(dst + 4, lit, len);
        return 4 + (int)len;
    }
    return mzo_emit_literal(dst, lit, len);
}

static int64_t encode_fast_asm(uint8_t *dst, const uint8_t *src, int n, const asm_opts *o) {
    const int tb = o->tableBits, hb = o->hashBytes;
    const int maxOffset = o->maxLen - 1;                         /* gen.go:291 */
    const int clampFar = maxOffset > kMaxCopy3Offset;            /* gen.go:466,472 */
    const int litOverhead = o->maxLen < 30 ? 1 : o->maxLen < 256 ? 2 : o->maxLen < 65536 ? 3 : 4; /* gen.go:1157-1169 */
    const int mlen = o->match8 ? 8 : 4;
    uint32_t *table = (uint32_t *)calloc((size_t)1 << tb, sizeof(uint32_t)); /* zero_loop, gen.go:337-353 */
    if (!table) return 0;
    const int sLimit = n - 17;                                   /* gen.go:369 */
    const int64_t dstLimit = (int64_t)(n - 17) - (n >> o->minSizeLog); /* gen.go:380-391 */
    int nextEmit = 0, s = 1, repeat = 1;
    int64_t d = 0;
    int candidate, base, length, offset;
    uint64_t cv;
#define BAIL_LITS(l) do { if (d + (l) + litOverhead >= dstLimit) { free(table); return 0; } } while (0) /* gen.go:395-417 */
#define BAIL() do { if (d >= dstLimit) { free(table); return 0; } } while (0)
#define CLAMP(c, minPos) do { if (clampFar && (c) <= (minPos)) (c) = (minPos); } while (0)        /* gen.go:477-482 */
#define CVEQ(pos, v) (o->match8 ? ld64(src, (pos)) == (v) : ld32(src, (pos)) == (uint32_t)(v))

search_loop:
    {
        int nextS = s + ((s - nextEmit) >> o->skipLog) + o->incLoop;        /* gen.go:449-455 */
        if ((uint32_t)nextS >= (uint32_t)sLimit) goto emit_remainder;      /* gen.go:458-461 */
        cv = ld64(src, s);
        int minPos = s - kMaxCopy3Offset + 2;                             /* gen.go:467-469 */
        uint32_t hash0 = hashNx(cv, tb, hb);
        uint32_t hash1 = hashNx(hb > 7 ? ld64(src, s + 1) : cv >> 8, tb, hb);
        candidate = (int)table[hash0];
        int candidate2 = (int)table[hash1];
        table[hash0] = (uint32_t)s;
        table[hash1] = (uint32_t)(s + 1);
        uint32_t hash2 = hashNx(hb > 6 ? ld64(src, s + 2) : cv >> 16, tb, hb);

        if ((uint32_t)(cv >> 8) == ld32(src, s - repeat + 1)) {            /* gen.go:566-583 */
            base = s + 1;
            if (o->checkBack) {                                            /* gen.go:593-612 */
                int i = base - repeat;
                while (i != 0 && base > nextEmit && src[i - 1] == src[base - 1]) {
                    base--;
                    i--;
                }
            }
            int litLen = base - nextEmit;
            BAIL_LITS(litLen);                                             /* gen.go:619 */
            d += emit_literal_asm(dst + d, src + nextEmit, (size_t)litLen, o->maxLen);
            s += 5;                                                        /* gen.go:631 */
            s += match_len_full(src, s, s - repeat, n - s);                /* gen.go:634-660 */
            d += mzo_emit_repeat(dst + d, s - base);
            nextEmit = s;
            goto search_loop;                                              /* gen.go:688 */
        }
        /* no_repeat_found, gen.go:690-798 */
        CLAMP(candidate, minPos);
        if (CVEQ(candidate, cv)) goto candidate_match;
        cv = hb > 7 ? ld64(src, s + 1) : cv >> 8;
        candidate = (int)table[hash2];
        CLAMP(candidate2, minPos);
        if (CVEQ(candidate2, cv)) {
            table[hash2] = (uint32_t)(s + 2);
            s++;
            candidate = candidate2;
            goto candidate_match;
        }
        table[hash2] = (uint32_t)(s + 2);
        cv = hb > 6 ? ld64(src, s + 2) : cv >> 8;
        CLAMP(candidate, minPos);
        if (CVEQ(candidate, cv)) {
            s += 2;
            goto candidate_match;
        }
        s = nextS;
        goto search_loop;
    }

candidate_match:
    if (o->checkBack)                                                      /* gen.go:803-824 */
        while (candidate != 0 && s > nextEmit && src[candidate - 1] == src[s - 1]) {
            s--;
            candidate--;
        }
    BAIL();                                                                /* gen.go:828 */
    base = s;
    repeat = s - candidate;
    s += mlen;
    candidate += mlen;
    length = match_len_full(src, s, candidate, n - s);                     /* gen.go:853-887 */
    s += length;
    length += mlen;
    offset = repeat;
    {
        int litLen = base - nextEmit, ne = nextEmit;
        nextEmit = s;
        if (litLen == 0) goto emit_copy;
        if (o->fuselits && litLen <= 3 && offset >= 64) {                  /* gen.go:907-943 */
            if (maxOffset > kMaxCopy2Offset && offset > kMaxCopy2Offset)
                d += mzo_emit_copy_lits3(dst + d, src + ne, litLen, offset, length);
            else
                d += mzo_emit_copy_lits2(dst + d, src + ne, litLen, offset, length);
            goto copy_done;
        }
        BAIL_LITS(litLen);                                                 /* gen.go:945 */
        d += emit_literal_asm(dst + d, src + ne, (size_t)litLen, o->maxLen);
    }
emit_copy:
    d += mzo_emit_copy(dst + d, offset, length);                           /* gen.go:951 */
copy_done:
    if ((uint32_t)s >= (uint32_t)sLimit) goto emit_remainder;              /* gen.go:955-958 */
    BAIL();                                                                /* gen.go:962-975 */
    {   /* immediate re-match, gen.go:977-1090 */
        uint64_t x = ld64(src, s - 2);
        uint32_t h0 = hashNx(x, tb, hb);
        cv = (hb > 6 || o->match8) ? ld64(src, s) : x >> 16;
        uint32_t h1 = hashNx(cv, tb, hb);
        candidate = (int)table[h1];
        table[h0] = (uint32_t)(s - 2);
        table[h1] = (uint32_t)s;
        base = s;
        s++;
        if (clampFar && candidate <= base - kMaxCopy3Offset) goto search_loop; /* gen.go:1007-1021 */
        if (!CVEQ(candidate, cv)) goto search_loop;
        repeat = base - candidate;
        BAIL();                                                            /* gen.go:1039 */
        s += mlen - 1;
        candidate += mlen;
        length = match_len_full(src, s, candidate, n - s);
        s += length;
        length += mlen;
        nextEmit = s;
        offset = repeat;
        goto emit_copy;
    }

emit_remainder:                                                            /* gen.go:1092-1119 */
    free(table);
    if (n - nextEmit != 0) {
        if (d + (n - nextEmit) + litOverhead >= dstLimit) return 0;
        d += emit_literal_asm(dst + d, src + nextEmit, (size_t)(n - nextEmit), o->maxLen);
    }
    return d;
#undef BAIL_LITS
#undef BAIL
#undef CLAMP
#undef CVEQ
}

/* encode_amd64.go:119-189 encodeBlock -> variants of gen.go:57-66 */
int64_t mzo_encode_block_l1_asm(uint8_t *dst, const uint8_t *src, size_t n) {
    if (n > MZO_MAX_BLOCK_SIZE) return 0;
    asm_opts o = {15, 6, 6, 8 << 20, 0, 1, 1, 4, 5};
    if (n > (2u << 20)) { /* encodeBlockAsm */ }
    else if (n > (512u << 10)) o.maxLen = 2 << 20;
    else if (n > (64u << 10)) { o.tableBits = 14; o.maxLen = 512 << 10; }
    else if (n > (16u << 10)) { o.tableBits = 13; o.skipLog = 5; o.maxLen = 64 << 10; o.incLoop = 3; }
    else if (n > (4u << 10)) { o.tableBits = 12; o.skipLog = 5; o.hashBytes = 5; o.maxLen = 16 << 10; o.incLoop = 3; }
    else if (n > (1u << 10)) { o.tableBits = 10; o.skipLog = 5; o.hashBytes = 4; o.maxLen = 4 << 10; o.incLoop = 3; }
    else if (n > kMinNonLiteralBlockSize) { o.tableBits = 9; o.skipLog = 4; o.hashBytes = 4; o.maxLen = 1 << 10; o.incLoop = 3; }
    else return 0;
    return encode_fast_asm(dst, src, (int)n, &o);
}

/* encode_amd64.go:37-107 encodeBlockFast -> variants of gen.go:68-76 */
int64_t mzo_encode_block_l0_asm(uint8_t *dst, const uint8_t *src, size_t n) {
    if (n > MZO_MAX_BLOCK_SIZE) return 0;
    asm_opts o = {14, 5, 8, 8 << 20, 1, 0, 0, 4, 3};
    if (n > (2u << 20)) { /* encodeFastBlockAsm */ }
    else if (n > (512u << 10)) { o.tableBits = 13; o.maxLen = 2 << 20; }
    else if (n > (64u << 10)) { o.tableBits = 13; o.maxLen = 512 << 10; }
    else if (n > (16u << 10)) { o.tableBits = 12; o.skipLog = 4; o.maxLen = 64 << 10; }
    else if (n > (4u << 10)) { o.tableBits = 11; o.skipLog = 4; o.maxLen = 16 << 10; }
    else if (n > (1u << 10)) { o.tableBits = 10; o.skipLog = 4; o.maxLen = 4 << 10; }
    else if (n > 32) { o.tableBits = 9; o.skipLog = 3; o.maxLen = 1 << 10; }
    else return 0;
    return encode_fast_asm(dst, src, (int)n, &o);
}

/* ---- L2, AMD64-assembly flavour -------------------------------------------
 * _generate/gen.go:1171-2038 genEncodeBetterBlockAsm(name, lTableBits, sTableBits,
 * skipLog, lHashBytes, maxLen) with the per-class options of gen.go:78-88, selected
 * by encode_amd64.go:201-271.  Differences from encodeBlockBetterGo: margins
 * (sLimit = len-17 / len-8, `>=` exits), bail tests `dst + lits + overhead >=
 * dstLimit` (gen.go:1490-1508,1716-1735), skip capped at maxSkip = 100 in the three
 * large classes (gen.go:1327-1354), matchLen to the end of the block, the far
 * 4-byte-match rejection at offset > 65599 (Go: > 65535) without an sLimit test
 * (gen.go:1786-1801), candidates clamped (CMOV) in the 8 MiB class, the per-class
 * tables and the 64 KiB literal quirk.  PINNED by tests/test_ref_asm.py against the
 * real assembly. */
typedef struct {
    int lBits, sBits, skipLog, lHashBytes, maxLen, maxSkip, outMargin, inMargin;
} better_opts;

static int64_t encode_better_asm(uint8_t *dst, const uint8_t *src, int n, const better_opts *o) {
    const int maxOffset = o->maxLen - 1;
    const int clampFar = maxOffset > kMaxCopy3Offset;
    const int litOverhead = o->maxLen < 30 ? 1 : o->maxLen < 256 ? 2 : o->maxLen < 65536 ? 3 : 4;
    uint32_t *lTab = (uint32_t *)calloc(((size_t)1 << o->lBits) + ((size_t)1 << o->sBits), sizeof(uint32_t));
    if (!lTab) return 0;
    uint32_t *sTab = lTab + ((size_t)1 << o->lBits);
    const int sLimit = n - o->inMargin;                                    /* gen.go:1272-1282 */
    const int64_t dstLimit = (int64_t)(n - o->outMargin) - (n >> 5);       /* gen.go:1284-1297 */
    int nextEmit = 0, s = 1, repeat = 1, nextS = 0;
    int64_t d = 0;
    int candidate, base, length, offset;
    uint64_t cv;
#define LH(v) hashNx((v), o->lBits, o->lHashBytes)
#define SH(v) hashNx((v), o->sBits, 4)
#define RET0() do { free(lTab); return 0; } while (0)
#define CLAMP(c, minPos) do { if (clampFar && (c) <= (minPos)) (c) = (minPos); } while (0)

search_loop:
    {
        uint32_t skip = (uint32_t)(s - nextEmit) >> o->skipLog;            /* gen.go:1324-1354 */
        if (o->maxSkip == 0 || skip <= (uint32_t)(o->maxSkip - 1)) nextS = s + (int)skip + 1;
        else nextS = s + o->maxSkip;
        if ((uint32_t)nextS >= (uint32_t)sLimit) goto emit_remainder;
        cv = ld64(src, s);
        uint32_t hash0 = LH(cv), hash1 = SH(cv);
        candidate = (int)lTab[hash0];
        int candidateS = (int)sTab[hash1];
        lTab[hash0] = (uint32_t)s;
        sTab[hash1] = (uint32_t)s;
        const int minPos = s - kMaxCopy3Offset + 2;                        /* gen.go:1396-1399 */
        CLAMP(candidate, minPos);
        const uint64_t longVal = ld64(src, candidate);
        if (longVal == cv) goto candidate_match;                           /* gen.go:1425-1429 */
        CLAMP(candidateS, minPos);
        const uint64_t shortVal = ld64(src, candidateS);

        if (((ld64(src, s - repeat) ^ cv) & (0xffffffffull << 8)) == 0) {  /* gen.go:1445-1459 */
            base = s + 1;
            for (int i = base - repeat; i != 0 && base > nextEmit && src[i - 1] == src[base - 1];) {
                base--;
                i--;
            }
            if (d + (base - nextEmit) + litOverhead >= dstLimit) RET0();   /* gen.go:1490-1508 */
            if (nextEmit != base) {
                d += emit_literal_asm(dst + d, src + nextEmit, (size_t)(base - nextEmit), o->maxLen);
                nextEmit = base;
            }
            s += 5;
            s += match_len_full(src, s, s - repeat, n - s);
            d += mzo_emit_repeat(dst + d, s - base);
            nextEmit = s;
            if ((uint32_t)s >= (uint32_t)sLimit) goto emit_remainder;      /* gen.go:1576-1577 */
            for (int64_t i0 = base + 1, i1 = s - 2; i0 < i1; i0 += 2, i1 -= 2) { /* gen.go:1587-1622 */
                lTab[LH(ld64(src, i0))] = (uint32_t)i0;
                sTab[SH(ld64(src, i0 + 1))] = (uint32_t)(i0 + 1);
                lTab[LH(ld64(src, i1))] = (uint32_t)i1;
                sTab[SH(ld64(src, i1 + 1))] = (uint32_t)(i1 + 1);
            }
            goto search_loop;
        }
        /* no_repeat_found, gen.go:1625-1690 */
        if ((uint32_t)longVal == (uint32_t)cv) goto candidate_match;
        if ((uint32_t)shortVal == (uint32_t)cv) {
            /* short match at s: try a long candidate at s+1 */
            cv >>= 8;
            hash0 = LH(cv);
            candidate = (int)lTab[hash0];
            s++;
            lTab[hash0] = (uint32_t)s;
            CLAMP(candidate, minPos);
            if (ld32(src, candidate) == (uint32_t)cv) goto candidate_match;
            s--;
            candidate = candidateS;
            goto candidate_match;
        }
        s = nextS;
        goto search_loop;
    }

candidate_match:
    while (candidate != 0 && s > nextEmit && src[candidate - 1] == src[s - 1]) { /* gen.go:1696-1717 */
        s--;
        candidate--;
    }
    if (d + (s - nextEmit) + litOverhead >= dstLimit) RET0();              /* gen.go:1720-1737 */
    base = s;
    s += 4;
    candidate += 4;
    length = match_len_full(src, s, candidate, n - s);
    offset = s - candidate;
    if (maxOffset > kMaxCopy2Offset && length == 0 && offset > kMaxCopy2Offset && offset != repeat) {
        s = nextS + 1;                                                     /* gen.go:1786-1801 */
        goto search_loop;
    }
    repeat = offset;
    {
        const int litLen = base - nextEmit, ne = nextEmit;
        s += length;
        length += 4;
        nextEmit = s;
        if (litLen == 0) {
            d += mzo_emit_copy(dst + d, offset, length);
        } else if (offset < kMinCopy2Offset) {
            d += emit_literal_asm(dst + d, src + ne, (size_t)litLen, o->maxLen);
            d += mzo_emit_copy(dst + d, offset, length);
        } else if (maxOffset > kMaxCopy2Offset && offset > kMaxCopy2Offset) { /* gen.go:1848-1866 */
            if (litLen > 3) {
                d += emit_literal_asm(dst + d, src + ne, (size_t)litLen, o->maxLen);
                d += mzo_emit_copy(dst + d, offset, length);
            } else {
                d += mzo_emit_copy_lits3(dst + d, src + ne, litLen, offset, length);
            }
        } else if (litLen > 4) {
            d += emit_literal_asm(dst + d, src + ne, (size_t)litLen, o->maxLen);
            d += mzo_emit_copy(dst + d, offset, length);
        } else {
            d += mzo_emit_copy_lits2(dst + d, src + ne, litLen, offset, length); /* gen.go:1826-1846 */
        }
    }
    if ((uint32_t)s >= (uint32_t)sLimit) goto emit_remainder;              /* gen.go:1899-1902 */
    if (d >= dstLimit) RET0();                                             /* gen.go:1905-1918 */
    {   /* index the match interior, gen.go:1921-1978 */
        int64_t i0 = base + 1, i1 = s - 2;
        lTab[LH(ld64(src, i0))] = (uint32_t)i0;
        lTab[LH(ld64(src, i1))] = (uint32_t)i1;
        sTab[SH(ld64(src, i0 + 1))] = (uint32_t)(i0 + 1);
        sTab[SH(ld64(src, i1 + 1))] = (uint32_t)(i1 + 1);
        int64_t i2 = (i0 + i1 + 1) >> 1;
        i0 += 1;
        i1 -= 1;
        for (; i2 < i1; i0 += 2, i2 += 2) {
            lTab[LH(ld64(src, i0))] = (uint32_t)i0;
            lTab[LH(ld64(src, i2))] = (uint32_t)i2;
        }
    }
    goto search_loop;

emit_remainder:                                                            /* gen.go:1980-2017 */
    free(lTab);
    if (d + (n - nextEmit) + litOverhead >= dstLimit) return 0;
    if (nextEmit != n) d += emit_literal_asm(dst + d, src + nextEmit, (size_t)(n - nextEmit), o->maxLen);
    return d;
#undef LH
#undef SH
#undef RET0
#undef CLAMP
}

/* encode_amd64.go:201-271 encodeBlockBetter -> variants of gen.go:78-88 */
int64_t mzo_encode_block_l2_asm(uint8_t *dst, const uint8_t *src, size_t n) {
    if (n > MZO_MAX_BLOCK_SIZE) return 0;
    better_opts o;
    if (n > (2u << 20)) o = (better_opts){17, 14, 8, 7, 8 << 20, 100, 17, 17};
    else if (n > (512u << 10)) o = (better_opts){17, 14, 7, 7, 2 << 20, 100, 17, 17};
    else if (n > (64u << 10)) o = (better_opts){16, 13, 7, 7, 512 << 10, 100, 11, 8};
    else if (n > (16u << 10)) o = (better_opts){15, 12, 6, 6, 64 << 10, 0, 11, 8};
    else if (n > (4u << 10)) o = (better_opts){14, 11, 6, 6, 16 << 10, 0, 11, 8};
    else if (n > (1u << 10)) o = (better_opts){12, 10, 5, 6, 4 << 10, 0, 11, 8};
    else if (n > kMinNonLiteralBlockSize) o = (better_opts){11, 8, 4, 6, 1 << 10, 0, 11, 8};
    else return 0;
    return encode_better_asm(dst, src, (int)n, &o);
}

/* ---- L2: encode_l2.go:61-338 (long 17 bit hash7 / short 14 bit hash4) and
 *          encode_l2.go:343-596 (long 15 bit hash6 / short 12 bit hash4).
 * As for L1, one parameterised body: the 64K variant drops guards that cannot
 * fire for len <= 64 KiB (minSrcPos, the far-4-byte-match bail, copy3). ----- */
static int64_t encode_l2(uint8_t *dst, const uint8_t *src, int n, int lBits, int lHashBytes, int sBits) {
    const int sLimit = n - kInputMargin; /* :65 */
    uint32_t *lTable = (uint32_t *)calloc(((size_t)1 << lBits) + ((size_t)1 << sBits), sizeof(uint32_t));
    if (!lTable) return 0;
    uint32_t *sTable = lTable + ((size_t)1 << lBits);
    const int dstLimit = n - (n >> 5) - 6; /* :92 */
    int nextEmit = 0;
    int s = 1;
    uint64_t cv = ld64(src, s);
    int repeat = 1; /* :102 */
    int d = 0;

#define LHASH(v) hashN((v), lBits, lHashBytes)
#define SHASH(v) hash4((v), sBits)

    for (;;) {
        int candidateL = 0;
        int nextS = 0;
        for (;;) {
            STAT(0); /* L2 search steps (shares the counter array with L1) */
            nextS = s + ((s - nextEmit) >> 7) + 1; /* :114 */
            if (nextS > sLimit) goto emit_remainder;
            int minSrcPos = s - kMaxCopy3Offset + 1; /* :118 */
            uint32_t hashL = LHASH(cv);
            uint32_t hashS = SHASH(cv);
            candidateL = (int)lTable[hashL];
            int candidateS = (int)sTable[hashS];
            lTable[hashL] = (uint32_t)s;
            sTable[hashS] = (uint32_t)s;

            uint64_t valLong = ld64(src, candidateL);
            uint64_t valShort = ld64(src, candidateS);

            if (candidateL > minSrcPos && cv == valLong) break; /* :130 */

            /* repeat at s+1, 4 bytes: :134-196 */
            const uint64_t repeatMask = 0xffffffffull << 8;
            if (repeat > 0 && (cv & repeatMask) == (ld64(src, s - repeat) & repeatMask)) {
                int base = s + 1;
                for (int i = base - repeat; base > nextEmit && i > 0 && src[i - 1] == src[base - 1];) {
                    i--;
                    base--;
                }
                if (d + (base - nextEmit) > dstLimit) { /* :147 */
                    free(lTable);
                    return 0;
                }
                d += mzo_emit_literal(dst + d, src + nextEmit, (size_t)(base - nextEmit));
                int cand = s - repeat + 4 + 1;
                s += 4 + 1;
                s = extend_tail(src, n, s, cand); /* :160-175 */
                d += mzo_emit_repeat(dst + d, s - base);
                nextEmit = s;
                if (s >= sLimit) goto emit_remainder; /* :179 */
                int index0 = base + 1;
                int index1 = s - 2;
                while (index0 < index1) { /* :186-195 */
                    uint64_t cv0 = ld64(src, index0);
                    uint64_t cv1 = ld64(src, index1);
                    lTable[LHASH(cv0)] = (uint32_t)index0;
                    sTable[SHASH(cv0 >> 8)] = (uint32_t)(index0 + 1);
                    lTable[LHASH(cv1)] = (uint32_t)index1;
                    sTable[SHASH(cv1 >> 8)] = (uint32_t)(index1 + 1);
                    index0 += 2;
                    index1 -= 2;
                }
                cv = ld64(src, s);
                continue;
            }

            if (candidateL >= minSrcPos && (uint32_t)cv == (uint32_t)valLong) break; /* :199 */

            if (candidateS >= minSrcPos && (uint32_t)cv == (uint32_t)valShort) { /* :204 */
                hashL = LHASH(cv >> 8);
                candidateL = (int)lTable[hashL];
                lTable[hashL] = (uint32_t)(s + 1);
                if (candidateL > minSrcPos && (uint32_t)(cv >> 8) == ld32(src, candidateL)) {
                    s++;
                    break;
                }
                candidateL = candidateS;
                break;
            }
            cv = ld64(src, nextS); /* :218 */
            s = nextS;
        }

        /* extend backwards, :223-226 */
        while (candidateL > 0 && s > nextEmit && src[candidateL - 1] == src[s - 1]) {
            candidateL--;
            s--;
        }
        if (d + (s - nextEmit) > dstLimit) { /* :229 */
            free(lTable);
            return 0;
        }
        int base = s;
        int offset = base - candidateL;
        s += 4;
        candidateL += 4;
        s = extend_tail(src, n, s, candidateL); /* :239-254 */
#ifdef MZO_STATS
        STAT(2);
        mzo_stats[12] += (uint64_t)(s - base);
#endif

        /* :257-264 drop far 4-byte matches */
        if (offset > 65535 && s - base <= 4 && repeat != offset) {
            s = nextS + 1;
            if (s >= sLimit) goto emit_remainder;
            cv = ld64(src, s);
            continue;
        }

        int nlits = base - nextEmit; /* :266-289 */
        if (nlits > 0) {
            if (offset <= kMaxCopy2Offset) {
                if (nlits > kMaxCopy2Lits || offset < 64) {
                    d += mzo_emit_literal(dst + d, src + nextEmit, (size_t)nlits);
                    d += mzo_emit_copy(dst + d, offset, s - base);
                } else {
                    d += mzo_emit_copy_lits2(dst + d, src + nextEmit, nlits, offset, s - base);
                }
            } else {
                if (nlits > kMaxCopy3Lits) {
                    d += mzo_emit_literal(dst + d, src + nextEmit, (size_t)nlits);
                    d += mzo_emit_copy(dst + d, offset, s - base);
                } else {
                    d += mzo_emit_copy_lits3(dst + d, src + nextEmit, nlits, offset, s - base);
                }
            }
        } else {
            d += mzo_emit_copy(dst + d, offset, s - base);
        }
        repeat = offset;
        nextEmit = s;
        if (s >= sLimit) goto emit_remainder; /* :293 */
        if (d > dstLimit) {                   /* :297 */
            free(lTable);
            return 0;
        }

        /* index short & long, :303-326 */
        int index0 = base + 1;
        int index1 = s - 2;
        uint64_t cv0 = ld64(src, index0);
        uint64_t cv1 = ld64(src, index1);
        lTable[LHASH(cv0)] = (uint32_t)index0;
        sTable[SHASH(cv0 >> 8)] = (uint32_t)(index0 + 1);
        lTable[LHASH(cv1)] = (uint32_t)index1;
        sTable[SHASH(cv1 >> 8)] = (uint32_t)(index1 + 1);
        index0 += 1;
        index1 -= 1;
        cv = ld64(src, s);
        int index2 = (index0 + index1 + 1) >> 1;
        while (index2 < index1) {
            lTable[LHASH(ld64(src, index0))] = (uint32_t)index0;
            lTable[LHASH(ld64(src, index2))] = (uint32_t)index2;
            index0 += 2;
            index2 += 2;
        }
    }

emit_remainder: /* :329-337 */
    free(lTable);
    if (nextEmit < n) {
        if (d + n - nextEmit > dstLimit) return 0;
        d += mzo_emit_literal(dst + d, src + nextEmit, (size_t)(n - nextEmit));
    }
    return d;
#undef LHASH
#undef SHASH
}

/* asm_none.go:68-76 encodeBlockBetter */
int64_t mzo_encode_block_l2(uint8_t *dst, const uint8_t *src, size_t n) {
    if (n < kMinNonLiteralBlockSize || n > MZO_MAX_BLOCK_SIZE) return 0;
    if (n <= (64 << 10)) return encode_l2(dst, src, (int)n, 15, 6, 12);
    return encode_l2(dst, src, (int)n, 17, 7, 14);
}

/* ---- decoder: decode.go:178-622 ---------------------------------------
 * One bounds-checked loop.  The Go function has a fast loop (src slack >= 11)
 * and a checked tail loop; they accept/reject identically and write identical
 * bytes on success (the fast loop's 4-byte fused-literal store :313-320 is
 * always overwritten by the >=4-byte copy that follows), so the checked tail
 * loop :362-611 is restated for every token. */
int mzo_decode_block(uint8_t *dst, size_t dst_len_, const uint8_t *src, size_t src_len_) {
    const int64_t dlen = (int64_t)dst_len_, slen = (int64_t)src_len_;
    int64_t d = 0, s = 0, length = 0;
    int64_t offset = 1; /* :186 */

    while (s < slen) {
        uint8_t tag = src[s];
        switch (tag & 3) {
        case 0: { /* literal / repeat :367-423 */
            uint32_t x = tag >> 3;
            if (x < 29) {
                s++;
                length = x + 1;
            } else if (x == 29) {
                s += 2;
                if (s > slen) return 1;
                length = (int64_t)src[s - 1] + 30;
            } else if (x == 30) {
                s += 3;
                if (s > slen) return 1;
                length = (int64_t)(src[s - 2] | (uint32_t)src[s - 1] << 8) + 30;
            } else {
                s += 4;
                if (s > slen) return 1;
                length = (int64_t)(src[s - 3] | (uint32_t)src[s - 2] << 8 | (uint32_t)src[s - 1] << 16) + 30;
            }
            if (tag & 4) break; /* repeat: goto doCopy2 */
            if (length > dlen - d || length > slen - s) return 1; /* :410 */
            memcpy(dst + d, src + s, (size_t)length);
            d += length;
            s += length;
            continue;
        }
        case 1: /* copy1 :425-448 */
            s += 2;
            if (s > slen) return 1;
            length = (src[s - 2] >> 2) & 15;
            offset = (int64_t)(ld16(src, s - 2) >> 6) + 1;
            if (length == 15) {
                s++;
                if (s > slen) return 1;
                length = (int64_t)src[s - 1] + 18;
            } else {
                length += 4;
            }
            break;
        case 2: /* copy2 :449-495 */
            s += 3;
            if (s > slen) return 1;
            length = src[s - 3] >> 2;
            offset = (int64_t)(src[s - 2] | (uint32_t)src[s - 1] << 8);
            if (length <= 60) {
                length += 4;
            } else if (length == 61) {
                s++;
                if (s > slen) return 1;
                length = (int64_t)src[s - 1] + 64;
            } else if (length == 62) {
                s += 2;
                if (s > slen) return 1;
                length = (int64_t)(src[s - 2] | (uint32_t)src[s - 1] << 8) + 64;
            } else {
                s += 3;
                if (s > slen) return 1;
                length = (int64_t)(src[s - 3] | (uint32_t)src[s - 2] << 8 | (uint32_t)src[s - 1] << 16) + 64;
            }
            offset += kMinCopy2Offset;
            break;
        default: { /* fused copy2 / copy3 :496-569 */
            s += 4;
            if (s > slen) return 1;
            uint32_t val = ld32(src, s - 4);
            int isCopy3 = (val & 4) != 0;
            int64_t litLen = (val >> 3) & 3;
            if (!isCopy3) {
                length = 4 + ((val >> 5) & 7);
                offset = (int64_t)((val >> 8) & 65535) + kMinCopy2Offset;
                s--;
                litLen++;
            } else {
                uint32_t lengthTmp = (val >> 5) & 63;
                offset = (int64_t)(val >> 11) + kMinCopy3Offset;
                if (lengthTmp >= 61) {
                    if (lengthTmp == 61) {
                        s++;
                        if (s > slen) return 1;
                        length = (int64_t)src[s - 1] + 64;
                    } else if (lengthTmp == 62) {
                        s += 2;
                        if (s > slen) return 1;
                        length = (int64_t)(src[s - 2] | (uint32_t)src[s - 1] << 8) + 64;
                    } else {
                        s += 3;
                        if (s > slen) return 1;
                        length = (int64_t)(src[s - 3] | (uint32_t)src[s - 2] << 8 | (uint32_t)src[s - 1] << 16) + 64;
                    }
                } else {
                    length = lengthTmp + 4;
                }
            }
            if (litLen > 0) { /* :556-568 */
                if (litLen > dlen - d || s + litLen > slen) return 1;
                memcpy(dst + d, src + s, (size_t)litLen);
                d += litLen;
                s += litLen;
            }
            break;
        }
        }
        /* doCopy2 :572-610 */
        if (offset <= 0 || d < offset || length > dlen - d) return 1;
        if (offset > length) {
            memcpy(dst + d, dst + d - offset, (size_t)length);
        } else {
            uint8_t *a = dst + d;
            const uint8_t *b = dst + d - offset;
            for (int64_t i = 0; i < length; i++) a[i] = b[i];
        }
        d += length;
    }
    if (d != dlen) return 1; /* :615 */
    return 0;
}

/* ---- wrappers ---------------------------------------------------------- */

/* encoding/binary PutUvarint */
static int put_uvarint(uint8_t *dst, uint64_t x) {
    int i = 0;
    while (x >= 0x80) {
        dst[i++] = (uint8_t)x | 0x80;
        x >>= 7;
    }
    dst[i] = (uint8_t)x;
    return i + 1;
}

/* encoding/binary Uvarint: n>0 ok, 0 short buffer, <0 overflow */
static int get_uvarint(const uint8_t *buf, size_t len, uint64_t *out) {
    uint64_t x = 0;
    unsigned sft = 0;
    for (size_t i = 0; i < len; i++) {
        uint8_t b = buf[i];
        if (i == 10) return -(int)(i + 1);
        if (b < 0x80) {
            if (i == 9 && b > 1) return -(int)(i + 1);
            *out = x | (uint64_t)b << sft;
            return (int)i + 1;
        }
        x |= (uint64_t)(b & 0x7f) << sft;
        sft += 7;
    }
    *out = 0;
    return 0;
}

/* encode.go:223-229 encodeUncompressed */
static int64_t encode_uncompressed(uint8_t *dst, size_t cap, const uint8_t *src, size_t n) {
    if (n == 0) {
        if (cap < 1) return MZO_ERR_DST_TOO_SMALL;
        dst[0] = 0;
        return 1;
    }
    if (cap < n + 2) return MZO_ERR_DST_TOO_SMALL;
    dst[0] = 0;
    dst[1] = 0;
    memcpy(dst + 2, src, n);
    return (int64_t)n + 2;
}

/* encode.go:74-139 Encode */
int64_t mzo_encode(uint8_t *dst, size_t dst_cap, const uint8_t *src, size_t n, int level) {
    int64_t maxlen = mzo_max_encoded_len((int64_t)n);
    if (maxlen < 0) return MZO_ERR_TOO_LARGE;
    if (n < kMinNonLiteralBlockSize) return encode_uncompressed(dst, dst_cap, src, n);
    /* Go allocates MaxEncodedLen when dst is short; a C caller must provide it.
     * The block encoders need headroom beyond n+2 while probing, as in Go
     * (dst[d:] where len(dst)=n+2, bails keep d below dstLimit). */
    if ((int64_t)dst_cap < maxlen) return MZO_ERR_DST_TOO_SMALL;
    if (level != -1 && level != 0 && level != 1 && level != 2) return MZO_ERR_INVALID_LEVEL;
    if (level == 0) return encode_uncompressed(dst, dst_cap, src, n);
    dst[0] = 0;
    int d = 1 + put_uvarint(dst + 1, n);
    /* The Go encoders may transiently write up to a few bytes past n+2-d while
     * emitting the token that trips the dstLimit bail; encode into a scratch
     * buffer with slack so the oracle never depends on that. */
    uint8_t *tmp = (uint8_t *)malloc(n + 64);
    if (!tmp) return MZO_ERR_DST_TOO_SMALL;
    int64_t m = level == 1 ? mzo_encode_block_l1(tmp, src, n) : level == 2 ? mzo_encode_block_l2(tmp, src, n) : mzo_encode_block_l0(tmp, src, n);
    if (m > 0) {
        memcpy(dst + d, tmp, (size_t)m);
        free(tmp);
        return d + m;
    }
    free(tmp);
    return encode_uncompressed(dst, dst_cap, src, n);
}

/* encode.go:168-207 TryEncode */
int64_t mzo_try_encode(uint8_t *dst, size_t dst_cap, const uint8_t *src, size_t n, int level) {
    int64_t maxlen = mzo_max_encoded_len((int64_t)n);
    if (maxlen < 0 || (int64_t)dst_cap < maxlen) return 0;
    if (n < kMinNonLiteralBlockSize) return 0;
    if (level != -1 && level != 1 && level != 2) return 0;
    dst[0] = 0;
    int d = 1 + put_uvarint(dst + 1, n);
    uint8_t *tmp = (uint8_t *)malloc(n + 64);
    if (!tmp) return 0;
    int64_t m = level == 1 ? mzo_encode_block_l1(tmp, src, n) : level == 2 ? mzo_encode_block_l2(tmp, src, n) : mzo_encode_block_l0(tmp, src, n);
    if (m > 0 && d + m < (int64_t)n) {
        memcpy(dst + d, tmp, (size_t)m);
        free(tmp);
        return d + m;
    }
    free(tmp);
    return 0;
}

/* decode.go:160-171 decodedLen */
static int decoded_len_hdr(const uint8_t *src, size_t n, int64_t *v, int *hdr) {
    uint64_t x;
    int k = get_uvarint(src, n, &x);
    if (k <= 0 || x > 0xffffffffull) return MZO_ERR_CORRUPT;
    *v = (int64_t)x;
    *hdr = k;
    return MZO_OK;
}

/* decode.go:120-156 isMinLZ */
int mzo_is_minlz(const uint8_t *src, size_t n, int *is_mlz, int *lits, size_t *hdr, int64_t *size) {
    *is_mlz = 0;
    *lits = 0;
    *hdr = 0;
    *size = 0;
    if (n <= 1) {
        if (n == 0) return MZO_ERR_CORRUPT;
        if (src[0] == 0) {
            *is_mlz = 1;
            *lits = 1;
            *hdr = 1;
            return MZO_OK;
        }
    }
    if (src[0] != 0) {
        int64_t v;
        int h;
        int e = decoded_len_hdr(src, n, &v, &h);
        if (e) return e;
        *size = v;
        return MZO_OK; /* not MinLZ: Snappy/S2 territory */
    }
    int64_t v;
    int h;
    int e = decoded_len_hdr(src + 1, n - 1, &v, &h);
    if (e) return e;
    if (v > MZO_MAX_BLOCK_SIZE) return MZO_ERR_TOO_LARGE;
    size_t rest = n - 1 - (size_t)h;
    if (rest == 0) return MZO_ERR_CORRUPT;
    *hdr = 1 + (size_t)h;
    if (v == 0) {
        *is_mlz = 1;
        *lits = 1;
        *size = (int64_t)rest;
        return MZO_OK;
    }
    if (v < (int64_t)rest) {
        *size = v;
        return MZO_ERR_CORRUPT;
    }
    *is_mlz = 1;
    *size = v;
    return MZO_OK;
}

/* decode.go:107-111 DecodedLen */
int64_t mzo_decoded_len(const uint8_t *src, size_t n) {
    int a, b;
    size_t h;
    int64_t v;
    int e = mzo_is_minlz(src, n, &a, &b, &h, &v);
    return e ? e : v;
}

/* decode.go:50-78 Decode (MinLZ blocks only; first byte != 0 is the S2/Snappy
 * fallback, which lives in host Go and is outside this path). */
int64_t mzo_decode(uint8_t *dst, size_t dst_cap, const uint8_t *src, size_t n) {
    int is_mlz, lits;
    size_t hdr;
    int64_t size;
    int e = mzo_is_minlz(src, n, &is_mlz, &lits, &hdr, &size);
    if (e) return e;
    if (lits) {
        if ((size_t)size > dst_cap) return MZO_ERR_DST_TOO_SMALL;
        memcpy(dst, src + hdr, (size_t)size);
        return size;
    }
    if (!is_mlz) return MZO_ERR_UNSUPPORTED;
    if ((size_t)size > dst_cap) return MZO_ERR_DST_TOO_SMALL;
    if (mzo_decode_block(dst, (size_t)size, src + hdr, n - hdr) != 0) return MZO_ERR_CORRUPT;
    return size;
}

/* ---- CRC32C: minlz.go:133-140 ------------------------------------------ */
static uint32_t crc_tab[8][256];
static pthread_once_t crc_once = PTHREAD_ONCE_INIT;
static void crc_init(void) {
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t c = i;
        for (int k = 0; k < 8; k++) c = (c & 1) ? (c >> 1) ^ 0x82f63b78u : c >> 1;
        crc_tab[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; i++)
        for (int t = 1; t < 8; t++) crc_tab[t][i] = (crc_tab[t - 1][i] >> 8) ^ crc_tab[0][crc_tab[t - 1][i] & 0xff];
}

uint32_t mzo_crc(const uint8_t *b, size_t n) {
    pthread_once(&crc_once, crc_init);
    uint32_t c = 0xffffffffu;
    while (n >= 8) {
        uint64_t v = ld64(b, 0) ^ c;
        c = crc_tab[7][v & 0xff] ^ crc_tab[6][(v >> 8) & 0xff] ^ crc_tab[5][(v >> 16) & 0xff] ^
            crc_tab[4][(v >> 24) & 0xff] ^ crc_tab[3][(v >> 32) & 0xff] ^ crc_tab[2][(v >> 40) & 0xff] ^
            crc_tab[1][(v >> 48) & 0xff] ^ crc_tab[0][v >> 56];
        b += 8;
        n -= 8;
    }
    while (n--) c = crc_tab[0][(c ^ *b++) & 0xff] ^ (c >> 8);
    c = ~c;
    return (c >> 15 | c << 17) + 0xa282ead8u;
}

/* ---- multi-threaded batch drivers (CPU baseline leg) ------------------- */
typedef struct {
    int level, nblk;
    const uint8_t *src;
    const uint64_t *src_off;
    uint8_t *dst;
    const uint64_t *dst_off;
    uint32_t *out_len;
    int32_t *status;
    volatile int next;
    int decode;
} batch_job;

static void *batch_worker(void *arg) {
    batch_job *j = (batch_job *)arg;
    for (;;) {
        int i = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
        if (i >= j->nblk) break;
        const uint8_t *s = j->src + j->src_off[i];
        size_t sn = (size_t)(j->src_off[i + 1] - j->src_off[i]);
        uint8_t *d = j->dst + j->dst_off[i];
        size_t dn = (size_t)(j->dst_off[i + 1] - j->dst_off[i]);
        if (j->decode) {
            j->status[i] = mzo_decode_block(d, dn, s, sn);
        } else {
            int64_t m = j->level == 1 ? mzo_encode_block_l1(d, s, sn) : j->level == 2 ? mzo_encode_block_l2(d, s, sn) : mzo_encode_block_l0(d, s, sn);
            j->out_len[i] = (uint32_t)m;
        }
    }
    return NULL;
}

static int run_batch(batch_job *j, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    j->next = 0;
    for (int t = 1; t < nthreads; t++)
        if (pthread_create(&th[t], NULL, batch_worker, j)) return -1;
    batch_worker(j);
    for (int t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
    return 0;
}

int mzo_encode_batch_mt(int level, int nblk, const uint8_t *src, const uint64_t *src_off, uint8_t *dst,
                        const uint64_t *dst_off, uint32_t *out_len, int nthreads) {
    if (level != -1 && level != 1 && level != 2) return MZO_ERR_INVALID_LEVEL;
    batch_job j = {level, nblk, src, src_off, dst, dst_off, out_len, NULL, 0, 0};
    return run_batch(&j, nthreads);
}

int mzo_decode_batch_mt(int nblk, const uint8_t *src, const uint64_t *src_off, uint8_t *dst,
                        const uint64_t *dst_off, int32_t *status, int nthreads) {
    batch_job j = {0, nblk, src, src_off, dst, dst_off, NULL, status, 0, 1};
    return run_batch(&j, nthreads);
}
