// mz_decode.cuh -- MinLZ block decode kernels (sm_100a).
//
// Replaces minLZDecode (reference decode.go:178-622 / decodeBlockAsm,
// asm_amd64.s:27062): tag dispatch on the low 2 bits, literal copy, repeat
// offset, overlapping back-reference copy.  Result contract is the
// reference's: status 0 and exactly dst_len bytes written, or status 1
// (decodeErrCodeCorrupt); a corrupt block never writes outside its dst range.
#pragma once

#include "mz_common.cuh"

namespace mz {

// One decoded token: `hdr` header bytes, then `lit` literal bytes in the
// stream, then a copy of `mlen` bytes from `off` back (mlen == 0: none).
struct Token {
    uint32_t hdr, lit, mlen, off;
    bool repeat;  // copy uses the previous offset
};

// Token field extraction from the 8 stream bytes at the token start.
// Layouts: SPEC.md section 2; decode.go:195-308.
__device__ __forceinline__ Token parse_token(uint64_t w) {
    Token t;
    uint32_t lo = (uint32_t)w;
    uint32_t tag = lo & 3;
    t.repeat = false;
    t.off = 0;
    if (tag == 0) {
        uint32_t x = (lo >> 3) & 31;
        uint32_t len;
        if (x < 29) {
            t.hdr = 1;
            len = x + 1;
        } else {
            uint32_t nb = x - 28;  // 1..3 extra length bytes
            t.hdr = 1 + nb;
            len = ((lo >> 8) & (0xffffffu >> (8 * (3 - nb)))) + 30;
        }
        if (lo & 4) {
            t.repeat = true;
            t.lit = 0;
            t.mlen = len;
        } else {
            t.lit = len;
            t.mlen = 0;
        }
    } else if (tag == 1) {
        uint32_t len = (lo >> 2) & 15;
        t.off = ((lo & 0xffff) >> 6) + 1;
        t.lit = 0;
        if (len == 15) {
            t.hdr = 3;
            t.mlen = ((lo >> 16) & 0xff) + 18;
        } else {
            t.hdr = 2;
            t.mlen = len + 4;
        }
    } else if (tag == 2) {
        uint32_t len = (lo >> 2) & 63;
        t.off = ((lo >> 8) & 0xffff) + kMinCopy2Offset;
        t.lit = 0;
        if (len <= 60) {
            t.hdr = 3;
            t.mlen = len + 4;
        } else {
            uint32_t nb = len - 60;
            t.hdr = 3 + nb;
            t.mlen = ((uint32_t)(w >> 24) & (0xffffffu >> (8 * (3 - nb)))) + 64;
        }
    } else if ((lo & 4) == 0) {  // fused copy2
        t.lit = ((lo >> 3) & 3) + 1;
        t.mlen = 4 + ((lo >> 5) & 7);
        t.off = ((lo >> 8) & 0xffff) + kMinCopy2Offset;
        t.hdr = 3;
    } else {  // copy3
        t.lit = (lo >> 3) & 3;
        uint32_t len = (lo >> 5) & 63;
        t.off = (lo >> 11) + kMinCopy3Offset;
        if (len < 61) {
            t.hdr = 4;
            t.mlen = len + 4;
        } else {
            uint32_t nb = len - 60;
            t.hdr = 4 + nb;
            t.mlen = ((uint32_t)(w >> 32) & (0xffffffu >> (8 * (3 - nb)))) + 64;
        }
    }
    return t;
}

// 8 stream bytes at s (zero filled past slen); never reads a word that holds
// no byte of [0, slen).
__device__ __forceinline__ uint64_t ldg_window(const uint8_t *sp, int64_t s, int64_t slen) {
    if (s + 8 <= slen) return ldg_u64_unaligned(sp + s);
    uint64_t w = 0;
    for (int i = 0; i < 8 && s + i < slen; i++) w |= (uint64_t)sp[s + i] << (8 * i);
    return w;
}

// ---------------------------------------------------------------------------
// v0: one warp per block, token-serial.  Every lane parses the same token
// (uniform control flow, broadcast loads); the 32 lanes share the byte copies.
// ---------------------------------------------------------------------------
template <int kWarps>
__global__ void __launch_bounds__(kWarps * 32)
decode_warp_serial_kernel(int nblk, const uint8_t *__restrict__ src, const uint64_t *__restrict__ sbeg,
                          const uint64_t *__restrict__ send, uint8_t *dst, const uint64_t *__restrict__ dbeg,
                          const uint64_t *__restrict__ dend, int32_t *__restrict__ status) {
    const int lane = lane_id();
    const int blk = blockIdx.x * kWarps + (threadIdx.x >> 5);
    if (blk >= nblk) return;

    const uint8_t *sp = src + sbeg[blk];
    const int64_t slen = (int64_t)(send[blk] - sbeg[blk]);
    uint8_t *dp = dst + dbeg[blk];
    const int64_t dlen = (int64_t)(dend[blk] - dbeg[blk]);

    int64_t s = 0, d = 0;
    uint32_t offset = 1;
    bool bad = false;

    while (s < slen) {
        Token t = parse_token(ldg_window(sp, s, slen));
        if (s + t.hdr > slen) {
            bad = true;
            break;
        }
        s += t.hdr;
        if (t.lit) {
            if ((int64_t)t.lit > dlen - d || (int64_t)t.lit > slen - s) {
                bad = true;
                break;
            }
            for (uint32_t i = lane; i < t.lit; i += 32) dp[d + i] = sp[s + i];
            s += t.lit;
            d += t.lit;
        }
        if (t.mlen) {
            if (!t.repeat) offset = t.off;
            if ((int64_t)offset > d || (int64_t)t.mlen > dlen - d) {
                bad = true;
                break;
            }
            __syncwarp();  // earlier tokens' stores (other lanes) -> visible
            const uint8_t *from = dp + d - offset;
            if (offset >= 32) {
                const bool overlap = offset < t.mlen;
                for (uint32_t base = 0; base < t.mlen; base += 32) {
                    uint32_t i = base + lane;
                    if (i < t.mlen) dp[d + i] = from[i];
                    if (overlap) __syncwarp();
                }
            } else {
                // short period: replicate the `offset`-byte pattern
                uint32_t r = lane % offset;
                uint32_t step = 32 % offset;
                for (uint32_t base = 0; base < t.mlen; base += 32) {
                    uint32_t i = base + lane;
                    if (i < t.mlen) dp[d + i] = from[r];
                    r += step;
                    if (r >= offset) r -= offset;
                }
            }
            d += t.mlen;
        }
    }
    if (!bad && d != dlen) bad = true;
    if (lane == 0) status[blk] = bad ? 1 : 0;
}

}  // namespace mz
