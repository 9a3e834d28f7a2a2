// mz_crc32c.cuh -- masked CRC-32C of every block of a batch (sm_100a).
//
// The stream format stores crc(uncompressed block) in every chunk
// (reference minlz.go:133-140 `crc`, SPEC.md stream section 3; computed at
// writer.go:672 and checked at reader.go:341-351).  On the GPU the blocks are
// already resident for the encode / decode kernels, so the checksum is one
// more pass over them: one warp per block, each lane folds a contiguous 1/32
// of the block with the byte-wise table, and the 32 partial registers are
// combined with precomputed "advance by 2^k zero bytes" GF(2) matrices (the
// CRC register update is linear, so crc(A||B) = advance(crc(A), |B|) ^ crc0(B)).
#pragma once

#include "mz_common.cuh"

namespace mz {

constexpr int kCrcWarps = 8;

struct CrcTables {
    uint32_t byte_table[256];  // reflected Castagnoli table
    uint32_t zeros[32][32];    // zeros[k][i]: image of register bit i after 2^k zero bytes (any 32-bit length)
};

__device__ __forceinline__ uint32_t crc_apply(const uint32_t *m, uint32_t v) {
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 32; i++) r ^= (v >> i) & 1 ? m[i] : 0;
    return r;
}

__global__ void __launch_bounds__(kCrcWarps * 32)
crc32c_blocks_kernel(int nblk, const uint8_t *__restrict__ src, const uint64_t *__restrict__ sbeg,
                     const uint64_t *__restrict__ send, uint32_t *__restrict__ out, const CrcTables *__restrict__ tabs) {
    __shared__ uint32_t T[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) T[i] = tabs->byte_table[i];
    __syncthreads();
    const int lane = lane_id();
    const int blk = blockIdx.x * kCrcWarps + (threadIdx.x >> 5);
    if (blk >= nblk) return;
    const uint8_t *p = src + sbeg[blk];
    const uint32_t n = (uint32_t)(send[blk] - sbeg[blk]);
    const uint32_t seg = (n + 31) / 32;
    const uint32_t lo = min(n, lane * seg), hi = min(n, lo + seg);
    uint32_t c = lane == 0 ? 0xffffffffu : 0u;
    uint32_t i = lo;
    // head bytes up to a 4-byte aligned address, then words, then the tail
    while (i < hi && ((reinterpret_cast<uintptr_t>(p + i)) & 3)) {
        c = T[(c ^ p[i]) & 0xff] ^ (c >> 8);
        i++;
    }
    for (; i + 4 <= hi; i += 4) {
        uint32_t w = *reinterpret_cast<const uint32_t *>(p + i) ^ c;
        c = T[w & 0xff] ^ (w >> 8);
        c = T[c & 0xff] ^ (c >> 8);
        c = T[c & 0xff] ^ (c >> 8);
        c = T[c & 0xff] ^ (c >> 8);
    }
    for (; i < hi; i++) c = T[(c ^ p[i]) & 0xff] ^ (c >> 8);
    // advance my register over the bytes that follow my segment
    uint32_t tail = n - hi;
    for (int k = 0; tail != 0; k++, tail >>= 1)
        if (tail & 1) c = crc_apply(tabs->zeros[k], c);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c ^= __shfl_xor_sync(kFullMask, c, o);
    if (lane == 0) {
        c = ~c;
        out[blk] = (c >> 15 | c << 17) + 0xa282ead8u;  // minlz.go:137-140
    }
}

}  // namespace mz
