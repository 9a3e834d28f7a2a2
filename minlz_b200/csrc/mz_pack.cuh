// mz_pack.cuh -- packs the per-block encoder output (one slot of
// MaxEncodedLen capacity per block) into a dense stream plus an offset table,
// so that a batch leaves the device with one D2H copy and can be handed to the
// decoder as is.  Plays the role of the ordered writer goroutine of the
// reference's stream layer (writer.go:214-272) for a batch.
#pragma once

#include "mz_common.cuh"

namespace mz {

// Exclusive prefix sum of len[0..n) into off[0..n] (uint64), one CTA.
__global__ void __launch_bounds__(1024) scan_lengths_kernel(int n, const uint32_t *__restrict__ len,
                                                            uint64_t *__restrict__ off) {
    __shared__ uint64_t warp_sum[32];
    __shared__ uint64_t carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        int i = base + threadIdx.x;
        uint64_t v = i < n ? len[i] : 0;
        uint64_t x = v;
        for (int d = 1; d < 32; d <<= 1) {
            uint64_t y = __shfl_up_sync(kFullMask, x, d);
            if (lane >= d) x += y;
        }
        if (lane == 31) warp_sum[warp] = x;
        __syncthreads();
        if (warp == 0) {
            uint64_t w = warp_sum[lane];
            for (int d = 1; d < 32; d <<= 1) {
                uint64_t y = __shfl_up_sync(kFullMask, w, d);
                if (lane >= d) w += y;
            }
            warp_sum[lane] = w;
        }
        __syncthreads();
        uint64_t carry = carry_s;
        uint64_t incl = x + (warp ? warp_sum[warp - 1] : 0) + carry;
        if (i < n) off[i] = incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) off[n] = carry_s;
}

// dst[off[b] .. off[b]+len[b]) = src[sbeg[b] .. +len[b]); one CTA per block.
__global__ void __launch_bounds__(256) pack_blocks_kernel(int n, const uint8_t *__restrict__ src,
                                                          const uint64_t *__restrict__ sbeg,
                                                          const uint32_t *__restrict__ len, uint8_t *__restrict__ dst,
                                                          const uint64_t *__restrict__ off) {
    for (int b = blockIdx.x; b < n; b += gridDim.x) {
        const uint8_t *s = src + sbeg[b];
        uint8_t *d = dst + off[b];
        const uint32_t m = len[b];
        // head bytes until d is 16-byte aligned
        uint32_t head = (uint32_t)((16 - (reinterpret_cast<uintptr_t>(d) & 15)) & 15);
        if (head > m) head = m;
        if (threadIdx.x < head) d[threadIdx.x] = s[threadIdx.x];
        const uint32_t body = (m - head) / 16;
        const uint8_t *sb = s + head;
        uint4 *db = reinterpret_cast<uint4 *>(d + head);
        const unsigned mis = (unsigned)(reinterpret_cast<uintptr_t>(sb) & 3);
        if (mis == 0 && (reinterpret_cast<uintptr_t>(sb) & 15) == 0) {
            const uint4 *s4 = reinterpret_cast<const uint4 *>(sb);
            for (uint32_t i = threadIdx.x; i < body; i += blockDim.x) db[i] = s4[i];
        } else {
            const uint32_t *sw = reinterpret_cast<const uint32_t *>(sb - mis);
            const unsigned sh = mis * 8;
            for (uint32_t i = threadIdx.x; i < body; i += blockDim.x) {
                uint32_t w0 = sw[4 * i], w1 = sw[4 * i + 1], w2 = sw[4 * i + 2], w3 = sw[4 * i + 3];
                uint4 v;
                if (mis == 0) {
                    v = make_uint4(w0, w1, w2, w3);
                } else {
                    uint32_t w4 = sw[4 * i + 4];
                    v = make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh),
                                   __funnelshift_r(w3, w4, sh));
                }
                db[i] = v;
            }
        }
        const uint32_t done = head + body * 16;
        if (done + threadIdx.x < m) d[done + threadIdx.x] = s[done + threadIdx.x];
    }
}

}  // namespace mz
