// mz_validate.cuh -- validate mode: decode-after-encode on the device.
//
// Mirrors the reference's debugValidateBlocks switch (minlz.go:52; encode.go:108-133,
// writer.go:584-600): every block the encoder compressed is decoded again and compared
// with its source before the call returns.  The decode is the product decode kernel; the
// two kernels here only build its argument table and compare the result.
#pragma once

#include "mz_common.cuh"

namespace mz {

// Token-stream ranges of the encoder's output slots: [dbeg[i], dbeg[i] + out_len[i]).
__global__ void validate_ranges_kernel(int n, const uint64_t *__restrict__ dbeg, const uint32_t *__restrict__ out_len,
                                       uint64_t *__restrict__ tbeg, uint64_t *__restrict__ tend) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        tbeg[i] = dbeg[i];
        tend[i] = dbeg[i] + out_len[i];
    }
}

// bad[i] = 1 when block i was compressed (out_len > 0) and its decode failed or differs
// from the source; *first_bad = smallest such index (INT_MAX when none).  One CTA per block.
__global__ void __launch_bounds__(256) validate_compare_kernel(int n, const uint8_t *__restrict__ src,
                                                               const uint64_t *__restrict__ sbeg,
                                                               const uint64_t *__restrict__ send,
                                                               const uint8_t *__restrict__ dec,
                                                               const uint32_t *__restrict__ out_len,
                                                               const int32_t *__restrict__ status, int *first_bad) {
    for (int b = blockIdx.x; b < n; b += gridDim.x) {
        if (out_len[b] == 0) continue;  // stored by the caller: nothing to decode
        const uint8_t *s = src + sbeg[b];
        const uint8_t *d = dec + sbeg[b];
        const uint64_t m = send[b] - sbeg[b];
        int diff = status[b] != 0;
        for (uint64_t i = threadIdx.x; i < m && !diff; i += blockDim.x) diff |= s[i] != d[i];
        if (__syncthreads_or(diff) && threadIdx.x == 0) atomicMin(first_bad, b);
    }
}

}  // namespace mz
