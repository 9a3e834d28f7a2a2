// mz_decode.cuh -- shared helpers of the MinLZ block decode kernels (sm_100a).
//
// The decode kernel itself is mz_decode_pc.cuh (parser / copier).  The first
// version of this file held a one-warp-per-block, token-serial kernel (70 ms for
// 4096 x 1 MiB); it was removed once the parser / copier
// kernel passed the same parity suite.
#pragma once

#include "mz_common.cuh"

namespace mz {

// 8 stream bytes at s (zero filled past slen); never reads a word that holds
// no byte of [0, slen).
__device__ __forceinline__ uint64_t ldg_window(const uint8_t *sp, int64_t s, int64_t slen) {
    if (s + 8 <= slen) return ldg_u64_unaligned(sp + s);
    uint64_t w = 0;
    for (int i = 0; i < 8 && s + i < slen; i++) w |= (uint64_t)sp[s + i] << (8 * i);
    return w;
}

}  // namespace mz
