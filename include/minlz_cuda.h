/*
 * minlz_cuda.h -- C ABI of libminlz_cuda.so: the MinLZ block encode / decode
 * hot path as hand-written sm_100a CUDA kernels, batch-first.
 *
 * This is the drop-in boundary for the per-architecture seam of minio/minlz
 * (the functions that build tags select between asm_amd64.s and the pure-Go
 * path).  Each entry point names the reference interface it replaces; paths
 * are relative to the reference tree.  INTEGRATION.md shows the cgo binding.
 *
 * Conventions
 *   - plain pointers and sizes, no C++/torch types; thread-safe and re-entrant
 *     (any goroutine / OS thread); the library owns no caller memory.
 *   - a batch is a flat byte buffer plus an offset table of nblk+1 uint64
 *     (block i = [off[i], off[i+1])).  Flat + offsets rather than pointer
 *     arrays because cgo may not pass Go memory holding Go pointers.
 *   - "blocks" entry points work at the reference's internal seam: token
 *     streams WITHOUT the 0x00 + uvarint(len) block header
 *     (encode_amd64.go:111-118, decode_amd64.go:21).
 *   - return value 0 = ok, negative = MZCU_ERR_*; per-block results go to
 *     out_len[] / status[] exactly as the reference's int returns.
 *   - there is NO CPU fallback: without a CUDA device every compute entry
 *     point fails with MZCU_ERR_CUDA.
 */
#ifndef MINLZ_CUDA_H
#define MINLZ_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MZCU_ABI_VERSION 2

/* minlz.go:24 MaxBlockSize */
#define MZCU_MAX_BLOCK_SIZE (8 << 20)

/* encode.go:20-42 levels (only the levels on the hot path are accepted) */
#define MZCU_LEVEL_SUPERFAST (-1) /* encode_l0.go  encodeBlockFast */
#define MZCU_LEVEL_UNCOMPRESSED 0
#define MZCU_LEVEL_FASTEST 1  /* encode_l1.go  encodeBlock        */
#define MZCU_LEVEL_BALANCED 2 /* encode_l2.go  encodeBlockBetter  */

/* decode.go:29-40 error values */
#define MZCU_OK 0
#define MZCU_ERR_CORRUPT (-1)       /* ErrCorrupt      */
#define MZCU_ERR_TOO_LARGE (-2)     /* ErrTooLarge     */
#define MZCU_ERR_UNSUPPORTED (-3)   /* first byte != 0: Snappy/S2 fallback stays in host Go (decode.go:59-68) */
#define MZCU_ERR_INVALID_LEVEL (-4) /* ErrInvalidLevel */
#define MZCU_ERR_DST_TOO_SMALL (-5) /* C callers must size dst; Go would allocate */
#define MZCU_ERR_CUDA (-6)          /* CUDA runtime failure / no device; see mzcu_last_error */
#define MZCU_ERR_INVALID_ARG (-7)
#define MZCU_ERR_VALIDATE (-8)      /* validate mode: an encoded block did not decode back to its source */

/* decode.go:25-27 decodeErrCodeCorrupt: per-block status of the decode seam */
#define MZCU_BLOCK_OK 0
#define MZCU_BLOCK_CORRUPT 1

/* Encoder FLAVOUR.  The reference ships two implementations of its encode loops,
 * selected at build time: the pure-Go functions (encode_l0/l1/l2.go; `-tags noasm`,
 * purego, every platform without assembly -- asm_none.go:15) and the generated
 * assembly (asm_amd64.s from _generate/gen.go; encode_amd64.go:15).  README.md:375:
 * "Using assembly/non-assembly versions will often produce slightly different
 * output".  Both are valid MinLZ; they differ in margins, bail-out tests, match
 * extension at the block tail and per-size-class tables.  This library mirrors
 * both, byte for byte:
 *   MZCU_FLAVOR_GO     (default) == encodeBlockGo / encodeFastBlockGo / encodeBlockBetterGo
 *   MZCU_FLAVOR_AMD64            == encodeFastBlockAsm* / encodeBlockAsm* / encodeBetterBlockAsm*
 *                                   (what `go build` produces on amd64), every size class.
 * The setting is process-wide, like the build tag it stands for.  Decoding is
 * unaffected (decodeBlockAsm == minLZDecodeGo by the reference's own tests). */
#define MZCU_FLAVOR_GO 0
#define MZCU_FLAVOR_AMD64 1
int mzcu_set_encoder_flavor(int flavor);
int mzcu_get_encoder_flavor(void);

int mzcu_abi_version(void);
/* Thread-local message of the last failing call on this thread. */
const char *mzcu_last_error(void);
/* Number of visible CUDA devices (0 when there is none); never fails. */
int mzcu_device_count(void);

/* ---- pure helpers (no device work) ------------------------------------ */

/* encode.go:234-244 MaxEncodedLen: n+2, 1 for n==0, -1 for n > 8 MiB. */
int64_t mzcu_max_encoded_len(int64_t src_len);
/* decode.go:107 DecodedLen / decode.go:120 isMinLZ on a full block.
 * Returns the decoded length or MZCU_ERR_*. */
int64_t mzcu_decoded_len(const uint8_t *block, size_t n);
/* decode.go:114 IsMinLZ: *is_minlz=1 when the block is MinLZ (first byte 0),
 * *size = decoded size.  Returns MZCU_OK or MZCU_ERR_*. */
int mzcu_is_minlz(const uint8_t *block, size_t n, int *is_minlz, int64_t *size);

/* ---- seam level, device pointers (what the benchmarks time) ------------
 *
 * replaces: encodeBlock / encodeBlockBetter (encode_amd64.go:119,201;
 *           asm_none.go:51,68) applied to every block of a batch.
 *   src/src_off : device; uncompressed blocks, 16 <= len <= 8 MiB each
 *                 (shorter blocks yield out_len 0, as the reference).
 *   dst/dst_off : device; per-block capacity >= MaxEncodedLen(len).
 *   out_len     : device uint32[nblk]; bytes written, 0 = not compressible
 *                 (caller stores the block raw, encode.go:137-138).
 *   stream      : cudaStream_t (as void*); the call is asynchronous.
 */
int mzcu_encode_blocks_dev(int device, int level, int nblk, const uint8_t *src, const uint64_t *src_off,
                           uint8_t *dst, const uint64_t *dst_off, uint32_t *out_len, void *stream);

/* replaces: minLZDecode (decode_amd64.go:21 / decode_other.go:23 /
 *           decode.go:178) applied to every block of a batch.
 *   src/src_off : device; token streams.
 *   dst/dst_off : device; dst_off[i+1]-dst_off[i] must equal the decoded
 *                 length of block i (len(dst) in the reference).
 *   status      : device int32[nblk]; 0 ok / 1 corrupt.  A corrupt block
 *                 never writes outside its own dst range.
 */
int mzcu_decode_blocks_dev(int device, int nblk, const uint8_t *src, const uint64_t *src_off, uint8_t *dst,
                           const uint64_t *dst_off, int32_t *status, void *stream);

/* Packs encoder output into a dense stream: block i (len[i] bytes at
 * src + src_off[i]) is copied to dst + dst_off[i], where dst_off (device
 * uint64[nblk+1], written by this call) is the exclusive prefix sum of len.
 * The batch analogue of the ordered result hand-off in writer.go:214-272; the
 * result is directly consumable by mzcu_decode_blocks_dev.  dst must hold
 * sum(len) bytes (<= sum of capacities). */
int mzcu_pack_blocks_dev(int device, int nblk, const uint8_t *src, const uint64_t *src_off, const uint32_t *len,
                         uint8_t *dst, uint64_t *dst_off, void *stream);

/* ---- seam level, host pointers (synchronous; H2D + kernel + D2H) -------
 * Same contracts with host memory; pinned memory (mzcu_host_alloc) makes the
 * copies asynchronous-capable and ~2x faster.  device < 0 = current device. */
int mzcu_encode_blocks(int device, int level, int nblk, const uint8_t *src, const uint64_t *src_off, uint8_t *dst,
                       const uint64_t *dst_off, uint32_t *out_len);
int mzcu_decode_blocks(int device, int nblk, const uint8_t *src, const uint64_t *src_off, uint8_t *dst,
                       const uint64_t *dst_off, int32_t *status);
/* Like mzcu_encode_blocks, but the token streams are written back to back:
 * block i occupies dst[dst_off_out[i] .. dst_off_out[i+1]) (dst_off_out has
 * nblk+1 entries and is an OUTPUT); an empty range means "not compressible".
 * One D2H copy for the whole batch; the result feeds mzcu_decode_blocks as is.
 * dst_cap >= sum of source lengths is always enough. */
int mzcu_encode_blocks_packed(int device, int level, int nblk, const uint8_t *src, const uint64_t *src_off,
                              uint8_t *dst, size_t dst_cap, uint64_t *dst_off_out);

/* ---- stream-layer helpers ------------------------------
 * The stream format frames every block as a chunk that carries the masked
 * CRC-32C of the UNCOMPRESSED block (minlz.go:133-140 crc; writer.go:672,
 * reader.go:341-351; SPEC.md stream section 3).  The blocks are resident on
 * the device for the codec kernels anyway, so the checksum is computed there.
 *
 * mzcu_crc32c_blocks[_dev]: crc[i] = masked CRC-32C of block i.
 * mzcu_stream_encode_blocks: mzcu_encode_blocks_packed + crc_out[i] of the
 *   source block (what Writer.write needs per block, writer.go:670-696).
 * mzcu_stream_decode_blocks: mzcu_decode_blocks + crc_out[i] of the decoded
 *   block (what Reader.Read checks, reader.go:334-351).  crc_out may be NULL. */
/* Any block below 4 GiB (the length-combine tables cover every 32-bit length; stream chunks are
 * below 16 MiB); the host form rejects larger blocks with MZCU_ERR_TOO_LARGE. */
int mzcu_crc32c_blocks_dev(int device, int nblk, const uint8_t *src, const uint64_t *src_off, uint32_t *crc, void *stream);
int mzcu_crc32c_blocks(int device, int nblk, const uint8_t *src, const uint64_t *src_off, uint32_t *crc);
int mzcu_stream_encode_blocks(int device, int level, int nblk, const uint8_t *src, const uint64_t *src_off, uint8_t *dst,
                              size_t dst_cap, uint64_t *dst_off_out, uint32_t *crc_out);
int mzcu_stream_decode_blocks(int device, int nblk, const uint8_t *src, const uint64_t *src_off, uint8_t *dst,
                              const uint64_t *dst_off, int32_t *status, uint32_t *crc_out);

/* ---- block API level (full blocks with header), host pointers ----------
 *
 * replaces: Encode (encode.go:74-139).  Writes 0x00 + uvarint(len) + tokens,
 * or the stored form 00 00 raw when len < 16 / level 0 / incompressible.
 * dst_cap must be >= mzcu_max_encoded_len(n).  Returns the encoded length. */
int64_t mzcu_encode(uint8_t *dst, size_t dst_cap, const uint8_t *src, size_t n, int level);
/* replaces: TryEncode (encode.go:168-207): 0 when Go returns nil. */
int64_t mzcu_try_encode(uint8_t *dst, size_t dst_cap, const uint8_t *src, size_t n, int level);
/* replaces: Decode (decode.go:50-78) for MinLZ blocks.  Returns the decoded
 * length; on MZCU_ERR_CORRUPT dst holds the partial output like Go's
 * `return dst, ErrCorrupt`. */
int64_t mzcu_decode(uint8_t *dst, size_t dst_cap, const uint8_t *block, size_t n);

/* Batch forms of Encode / Decode over full blocks (header included in dst /
 * src).  enc_len[i] receives the encoded size of block i (always > 0),
 * dec_len[i] the decoded size or a negative MZCU_ERR_*. */
int mzcu_encode_batch(int device, int level, int nblk, const uint8_t *src, const uint64_t *src_off, uint8_t *dst,
                      const uint64_t *dst_off, uint64_t *enc_len);
int mzcu_decode_batch(int device, int nblk, const uint8_t *src, const uint64_t *src_off, uint8_t *dst,
                      const uint64_t *dst_off, int64_t *dec_len);

/* ---- several devices behind one call ------------------------------------
 * replaces: the goroutine fan-out of Writer.EncodeBuffer (writer.go:441-563, one goroutine per
 * block at :501) and Reader.DecodeConcurrent (reader.go:575-992, :830-859) for a host that owns
 * several GPUs.  One batch is sharded over `devices[0..ndev)` in stream order -- device k takes
 * blocks [nblk*k/ndev, nblk*(k+1)/ndev) -- one library thread per device, each bound to its
 * device's NUMA node; blocks are independent, so nothing is exchanged between devices.
 *
 * encode: block i's token stream lands at dst[dst_off_out[i] .. + out_len[i]) (both arrays have
 *   nblk entries and are OUTPUTS; out_len 0 = not compressible; the ranges of different devices
 *   are not adjacent).  dst_cap must be >= the total source bytes.  crc_out (nullable) =
 *   masked CRC-32C of every source block.
 * decode: block i's token stream is src[src_beg[i] .. + src_len[i]) -- the layout the encode call
 *   produced -- and decodes into dst[dst_off[i] .. dst_off[i+1]). */
int mzcu_stream_encode_blocks_multi(int ndev, const int *devices, int level, int nblk, const uint8_t *src,
                                    const uint64_t *src_off, uint8_t *dst, size_t dst_cap, uint64_t *dst_off_out,
                                    uint32_t *out_len, uint32_t *crc_out);
int mzcu_stream_decode_blocks_multi(int ndev, const int *devices, int nblk, const uint8_t *src, const uint64_t *src_beg,
                                    const uint32_t *src_len, uint8_t *dst, const uint64_t *dst_off, int32_t *status,
                                    uint32_t *crc_out);

/* ---- asynchronous host calls ----------------------------------------------
 * The reference's Writer keeps `concurrency` blocks in flight (writer.go:214-272: results are
 * handed over in order while later blocks still encode).  The batch analogue: submit returns a
 * job id (> 0) at once; the call runs on a library thread with its own workspace and streams, so
 * the D2H of batch k overlaps the H2D and kernels of batch k+1.  mzcu_wait blocks, returns the
 * call's result and retires the job.  All buffers must stay valid until mzcu_wait returns.
 * Jobs reach the device in submission order (their kernels are launched in that order), so a host
 * decides the overlap by the order it submits in; three jobs in flight - encode k+1 and k+2
 * submitted before decode k - keep the device and both PCIe directions busy. */
int64_t mzcu_submit_stream_encode_blocks(int device, int level, int nblk, const uint8_t *src, const uint64_t *src_off,
                                         uint8_t *dst, size_t dst_cap, uint64_t *dst_off_out, uint32_t *crc_out);
int64_t mzcu_submit_stream_decode_blocks(int device, int nblk, const uint8_t *src, const uint64_t *src_off, uint8_t *dst,
                                         const uint64_t *dst_off, int32_t *status, uint32_t *crc_out);
int mzcu_wait(int64_t job);

/* ---- validate mode ----------------------------------------------------------
 * replaces: debugValidateBlocks (minlz.go:52; encode.go:108-133, writer.go:584-600): every block
 * the encoder compressed is decoded again ON THE DEVICE and compared with its source before an
 * encode call returns; a mismatch fails the call with MZCU_ERR_VALIDATE and names the block in
 * mzcu_last_error.  Off by default; MZCU_VALIDATE=1 in the environment turns it on at start. */
int mzcu_set_validate(int on);
int mzcu_get_validate(void);
/* The validate pass on its own (device pointers, the layout mzcu_encode_blocks_dev produced):
 * MZCU_OK, or MZCU_ERR_VALIDATE naming the first block that does not decode back to its source.
 * Synchronises `stream`. */
int mzcu_validate_blocks_dev(int device, int nblk, const uint8_t *src, const uint64_t *src_off, const uint8_t *enc,
                             const uint64_t *enc_off, const uint32_t *out_len, void *stream);

/* ---- host placement ---------------------------------------------------------
 * Binds the calling thread to the CPUs of the NUMA node `device` is attached to and prefers that
 * node for its later allocations (staging buffers).  Returns the node, or -1 when there is
 * nothing to bind to (single-node host, no affinity information); never fails the caller. */
int mzcu_bind_host_to_device(int device);

/* ---- pinned host memory for callers that want fast H2D / D2H ---------- */
void *mzcu_host_alloc(size_t n);
void mzcu_host_free(void *p);

/* Last kernel duration in milliseconds measured with CUDA events around the
 * kernels of the most recent host-pointer call on this thread (0 if none). */
float mzcu_last_kernel_ms(void);

#ifdef __cplusplus
}
#endif
#endif /* MINLZ_CUDA_H */
