/*
 * minlz_oracle.h -- CPU restatement of the MinLZ block codec hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the parity oracle for the CUDA kernels in
 * minlz_b200/csrc.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product path
 * (libminlz_cuda.so) never links, loads or calls anything in this directory.
 *
 * What it restates (reference = minio/minlz, pure-Go "noasm" path):
 *   decoder   decode.go:178-622  (minLZDecodeGo), cross-read with
 *             internal/reference/decoder.go:26-373
 *   L0        encode_l0.go:32-279 (encodeFastBlockGo), :281-522 (...Go64K)
 *   L1        encode_l1.go:39-283 (encodeBlockGo), :285-524 (encodeBlockGo64K)
 *   L2        encode_l2.go:61-338 (encodeBlockBetterGo), :343-596 (...Go64K)
 *   emitters  asm_none.go:84-323, encode.go:247-282
 *   wrappers  encode.go:74-139,168-244  decode.go:50-171  asm_none.go:51-76
 *   crc       minlz.go:133-140
 *
 *   amd64     _generate/gen.go:257-1155 (genEncodeBlockAsm), :1171-2038
 *   flavour   (genEncodeBetterBlockAsm), class tables gen.go:57-88,
 *             dispatch encode_amd64.go:37-271
 *
 * Pinning: PINNED against the reference itself.  oracle/_ref (p9_to_gas.py +
 * ref_shim.c) runs the reference's own asm_amd64.s here; tests/test_ref_asm.py
 * requires the decoder, the emitters, matchLen and the amd64-flavour encoders
 * (all levels, every size class) to agree with it byte for byte.  The decoder is
 * also pinned to testdata/Mark.Twain-Tom.Sawyer.txt.mzb -> .txt and the emitters
 * to TestEmitLiteral / TestEmitCopy (minlz_test.go:871-1026).  The Go-flavour
 * encoders (Go source, nothing to run here) are anchored through the real
 * decoder and through the tail-only difference to the assembly for blocks
 * > 512 KiB.
 */
#ifndef MINLZ_ORACLE_H
#define MINLZ_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MZO_MAX_BLOCK_SIZE (8 << 20) /* minlz.go:24 */

/* error codes of the wrapper functions (decode.go:29-40) */
#define MZO_OK 0
#define MZO_ERR_CORRUPT (-1)
#define MZO_ERR_TOO_LARGE (-2)
#define MZO_ERR_UNSUPPORTED (-3) /* first byte != 0: Snappy/S2 fallback is host Go, not here */
#define MZO_ERR_INVALID_LEVEL (-4)
#define MZO_ERR_DST_TOO_SMALL (-5)

/* encode.go:234-244 */
int64_t mzo_max_encoded_len(int64_t src_len);

/* asm_none.go:84-323, encode.go:247-282: return bytes written to dst. */
int mzo_emit_literal(uint8_t *dst, const uint8_t *lit, size_t n);
int mzo_emit_repeat(uint8_t *dst, int length);
int mzo_emit_copy(uint8_t *dst, int offset, int length);
int mzo_emit_copy_lits2(uint8_t *dst, const uint8_t *lits, int nlits, int offset, int length);
int mzo_emit_copy_lits3(uint8_t *dst, const uint8_t *lits, int nlits, int offset, int length);

/* asm_none.go:51-76 dispatch: returns bytes written, 0 = not compressible.
 * dst must hold mzo_max_encoded_len(n) bytes; header is NOT written. */
int64_t mzo_encode_block_l0(uint8_t *dst, const uint8_t *src, size_t n); /* LevelSuperFast, encode_l0.go */
int64_t mzo_encode_block_l1(uint8_t *dst, const uint8_t *src, size_t n);
int64_t mzo_encode_block_l2(uint8_t *dst, const uint8_t *src, size_t n);
/* The same seam as built for amd64 (encode_amd64.go:37-189 -> the generated
 * assembly of _generate/gen.go:257-1155), restated; byte-identical to the real
 * assembly run through oracle/_ref (tests/test_ref_asm.py). */
int64_t mzo_encode_block_l0_asm(uint8_t *dst, const uint8_t *src, size_t n);
int64_t mzo_encode_block_l1_asm(uint8_t *dst, const uint8_t *src, size_t n);
int64_t mzo_encode_block_l2_asm(uint8_t *dst, const uint8_t *src, size_t n);

/* decode.go:178 minLZDecodeGo: dst_len must equal the decoded length, src is
 * the token stream without 0x00 + uvarint.  Returns 0 ok / 1 corrupt. */
int mzo_decode_block(uint8_t *dst, size_t dst_len, const uint8_t *src, size_t src_len);

/* encode.go:74 Encode (levels 0,1,2).  Returns encoded length or MZO_ERR_*. */
int64_t mzo_encode(uint8_t *dst, size_t dst_cap, const uint8_t *src, size_t n, int level);
/* encode.go:168 TryEncode: encoded length, 0 if "nil" (incompressible/invalid). */
int64_t mzo_try_encode(uint8_t *dst, size_t dst_cap, const uint8_t *src, size_t n, int level);

/* decode.go:120 isMinLZ.  Fills *is_mlz,*lits,*hdr (offset of block payload),
 * *size.  Returns MZO_OK or error. */
int mzo_is_minlz(const uint8_t *src, size_t n, int *is_mlz, int *lits, size_t *hdr, int64_t *size);
/* decode.go:107 DecodedLen */
int64_t mzo_decoded_len(const uint8_t *src, size_t n);
/* decode.go:50 Decode.  Returns decoded length or MZO_ERR_*.  On
 * MZO_ERR_CORRUPT from the token loop dst holds the partial output. */
int64_t mzo_decode(uint8_t *dst, size_t dst_cap, const uint8_t *src, size_t n);

/* minlz.go:137-140 masked CRC32C (Castagnoli). */
uint32_t mzo_crc(const uint8_t *b, size_t n);

/* Multi-threaded batch drivers for the CPU baseline leg of bench.py:
 * one block per task over nthreads pthreads.  Offsets are nblk+1 entries. */
int mzo_encode_batch_mt(int level, int nblk, const uint8_t *src, const uint64_t *src_off,
                        uint8_t *dst, const uint64_t *dst_off, uint32_t *out_len, int nthreads);
int mzo_decode_batch_mt(int nblk, const uint8_t *src, const uint64_t *src_off, uint8_t *dst,
                        const uint64_t *dst_off, int32_t *status, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
