//go:build cuda && cgo

// CUDA build of the per-architecture seam.  Replaces encode_amd64.go /
// asm_none.go (add `&& !cuda` to their build constraints) by providing the
// same unexported functions with the same contracts
// (encode_amd64.go:111-118): dst has MaxEncodedLen(len(src)) capacity, the
// block header is already written, the return value is the number of bytes
// written or 0 for "not compressible".

package minlz

func encodeOne(dst, src []byte, level int) int {
	if len(src) < minNonLiteralBlockSize {
		return 0
	}
	off, err := EncodeBlocks(dst, src, []uint64{0, uint64(len(src))}, level)
	if err != nil {
		panic(err) // the reference seam has no error path; a device failure is fatal
	}
	return int(off[1])
}

func encodeBlock(dst, src []byte) (d int)       { return encodeOne(dst, src, LevelFastest) }
func encodeBlockBetter(dst, src []byte) (d int) { return encodeOne(dst, src, LevelBalanced) }

func encodeBlockFast(dst, src []byte) (d int)   { return encodeOne(dst, src, LevelSuperFast) }

// LevelSmallest is not on the accelerated path; it keeps the pure-Go
// implementation (encode_l3.go).
