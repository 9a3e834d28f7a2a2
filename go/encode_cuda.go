//go:build cuda && cgo

// CUDA build of the per-architecture seam.  Replaces encode_amd64.go /
// asm_none.go (add `&& !cuda` to their build constraints) by providing the
// same unexported functions with the same contracts
// (encode_amd64.go:111-118): dst has MaxEncodedLen(len(src)) capacity, the
// block header is already written, the return value is the number of bytes
// written or 0 for "not compressible".
//
// One block per call is the worst way to use a GPU (a launch plus two PCIe hops
// per block); these exist so that every caller of the seam keeps working.  The
// batch points -- Writer.EncodeBuffer / ReadFrom and Reader.DecodeConcurrent --
// are rerouted in writer_cuda.go / reader_cuda.go.

package minlz

func encodeOne(dst, src []byte, level int) int {
	if len(src) < minNonLiteralBlockSize {
		return 0
	}
	off, n, _, err := EncodeBlocks(dst, src, []uint64{0, uint64(len(src))}, level)
	if err != nil {
		// The reference seam has no error path and this build has no CPU encoder to
		// fall back to: a device failure is fatal, loudly.
		panic(err)
	}
	if n[0] > 0 && off[0] != 0 {
		copy(dst, dst[off[0]:off[0]+uint64(n[0])])
	}
	return int(n[0])
}

func encodeBlock(dst, src []byte) (d int)       { return encodeOne(dst, src, LevelFastest) }
func encodeBlockBetter(dst, src []byte) (d int) { return encodeOne(dst, src, LevelBalanced) }
func encodeBlockFast(dst, src []byte) (d int)   { return encodeOne(dst, src, LevelSuperFast) }

// LevelSmallest is not on the accelerated path; it keeps the pure-Go
// implementation (encode_l3.go).
