//go:build cuda && cgo

// CUDA build of minLZDecode (replaces decode_amd64.go / decode_arm64.go /
// decode_other.go, which get `&& !cuda`).  Contract as decode.go:173-177:
// len(dst) is the exact decoded length, src excludes 0x00 + uvarint; returns 0
// or decodeErrCodeCorrupt.  Batch callers use DecodeBlocks (reader_cuda.go).

package minlz

func minLZDecode(dst, src []byte) int {
	if dst == nil {
		panic("minlz: nil dst") // decode_amd64.go:22-24
	}
	st, _, err := DecodeBlocks(dst, []uint64{0, uint64(len(dst))}, src, []uint64{0}, []uint32{uint32(len(src))})
	if err != nil {
		panic(err) // no CPU decoder in this build
	}
	return int(st[0])
}
