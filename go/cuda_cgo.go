//go:build cuda && cgo

// Thin cgo layer over libminlz_cuda.so (include/minlz_cuda.h).  This file and
// its siblings are what a maintainer adds to minio/minlz to select the CUDA
// backend with `-tags cuda`; they contain no codec logic.  They could not be
// compiled in the build container (no Go toolchain there); see INTEGRATION.md.

package minlz

/*
#cgo CFLAGS: -I${SRCDIR}/include
#cgo LDFLAGS: -L${SRCDIR}/lib -lminlz_cuda
#include <stdlib.h>
#include "minlz_cuda.h"
*/
import "C"

import (
	"errors"
	"unsafe"
)

var errCuda = errors.New("minlz: cuda backend failure")

func cudaErr(rc C.int) error {
	switch rc {
	case C.MZCU_OK:
		return nil
	case C.MZCU_ERR_CORRUPT:
		return ErrCorrupt
	case C.MZCU_ERR_TOO_LARGE:
		return ErrTooLarge
	case C.MZCU_ERR_INVALID_LEVEL:
		return ErrInvalidLevel
	case C.MZCU_ERR_UNSUPPORTED:
		return ErrUnsupported
	default:
		return errors.Join(errCuda, errors.New(C.GoString(C.mzcu_last_error())))
	}
}

// Encoder flavours: which of the reference's two builds the CUDA encoders mirror
// byte for byte (minlz_cuda.h).  An amd64 deployment that wants the bytes it has
// today calls SetEncoderFlavor(FlavorAMD64) once at start-up.
const (
	FlavorGo    = int(C.MZCU_FLAVOR_GO)    // encodeBlockGo / encodeBlockBetterGo / encodeFastBlockGo (noasm, purego)
	FlavorAMD64 = int(C.MZCU_FLAVOR_AMD64) // encodeBlockAsm* / encodeBetterBlockAsm* / encodeFastBlockAsm*
)

// SetEncoderFlavor is process-wide, like the build tag it stands for.
func SetEncoderFlavor(f int) error { return cudaErr(C.mzcu_set_encoder_flavor(C.int(f))) }

func bptr(b []byte) *C.uint8_t {
	if len(b) == 0 {
		return nil
	}
	return (*C.uint8_t)(unsafe.Pointer(unsafe.SliceData(b)))
}

// EncodeBlocks runs encodeBlock / encodeBlockBetter over a batch with one GPU
// launch.  src holds the blocks back to back, srcOff has len(blocks)+1 entries.
// The token streams come back packed in dst with their offsets; an empty range
// means "not compressible" exactly like a 0 return of encodeBlock.
func EncodeBlocks(dst, src []byte, srcOff []uint64, level int) (dstOff []uint64, err error) {
	n := len(srcOff) - 1
	dstOff = make([]uint64, n+1)
	rc := C.mzcu_encode_blocks_packed(-1, C.int(level), C.int(n), bptr(src),
		(*C.uint64_t)(unsafe.Pointer(unsafe.SliceData(srcOff))), bptr(dst), C.size_t(len(dst)),
		(*C.uint64_t)(unsafe.Pointer(unsafe.SliceData(dstOff))))
	return dstOff, cudaErr(rc)
}

// DecodeBlocks runs minLZDecode over a batch with one GPU launch; status[i] is
// the reference's return code (0 ok, 1 = decodeErrCodeCorrupt).
func DecodeBlocks(dst []byte, dstOff []uint64, src []byte, srcOff []uint64) (status []int32, err error) {
	n := len(srcOff) - 1
	status = make([]int32, n)
	rc := C.mzcu_decode_blocks(-1, C.int(n), bptr(src),
		(*C.uint64_t)(unsafe.Pointer(unsafe.SliceData(srcOff))), bptr(dst),
		(*C.uint64_t)(unsafe.Pointer(unsafe.SliceData(dstOff))),
		(*C.int32_t)(unsafe.Pointer(unsafe.SliceData(status))))
	return status, cudaErr(rc)
}
