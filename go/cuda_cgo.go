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
	"runtime"
	"sync"
	"unsafe"
)

// cudaEnabled gates the batch-point branches in writer.go / reader.go (a `const false` twin lives in a !cuda file).
const cudaEnabled = true

var errCuda = errors.New("minlz: cuda backend failure")

// call runs one C entry point and, on failure, fetches its message.  The message is
// thread-local in the library, so the goroutine stays on its OS thread for both calls.
func call(f func() C.int) error {
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	rc := f()
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
func SetEncoderFlavor(f int) error {
	return call(func() C.int { return C.mzcu_set_encoder_flavor(C.int(f)) })
}

// SetValidate turns on decode-after-encode on the device (debugValidateBlocks, minlz.go:52).
func SetValidate(on bool) {
	v := C.int(0)
	if on {
		v = 1
	}
	C.mzcu_set_validate(v)
}

// cudaDevices is the device list every batch call shards over (stream order =
// list order).  Default: device 0.  SetDevices([]int{0,1,...,7}) uses a whole box.
var (
	devMu       sync.RWMutex
	cudaDevices = []C.int{0}
)

func SetDevices(devs []int) {
	d := make([]C.int, len(devs))
	for i, v := range devs {
		d[i] = C.int(v)
	}
	devMu.Lock()
	cudaDevices = d
	devMu.Unlock()
}

func devices() []C.int {
	devMu.RLock()
	defer devMu.RUnlock()
	return cudaDevices
}

func bptr(b []byte) *C.uint8_t {
	if len(b) == 0 {
		return nil
	}
	return (*C.uint8_t)(unsafe.Pointer(unsafe.SliceData(b)))
}

func u64ptr(b []uint64) *C.uint64_t { return (*C.uint64_t)(unsafe.Pointer(unsafe.SliceData(b))) }
func u32ptr(b []uint32) *C.uint32_t { return (*C.uint32_t)(unsafe.Pointer(unsafe.SliceData(b))) }

// ---- pinned staging buffers (mzcu_host_alloc), pooled by capacity class ----
//
// H2D / D2H from pageable Go memory is staged by the driver at half the speed;
// batch callers copy their blocks into one of these instead (they are C memory,
// so cgo's pointer rules do not apply to them either).
type pinned struct {
	b []byte
}

var pinnedPool sync.Map // capacity class (log2) -> *sync.Pool

func getPinned(n int) *pinned {
	cls := 20 // 1 MiB minimum
	for (1 << cls) < n {
		cls++
	}
	p, _ := pinnedPool.LoadOrStore(cls, &sync.Pool{})
	if v := p.(*sync.Pool).Get(); v != nil {
		return v.(*pinned)
	}
	ptr := C.mzcu_host_alloc(C.size_t(1 << cls))
	if ptr == nil {
		panic("minlz: mzcu_host_alloc failed") // out of pinned memory: there is no pageable fallback
	}
	return &pinned{b: unsafe.Slice((*byte)(ptr), 1<<cls)}
}

func putPinned(p *pinned) {
	cls := 0
	for (1 << cls) < cap(p.b) {
		cls++
	}
	pool, _ := pinnedPool.LoadOrStore(cls, &sync.Pool{})
	pool.(*sync.Pool).Put(p)
}

// EncodeBlocks runs encodeBlock / encodeBlockBetter / encodeBlockFast over a batch with
// one call, sharded over the device list.  src holds the blocks back to back, srcOff has
// len(blocks)+1 entries.  Block i's token stream is dst[dstOff[i] : dstOff[i]+outLen[i]];
// outLen 0 means "not compressible" exactly like a 0 return of encodeBlock.  crc[i] is
// the masked CRC-32C of block i (what the stream writer stores, writer.go:672).
func EncodeBlocks(dst, src []byte, srcOff []uint64, level int) (dstOff []uint64, outLen, crc []uint32, err error) {
	n := len(srcOff) - 1
	dstOff = make([]uint64, n)
	outLen = make([]uint32, n)
	crc = make([]uint32, n)
	devs := devices()
	err = call(func() C.int {
		return C.mzcu_stream_encode_blocks_multi(C.int(len(devs)), &devs[0], C.int(level), C.int(n), bptr(src),
			u64ptr(srcOff), bptr(dst), C.size_t(len(dst)), u64ptr(dstOff), u32ptr(outLen), u32ptr(crc))
	})
	return
}

// DecodeBlocks runs minLZDecode over a batch with one call; status[i] is the reference's
// return code (0 ok, 1 = decodeErrCodeCorrupt); crc[i] the masked CRC-32C of the decoded
// block (reader.go:341-351).
func DecodeBlocks(dst []byte, dstOff []uint64, src []byte, srcBeg []uint64, srcLen []uint32) (status []int32, crc []uint32, err error) {
	n := len(srcBeg)
	status = make([]int32, n)
	crc = make([]uint32, n)
	devs := devices()
	err = call(func() C.int {
		return C.mzcu_stream_decode_blocks_multi(C.int(len(devs)), &devs[0], C.int(n), bptr(src), u64ptr(srcBeg),
			u32ptr(srcLen), bptr(dst), u64ptr(dstOff), (*C.int32_t)(unsafe.Pointer(unsafe.SliceData(status))), u32ptr(crc))
	})
	return
}
