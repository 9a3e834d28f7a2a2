//go:build cuda && cgo

// Batch hook for the stream Writer.  WriterCustomEncoder (writer.go:1293-1304)
// is called once per block from one goroutine per block (writer.go:670); the
// GPU wants all blocks of a buffer at once, so this collector parks the
// callers until a batch is full (or a short timer fires), submits one
// EncodeBlocks call and hands every caller its slice.  Contract of the hook:
// return bytes used in dst, 0 = incompressible (writer emits chunk 0x01),
// <0 = fall through to the built-in encoder.

package minlz

import (
	"sync"
	"time"
)

type cudaBatcher struct {
	mu      sync.Mutex
	level   int
	pending []*cudaReq
	timer   *time.Timer
	maxBlk  int
}

type cudaReq struct {
	dst, src []byte
	n        int
	done     chan struct{}
}

// WriterCUDA returns a WriterOption that encodes blocks on the GPU in batches
// of up to maxBlocks (use the writer's concurrency).
func WriterCUDA(level, maxBlocks int) WriterOption {
	b := &cudaBatcher{level: level, maxBlk: maxBlocks}
	return WriterCustomEncoder(b.encode)
}

func (b *cudaBatcher) encode(dst, src []byte) int {
	r := &cudaReq{dst: dst, src: src, done: make(chan struct{})}
	b.mu.Lock()
	b.pending = append(b.pending, r)
	if len(b.pending) >= b.maxBlk {
		batch := b.pending
		b.pending = nil
		b.mu.Unlock()
		b.flush(batch)
	} else {
		if b.timer == nil {
			b.timer = time.AfterFunc(200*time.Microsecond, b.timeout)
		}
		b.mu.Unlock()
	}
	<-r.done
	return r.n
}

func (b *cudaBatcher) timeout() {
	b.mu.Lock()
	batch := b.pending
	b.pending = nil
	b.timer = nil
	b.mu.Unlock()
	if len(batch) > 0 {
		b.flush(batch)
	}
}

func (b *cudaBatcher) flush(batch []*cudaReq) {
	total := 0
	off := make([]uint64, len(batch)+1)
	for i, r := range batch {
		total += len(r.src)
		off[i+1] = uint64(total)
	}
	flat := make([]byte, total) // a production build pools pinned buffers (mzcu_host_alloc)
	for i, r := range batch {
		copy(flat[off[i]:], r.src)
	}
	out := make([]byte, total)
	doff, err := EncodeBlocks(out, flat, off, b.level)
	for i, r := range batch {
		if err != nil {
			r.n = -1 // fall back to the built-in encoder for this block
		} else {
			r.n = copy(r.dst, out[doff[i]:doff[i+1]])
		}
		close(r.done)
	}
}
