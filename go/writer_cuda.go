//go:build cuda && cgo

// Batch points of the stream Writer on the GPU.
//
// The reference fans a buffer out to one goroutine per block (EncodeBuffer,
// writer.go:441-563; each goroutine: crc, uvarint, encodeBlock, chunk header,
// ordered hand-off through w.output).  Here the whole buffer -- every block of it --
// goes to the devices in ONE call that also returns the CRC-32C of every raw block;
// what is left on the host is the framing: 8 header bytes per chunk.
//
// Wiring: EncodeBuffer and ReadFrom call w.encodeBufferCUDA when w.customEnc == nil
// and the level is one of SuperFast / Fastest / Balanced (one `if` at writer.go:441
// and :316); everything else in writer.go is untouched.

package minlz

import (
	"encoding/binary"
	"sync"
)

// cudaBatchBlocks bounds one device call (pinned staging = 2 x batch bytes).
const cudaBatchBlocks = 4096

// cudaLevel reports whether the Writer's level (writer.go:135, switch at :573-582) is one
// of the levels on the accelerated path.
func (w *Writer) cudaLevel() (int, bool) {
	switch int(w.level) {
	case LevelSuperFast, LevelFastest, LevelBalanced:
		return int(w.level), true
	}
	return 0, false
}

// encodeBufferCUDA mirrors EncodeBuffer's concurrent path: same chunk bytes, same
// order, same bookkeeping (uncompWritten, stream header, pooled buffers).
func (w *Writer) encodeBufferCUDA(buf []byte, level int) error {
	if err := w.err(nil); err != nil {
		return err
	}
	if len(w.ibuf) > 0 { // flush queued data first (writer.go:451-456)
		if err := w.AsyncFlush(); err != nil {
			return err
		}
	}
	if !w.wroteStreamHeader {
		w.wroteStreamHeader = true
		hWriter := make(chan result)
		w.output <- hWriter
		hWriter <- result{startOffset: w.uncompWritten, b: makeHeader(w.blockSize)}
	}
	for len(buf) > 0 {
		// one device call per run of up to cudaBatchBlocks blocks
		part := buf
		if len(part) > cudaBatchBlocks*w.blockSize {
			part = part[:cudaBatchBlocks*w.blockSize]
		}
		buf = buf[len(part):]
		nblk := (len(part) + w.blockSize - 1) / w.blockSize
		off := make([]uint64, nblk+1)
		for i := range off {
			off[i] = uint64(min(i*w.blockSize, len(part)))
		}
		in, out := getPinned(len(part)), getPinned(len(part))
		copy(in.b, part)
		dstOff, outLen, crcs, err := EncodeBlocks(out.b[:len(part)], in.b[:len(part)], off, level)
		putPinned(in)
		if err != nil {
			putPinned(out)
			return w.err(err) // sticky, like every other writer error: no CPU fallback
		}
		for i := 0; i < nblk; i++ {
			uncompressed := part[off[i]:off[i+1]]
			obuf := w.buffers.Get().([]byte)[:len(uncompressed)+obufHeaderLen]
			output := make(chan result)
			w.output <- output // reserves this block's place in the stream (writer.go:498-499)
			res := result{startOffset: w.uncompWritten}
			w.uncompWritten += int64(len(uncompressed))

			chunkType := uint8(chunkTypeUncompressedData)
			chunkLen := 4 + len(uncompressed)
			if n2 := int(outLen[i]); n2 > 0 {
				n := binary.PutUvarint(obuf[obufHeaderLen:], uint64(len(uncompressed)))
				copy(obuf[obufHeaderLen+n:], out.b[dstOff[i]:dstOff[i]+uint64(n2)])
				chunkType = uint8(chunkTypeMinLZCompressedData)
				chunkLen = 4 + n + n2
				obuf = obuf[:obufHeaderLen+n+n2]
			} else {
				copy(obuf[obufHeaderLen:], uncompressed) // writer.go:515-516
			}
			checksum := crcs[i]
			obuf[0] = chunkType
			obuf[1] = uint8(chunkLen >> 0)
			obuf[2] = uint8(chunkLen >> 8)
			obuf[3] = uint8(chunkLen >> 16)
			obuf[4] = uint8(checksum >> 0)
			obuf[5] = uint8(checksum >> 8)
			obuf[6] = uint8(checksum >> 16)
			obuf[7] = uint8(checksum >> 24)
			res.b = obuf
			res.pooled = obuf
			go func() { output <- res }() // the ordered writer goroutine drains w.output (writer.go:219-272)
		}
		putPinned(out)
	}
	return nil
}

// WriterCUDA keeps the per-block hook usable (WriterCustomEncoder, writer.go:1293-1304:
// one call per block from one goroutine per block): callers are parked until
// `maxBlocks` of them are waiting -- use the writer's concurrency -- or until the
// Writer flushes, then one EncodeBlocks call serves them all.  There is no timer: a
// batch is cut by count or by flushCUDA(), never by the clock.
type cudaBatcher struct {
	mu      sync.Mutex
	level   int
	maxBlk  int
	pending []*cudaReq
}

type cudaReq struct {
	dst, src []byte
	n        int
	err      error
	done     chan struct{}
}

func WriterCUDA(level, maxBlocks int) (WriterOption, func()) {
	b := &cudaBatcher{level: level, maxBlk: maxBlocks}
	return WriterCustomEncoder(b.encode), b.flushPending
}

func (b *cudaBatcher) encode(dst, src []byte) int {
	r := &cudaReq{dst: dst, src: src, done: make(chan struct{})}
	b.mu.Lock()
	b.pending = append(b.pending, r)
	var batch []*cudaReq
	if len(b.pending) >= b.maxBlk {
		batch, b.pending = b.pending, nil
	}
	b.mu.Unlock()
	if batch != nil {
		b.run(batch)
	}
	<-r.done
	if r.err != nil {
		panic(r.err) // the hook has no error return and <0 would mean "use the CPU encoder": fail loudly instead
	}
	return r.n
}

// flushPending submits whatever is parked (call it from Flush / Close paths).
func (b *cudaBatcher) flushPending() {
	b.mu.Lock()
	batch := b.pending
	b.pending = nil
	b.mu.Unlock()
	if len(batch) > 0 {
		b.run(batch)
	}
}

func (b *cudaBatcher) run(batch []*cudaReq) {
	total := 0
	off := make([]uint64, len(batch)+1)
	for i, r := range batch {
		total += len(r.src)
		off[i+1] = uint64(total)
	}
	in, out := getPinned(total), getPinned(total)
	for i, r := range batch {
		copy(in.b[off[i]:], r.src)
	}
	dstOff, outLen, _, err := EncodeBlocks(out.b[:total], in.b[:total], off, b.level)
	for i, r := range batch {
		if err != nil {
			r.err = err
		} else {
			r.n = copy(r.dst, out.b[dstOff[i]:dstOff[i]+uint64(outLen[i])])
		}
		close(r.done)
	}
	putPinned(in)
	putPinned(out)
}
