//go:build cuda && cgo

// Batch point of the stream Reader on the GPU.
//
// Reader.DecodeConcurrent (reader.go:575-992) reads chunks serially and hands every
// compressed block to a goroutine (reader.go:830-859: minLZDecode, CRC, ordered write).
// Here the reader loop collects the compressed chunks of a BATCH into one pinned buffer,
// one device call decodes them all and returns the CRC-32C of every decoded block, and
// the blocks are written out in stream order.  Chunk parsing, limits and error values are
// the reference's; Snappy / S2 legacy chunks (type 0x00) still go to klauspost/compress/s2
// on the CPU -- that decoder is not part of this path (decode.go:59-68).
//
// Wiring: DecodeConcurrent / WriteTo call r.decodeConcurrentCUDA(w, batchBlocks) first.

package minlz

import (
	"errors"
	"io"
)

type cudaBlock struct {
	beg      uint64 // token stream = comp[beg : beg+len]
	len      uint32
	n        int    // decoded length
	checksum uint32 // stored CRC
	compCRC  bool   // chunk 0x03: the CRC covers the compressed bytes (checked on the host)
	raw      []byte // uncompressed chunk 0x01: copied as is
}

func (r *Reader) decodeConcurrentCUDA(w io.Writer, batchBlocks int) (written int64, err error) {
	if r.i > 0 || r.j > 0 {
		return 0, errors.New("DecodeConcurrent called after Read")
	}
	if batchBlocks <= 0 {
		batchBlocks = 1024
	}
	var blocks []cudaBlock
	var comp *pinned
	compUsed := 0

	flush := func() error {
		if len(blocks) == 0 {
			return nil
		}
		total := 0
		dstOff := make([]uint64, 0, len(blocks)+1)
		srcBeg := make([]uint64, 0, len(blocks))
		srcLen := make([]uint32, 0, len(blocks))
		for _, b := range blocks {
			if b.raw == nil {
				dstOff = append(dstOff, uint64(total))
				srcBeg = append(srcBeg, b.beg)
				srcLen = append(srcLen, b.len)
				total += b.n
			}
		}
		dstOff = append(dstOff, uint64(total))
		out := getPinned(total)
		defer putPinned(out)
		var status []int32
		var crcs []uint32
		if len(srcBeg) > 0 {
			var e error
			status, crcs, e = DecodeBlocks(out.b[:total], dstOff, comp.b[:compUsed], srcBeg, srcLen)
			if e != nil {
				return e
			}
		}
		k := 0
		for _, b := range blocks {
			buf := b.raw
			if buf == nil {
				if status[k] != 0 {
					return ErrCorrupt // reader.go:836-843
				}
				if !r.ignoreCRC {
					got := crcs[k]
					if b.compCRC {
						got = crc(comp.b[b.beg : b.beg+uint64(b.len)])
					}
					if got != b.checksum {
						return ErrCRC // reader.go:848-855
					}
				}
				buf = out.b[dstOff[k]:dstOff[k+1]]
				k++
			}
			n, e := w.Write(buf)
			written += int64(n)
			if e != nil {
				return e
			}
			if n != len(buf) {
				return io.ErrShortWrite
			}
		}
		blocks = blocks[:0]
		compUsed = 0
		return nil
	}
	defer func() {
		if comp != nil {
			putPinned(comp)
		}
	}()

	for {
		if !r.readFull(r.tmp[:4], !r.wantEOF) {
			if r.err == io.EOF {
				r.err = nil
			}
			if r.err == nil {
				r.err = flush()
			}
			return written, r.err
		}
		chunkType := r.tmp[0]
		chunkLen := int(r.tmp[1]) | int(r.tmp[2])<<8 | int(r.tmp[3])<<16
		if !r.readHeader {
			if chunkType == ChunkTypeStreamIdentifier {
				r.readHeader = true
			} else if chunkType <= maxNonSkippableChunk && chunkType != chunkTypeEOF {
				r.err = ErrCorrupt
				return written, r.err
			}
		}
		switch chunkType {
		case chunkTypeMinLZCompressedData, chunkTypeMinLZCompressedDataCompCRC, chunkTypeUncompressedData:
			if chunkLen < checksumSize || chunkLen > r.maxBufSize {
				r.err = ErrCorrupt
				return written, r.err
			}
			if comp == nil {
				comp = getPinned(batchBlocks * r.maxBufSize / 2)
			}
			if compUsed+chunkLen > len(comp.b) || len(blocks) >= batchBlocks {
				if r.err = flush(); r.err != nil {
					return written, r.err
				}
			}
			buf := comp.b[compUsed : compUsed+chunkLen]
			if !r.readFull(buf, false) {
				return written, r.err
			}
			checksum := uint32(buf[0]) | uint32(buf[1])<<8 | uint32(buf[2])<<16 | uint32(buf[3])<<24
			buf = buf[checksumSize:]
			if chunkType == chunkTypeUncompressedData { // reader.go:866-905
				if len(buf) > r.maxBlock {
					r.err = ErrTooLarge
					return written, r.err
				}
				if !r.ignoreCRC && crc(buf) != checksum {
					r.err = ErrCRC
					return written, r.err
				}
				blocks = append(blocks, cudaBlock{raw: buf, n: len(buf)})
				compUsed += chunkLen
				continue
			}
			n, hdrSize, e := decodedLen(buf) // reader.go:810-826
			if e != nil {
				r.err = e
				return written, r.err
			}
			if n > r.maxBlock {
				r.err = ErrTooLarge
				return written, r.err
			}
			buf = buf[hdrSize:]
			if n == 0 || n < len(buf) {
				r.err = ErrCorrupt
				return written, r.err
			}
			blocks = append(blocks, cudaBlock{
				beg: uint64(compUsed + checksumSize + hdrSize), len: uint32(len(buf)), n: n, checksum: checksum,
				compCRC: chunkType == chunkTypeMinLZCompressedDataCompCRC,
			})
			compUsed += chunkLen
		default:
			// Stream identifier, EOF, index, padding, legacy and skippable chunks: drain the batch so
			// that stream order is kept, then let the reference's own chunk handling take over.
			if r.err = flush(); r.err != nil {
				return written, r.err
			}
			if r.err = r.handleOtherChunkCUDA(chunkType, chunkLen, w, &written); r.err != nil {
				return written, r.err
			}
		}
	}
}
