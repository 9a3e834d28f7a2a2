// minlz.hpp -- C++ host side of the MinLZ block / stream API over the C ABI of
// libminlz_cuda.so (include/minlz_cuda.h).  Header only, C++17.
//
// The reference's host code is Go; this image has no Go toolchain, so the host
// side above the C ABI is written in C++ (the cgo files a maintainer would add
// are under go/, uncompiled).  Names, argument meaning and error behaviour
// mirror the Go package so callers and tests read like the reference's:
//
//   block API   Encode / AppendEncoded / TryEncode / MaxEncodedLen        encode.go:74-244
//               Decode / AppendDecoded / DecodedLen / IsMinLZ             decode.go:50-171
//               EncodeBlocks / DecodeBlocks (the seam, batched)           encode_amd64.go:111-118
//   flavour     SetEncoderFlavor (Go functions vs amd64 assembly bytes)   asm_none.go:15 / encode_amd64.go:15
//   index       Index::add / Find / reduce / appendTo / Load, IndexStream index.go:33-550
//   streams     Writer (Write / EncodeBuffer / Flush / Close / CloseIndex / Written; level,
//               block size, concurrency, index options)                   writer.go:40-1312
//               Reader (Read / ReadAll / WriteTo / Skip / Seek / ReadAt;  reader.go:248-1489
//               max block size, ignore CRC, ignore stream identifier)
//
// Go returns errors; here the same sentinel errors are thrown as minlz::Error
// (code() tells which).  All compression, decompression and CRC work happens in
// the CUDA kernels behind the C ABI: there is no CPU codec and no fallback here.
// Out of scope, as in the Python mirror: Snappy/S2 fallback, search tables,
// sidecars, padding on write, LevelSmallest.
#ifndef MINLZ_HPP
#define MINLZ_HPP

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <functional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "minlz_cuda.h"

namespace minlz {

using Bytes = std::vector<uint8_t>;

constexpr int LevelSuperFast = MZCU_LEVEL_SUPERFAST;      // encode.go:20-42
constexpr int LevelUncompressed = MZCU_LEVEL_UNCOMPRESSED;
constexpr int LevelFastest = MZCU_LEVEL_FASTEST;
constexpr int LevelBalanced = MZCU_LEVEL_BALANCED;
constexpr int MaxBlockSize = MZCU_MAX_BLOCK_SIZE;         // minlz.go:24
constexpr int FlavorGo = MZCU_FLAVOR_GO;
constexpr int FlavorAMD64 = MZCU_FLAVOR_AMD64;

// decode.go:29-40 and io errors the reference returns on these paths
enum class Err { Corrupt, TooLarge, Unsupported, InvalidLevel, CRC, Cuda, UnexpectedEOF, CantSeek, InvalidArg };

class Error : public std::runtime_error {
   public:
    Error(Err e, const std::string &msg) : std::runtime_error(msg), err_(e) {}
    Err code() const { return err_; }
    Bytes partial;  // Decode: what was written before the block turned out corrupt (decode.go:74-76)
   private:
    Err err_;
};

namespace detail {
[[noreturn]] inline void raise(int rc) {
    const char *m = mzcu_last_error();
    const std::string tail = (m && *m) ? std::string(": ") + m : std::string();
    switch (rc) {
        case MZCU_ERR_CORRUPT: throw Error(Err::Corrupt, "minlz: corrupt input");
        case MZCU_ERR_TOO_LARGE: throw Error(Err::TooLarge, "minlz: decoded block is too large");
        case MZCU_ERR_UNSUPPORTED: throw Error(Err::Unsupported, "minlz: unsupported input");
        case MZCU_ERR_INVALID_LEVEL: throw Error(Err::InvalidLevel, "minlz: invalid compression level");
        case MZCU_ERR_CUDA: throw Error(Err::Cuda, "minlz: cuda backend failure" + tail);
        default: throw Error(Err::InvalidArg, "minlz: error " + std::to_string(rc) + tail);
    }
}
inline void check(int rc) {
    if (rc < 0) raise(rc);
}
inline const uint8_t *ptr(const uint8_t *p, size_t n) { return n ? p : nullptr; }
inline void put_uvarint(Bytes &b, uint64_t x) {
    while (x >= 0x80) {
        b.push_back(uint8_t(x) | 0x80);
        x >>= 7;
    }
    b.push_back(uint8_t(x));
}
// encoding/binary.Uvarint: value and bytes read (<= 0 on failure)
inline int uvarint(const uint8_t *p, size_t n, uint64_t *out) {
    uint64_t x = 0;
    unsigned s = 0;
    for (size_t i = 0; i < n; i++) {
        if (i == 10) return -int(i + 1);
        const uint8_t b = p[i];
        if (b < 0x80) {
            if (i == 9 && b > 1) return -int(i + 1);
            *out = x | uint64_t(b) << s;
            return int(i + 1);
        }
        x |= uint64_t(b & 0x7f) << s;
        s += 7;
    }
    *out = 0;
    return 0;
}
// encoding/binary.PutVarint / Varint (zig-zag)
inline void put_varint(Bytes &b, int64_t x) {
    uint64_t ux = uint64_t(x) << 1;
    if (x < 0) ux = ~ux;
    put_uvarint(b, ux);
}
inline int varint(const uint8_t *p, size_t n, int64_t *out) {
    uint64_t ux = 0;
    const int k = uvarint(p, n, &ux);
    int64_t x = int64_t(ux >> 1);
    if (ux & 1) x = ~x;
    *out = x;
    return k;
}
inline int64_t go_div2(int64_t x) { return x / 2; }  // C++ truncates toward zero like Go
}  // namespace detail

inline int DeviceCount() { return mzcu_device_count(); }
// Which of the reference's two builds the encoders mirror byte for byte (process-wide).
inline void SetEncoderFlavor(int flavor) { detail::check(mzcu_set_encoder_flavor(flavor)); }
inline int EncoderFlavor() { return mzcu_get_encoder_flavor(); }

// ---------------------------------------------------------------- block API ----

inline int64_t MaxEncodedLen(int64_t srcLen) { return mzcu_max_encoded_len(srcLen); }  // encode.go:234

inline bool valid_level(int level) {
    return level == LevelSuperFast || level == LevelUncompressed || level == LevelFastest || level == LevelBalanced;
}

// encode.go:74-139
inline Bytes Encode(const uint8_t *src, size_t n, int level) {
    const int64_t cap = MaxEncodedLen(int64_t(n));
    if (cap < 0) throw Error(Err::TooLarge, "minlz: decoded block is too large");
    if (!valid_level(level)) {
        if (n < 16) {  // encode.go:83-85 runs before the level switch
            Bytes out;
            out.push_back(0);
            if (n) {
                out.push_back(0);
                out.insert(out.end(), src, src + n);
            }
            return out;
        }
        throw Error(Err::InvalidLevel, "minlz: invalid compression level");
    }
    Bytes out(size_t(std::max<int64_t>(cap, 1)));
    const int64_t r = mzcu_encode(out.data(), out.size(), detail::ptr(src, n), n, level);
    if (r < 0) detail::raise(int(r));
    out.resize(size_t(r));
    return out;
}
inline Bytes Encode(const Bytes &src, int level) { return Encode(src.data(), src.size(), level); }

// encode.go:144-163
inline void AppendEncoded(Bytes &dst, const Bytes &src, int level) {
    const Bytes e = Encode(src, level);
    dst.insert(dst.end(), e.begin(), e.end());
}

// encode.go:168-207: false when Go returns nil
inline bool TryEncode(Bytes &dst, const Bytes &src, int level) {
    const int64_t cap = MaxEncodedLen(int64_t(src.size()));
    if (cap < 0 || src.size() < 16 || !(level == LevelSuperFast || level == LevelFastest || level == LevelBalanced))
        return false;
    dst.resize(size_t(cap));
    const int64_t r = mzcu_try_encode(dst.data(), dst.size(), src.data(), src.size(), level);
    if (r < 0) detail::raise(int(r));
    dst.resize(size_t(r));
    return r > 0;
}

// decode.go:107-118
inline int64_t DecodedLen(const uint8_t *block, size_t n) {
    const int64_t r = mzcu_decoded_len(detail::ptr(block, n), n);
    if (r < 0) detail::raise(int(r));
    return r;
}
inline int64_t DecodedLen(const Bytes &b) { return DecodedLen(b.data(), b.size()); }
inline std::pair<bool, int64_t> IsMinLZ(const Bytes &b) {
    int ok = 0;
    int64_t size = 0;
    detail::check(mzcu_is_minlz(detail::ptr(b.data(), b.size()), b.size(), &ok, &size));
    return {ok != 0, size};
}

// decode.go:50-78 (MinLZ blocks; a first byte != 0 is Snappy/S2 territory: Err::Unsupported)
inline Bytes Decode(const uint8_t *block, size_t n) {
    // nothing is allocated before the block is known to be MinLZ of a legal size
    int ok = 0;
    int64_t dlen = 0;
    detail::check(mzcu_is_minlz(detail::ptr(block, n), n, &ok, &dlen));
    if (!ok) throw Error(Err::Unsupported, "minlz: unsupported input");
    if (dlen > MaxBlockSize) throw Error(Err::TooLarge, "minlz: decoded block is too large");
    Bytes out(size_t(std::max<int64_t>(dlen, 1)));
    const int64_t r = mzcu_decode(out.data(), size_t(dlen), detail::ptr(block, n), n);
    if (r == MZCU_ERR_CORRUPT) {
        Error e(Err::Corrupt, "minlz: corrupt input");
        out.resize(size_t(dlen));
        e.partial = std::move(out);
        throw e;
    }
    if (r < 0) detail::raise(int(r));
    out.resize(size_t(r));
    return out;
}
inline Bytes Decode(const Bytes &b) { return Decode(b.data(), b.size()); }
inline void AppendDecoded(Bytes &dst, const Bytes &block) {  // decode.go:85-103
    const Bytes d = Decode(block);
    dst.insert(dst.end(), d.begin(), d.end());
}

// The seam over a batch: encodeBlock / encodeBlockBetter / encodeBlockFast of every block with one
// launch.  src holds the blocks back to back, src_off has nblk+1 entries.  Returns the packed token
// streams; dst_off[i]..dst_off[i+1] is block i, an empty range means "not compressible".
inline Bytes EncodeBlocks(int level, const uint8_t *src, const std::vector<uint64_t> &src_off,
                          std::vector<uint64_t> &dst_off, std::vector<uint32_t> *crc = nullptr, int device = -1) {
    const int nblk = int(src_off.size()) - 1;
    dst_off.assign(size_t(nblk) + 1, 0);
    if (nblk <= 0) return {};
    Bytes dst(size_t(src_off.back() - src_off.front()) + 64);
    if (crc) crc->assign(size_t(nblk), 0);
    detail::check(mzcu_stream_encode_blocks(device, level, nblk, src, src_off.data(), dst.data(), dst.size(),
                                            dst_off.data(), crc ? crc->data() : nullptr));
    dst.resize(size_t(dst_off.back()));
    return dst;
}

// minLZDecode over a batch: status[i] is the reference's return code (0 ok, 1 corrupt).
inline std::vector<int32_t> DecodeBlocks(const uint8_t *src, const std::vector<uint64_t> &src_off, uint8_t *dst,
                                         const std::vector<uint64_t> &dst_off, std::vector<uint32_t> *crc = nullptr,
                                         int device = -1) {
    const int nblk = int(src_off.size()) - 1;
    std::vector<int32_t> status(size_t(std::max(nblk, 0)), 0);
    if (nblk <= 0) return status;
    if (crc) crc->assign(size_t(nblk), 0);
    detail::check(mzcu_stream_decode_blocks(device, nblk, src, src_off.data(), dst, dst_off.data(), status.data(),
                                            crc ? crc->data() : nullptr));
    return status;
}

// -------------------------------------------------------------------- index ----

constexpr char IndexHeader[] = "s2idx\x00";   // index.go:27 (6 bytes incl. the NUL)
constexpr char IndexTrailer[] = "\x00xdi2s";  // index.go:28
constexpr int maxIndexEntries = 1 << 16;
constexpr int64_t minIndexDist = 1 << 20;
constexpr uint8_t chunkTypeIndex = 0x40, legacyIndexChunk = 0x99;

struct OffsetPair {
    int64_t CompressedOffset, UncompressedOffset;
    bool operator==(const OffsetPair &o) const {
        return CompressedOffset == o.CompressedOffset && UncompressedOffset == o.UncompressedOffset;
    }
};

// index.go:33-410
class Index {
   public:
    int64_t TotalUncompressed = -1, TotalCompressed = -1;
    std::vector<OffsetPair> Offsets;
    int64_t estBlockUncomp = 0;

    void reset(int64_t maxBlock) {  // :55-68
        while (maxBlock < minIndexDist) maxBlock *= 2;
        estBlockUncomp = maxBlock;
        TotalCompressed = TotalUncompressed = -1;
        Offsets.clear();
    }
    void add(int64_t compressedOffset, int64_t uncompressedOffset) {  // :80-105
        if (!Offsets.empty()) {
            const OffsetPair &latest = Offsets.back();
            if (uncompressedOffset - latest.UncompressedOffset < estBlockUncomp) return;
            if (latest.UncompressedOffset > uncompressedOffset || latest.CompressedOffset > compressedOffset)
                throw Error(Err::InvalidArg, "minlz: internal error: earlier offset received");
        }
        Offsets.push_back({compressedOffset, uncompressedOffset});
        if (int(Offsets.size()) > maxIndexEntries) reduceLight();
    }
    // :114-144: entry at or before the (uncompressed) offset; negative = from the end
    OffsetPair Find(int64_t offset) const {
        if (TotalUncompressed < 0) throw Error(Err::Corrupt, "minlz: corrupt input");
        if (offset < 0) {
            offset += TotalUncompressed;
            if (offset < 0) throw Error(Err::UnexpectedEOF, "unexpected EOF");
        }
        if (offset > TotalUncompressed) throw Error(Err::UnexpectedEOF, "unexpected EOF");
        OffsetPair r{0, 0};
        if (Offsets.size() > 200) {
            auto it = std::upper_bound(Offsets.begin(), Offsets.end(), offset,
                                       [](int64_t v, const OffsetPair &p) { return v < p.UncompressedOffset; });
            size_t n = size_t(it - Offsets.begin());
            if (n == 0) n = 1;
            return Offsets[n - 1];
        }
        for (const OffsetPair &p : Offsets) {
            if (p.UncompressedOffset > offset) break;
            r = p;
        }
        return r;
    }
    void reduce() {  // :147-169
        if (int(Offsets.size()) < maxIndexEntries) return;
        int64_t removeN = (int64_t(Offsets.size()) + 1) / maxIndexEntries;
        while (estBlockUncomp * (removeN + 1) < minIndexDist && int64_t(Offsets.size()) / (removeN + 1) > 1000) removeN++;
        size_t j = 0;
        for (size_t idx = 0; idx < Offsets.size(); idx += size_t(removeN) + 1) Offsets[j++] = Offsets[idx];
        Offsets.resize(j);
        estBlockUncomp += estBlockUncomp * removeN;
    }
    void reduceLight() {  // :172-185 (incl. the loop's own idx++ after the inner scan)
        estBlockUncomp *= 2;
        size_t j = 0;
        for (size_t idx = 0; idx < Offsets.size(); idx++) {
            const OffsetPair base = Offsets[idx];
            Offsets[j++] = base;
            while (idx < Offsets.size() && Offsets[idx].UncompressedOffset - base.UncompressedOffset < estBlockUncomp) idx++;
        }
        Offsets.resize(j);
    }
    // :187-270
    Bytes appendTo(Bytes b, int64_t uncompTotal, int64_t compTotal) {
        reduce();
        const size_t init = b.size();
        b.insert(b.end(), {chunkTypeIndex, 0, 0, 0});
        b.insert(b.end(), IndexHeader, IndexHeader + 6);
        detail::put_varint(b, uncompTotal);
        detail::put_varint(b, compTotal);
        detail::put_varint(b, estBlockUncomp);
        detail::put_varint(b, int64_t(Offsets.size()));
        uint8_t hasUncompressed = 0;
        for (size_t i = 0; i < Offsets.size(); i++) {
            if (i == 0) {
                if (Offsets[i].UncompressedOffset != 0) {
                    hasUncompressed = 1;
                    break;
                }
                continue;
            }
            if (Offsets[i].UncompressedOffset != Offsets[i - 1].UncompressedOffset + estBlockUncomp) {
                hasUncompressed = 1;
                break;
            }
        }
        b.push_back(hasUncompressed);
        if (hasUncompressed)
            for (size_t i = 0; i < Offsets.size(); i++) {
                int64_t u = Offsets[i].UncompressedOffset;
                if (i > 0) u -= Offsets[i - 1].UncompressedOffset + estBlockUncomp;
                detail::put_varint(b, u);
            }
        int64_t cPredict = estBlockUncomp / 2;
        for (size_t i = 0; i < Offsets.size(); i++) {
            int64_t c = Offsets[i].CompressedOffset;
            if (i > 0) {
                c -= Offsets[i - 1].CompressedOffset + cPredict;
                cPredict += detail::go_div2(c);
            }
            detail::put_varint(b, c);
        }
        const uint32_t total = uint32_t(b.size() - init + 4 + 6);
        for (int k = 0; k < 4; k++) b.push_back(uint8_t(total >> (8 * k)));
        b.insert(b.end(), IndexTrailer, IndexTrailer + 6);
        const size_t chunkLen = b.size() - init - 4;
        b[init + 1] = uint8_t(chunkLen);
        b[init + 2] = uint8_t(chunkLen >> 8);
        b[init + 3] = uint8_t(chunkLen >> 16);
        return b;
    }
    // :273-410: returns the number of bytes consumed
    size_t Load(const uint8_t *b, size_t n) {
        if (n <= 4 + 6 + 6) throw Error(Err::UnexpectedEOF, "unexpected EOF");
        if (b[0] != chunkTypeIndex && b[0] != legacyIndexChunk) throw Error(Err::Corrupt, "minlz: corrupt input");
        const size_t chunkLen = size_t(b[1]) | size_t(b[2]) << 8 | size_t(b[3]) << 16;
        size_t p = 4;
        if (n - p < chunkLen) throw Error(Err::UnexpectedEOF, "unexpected EOF");
        if (std::memcmp(b + p, IndexHeader, 6) != 0) throw Error(Err::Unsupported, "minlz: unsupported input");
        p += 6;
        auto rd = [&](bool nonneg) {
            int64_t v;
            const int k = detail::varint(b + p, n - p, &v);
            if (k <= 0 || (nonneg && v < 0)) throw Error(Err::Corrupt, "minlz: corrupt input");
            p += size_t(k);
            return v;
        };
        TotalUncompressed = rd(true);
        TotalCompressed = rd(false);
        estBlockUncomp = rd(true);
        const int64_t entries = rd(true);
        if (entries > maxIndexEntries) throw Error(Err::Corrupt, "minlz: corrupt input");
        if (n - p < 1) throw Error(Err::UnexpectedEOF, "unexpected EOF");
        const uint8_t hasUncompressed = b[p++];
        if ((hasUncompressed & 1) != hasUncompressed) throw Error(Err::Corrupt, "minlz: corrupt input");
        Offsets.assign(size_t(entries), {0, 0});
        for (size_t i = 0; i < Offsets.size(); i++) {
            int64_t u = hasUncompressed ? rd(false) : 0;
            if (i > 0) {
                const int64_t prev = Offsets[i - 1].UncompressedOffset;
                u += prev + estBlockUncomp;
                if (u <= prev) throw Error(Err::Corrupt, "minlz: corrupt input");
            }
            if (u < 0) throw Error(Err::Corrupt, "minlz: corrupt input");
            Offsets[i].UncompressedOffset = u;
        }
        int64_t cPredict = estBlockUncomp / 2;
        for (size_t i = 0; i < Offsets.size(); i++) {
            int64_t c = rd(false);
            if (i > 0) {
                const int64_t cNew = cPredict + detail::go_div2(c);
                const int64_t prev = Offsets[i - 1].CompressedOffset;
                c += prev + cPredict;
                if (c <= prev) throw Error(Err::Corrupt, "minlz: corrupt input");
                cPredict = cNew;
            }
            if (c < 0) throw Error(Err::Corrupt, "minlz: corrupt input");
            Offsets[i].CompressedOffset = c;
        }
        if (n - p < 4 + 6) throw Error(Err::UnexpectedEOF, "unexpected EOF");
        p += 4;
        if (std::memcmp(b + p, IndexTrailer, 6) != 0) throw Error(Err::Corrupt, "minlz: corrupt input");
        return p + 6;
    }
    size_t Load(const Bytes &b) { return Load(b.data(), b.size()); }
};

// ------------------------------------------------------------------ streams ----

constexpr uint8_t chunkTypeLegacy = 0x00, chunkTypeUncompressedData = 0x01, chunkTypeMinLZCompressedData = 0x02,
                  chunkTypeMinLZCompressedDataCompCRC = 0x03, chunkTypeEOF = 0x20, maxNonSkippableChunk = 0x3f,
                  ChunkTypeStreamIdentifier = 0xff;
constexpr int defaultBlockSize = 2 << 20, minBlockSize = 4 << 10;

inline Bytes makeHeader(int blockSize) {  // writer.go:1553-1556
    Bytes h = {0xff, 0x06, 0x00, 0x00, 'M', 'i', 'n', 'L', 'z'};
    int bits = 0;
    while ((1 << bits) < blockSize) bits++;
    h.push_back(uint8_t(bits - 10));
    return h;
}

struct WriterOptions {
    int Level = LevelBalanced;          // writer.go:40
    int BlockSize = defaultBlockSize;   // WriterBlockSize
    int Concurrency = 0;                // WriterConcurrency: blocks per GPU call (0 = 256 MiB worth)
    bool CreateIndex = true;            // WriterCreateIndex
    bool AddIndex = false;              // WriterAddIndex
    int Device = -1;
};

// writer.go Writer: frames blocks of <= block size into a MinLZ stream; one batched GPU call
// (encode + CRC-32C of every block) per `Concurrency` blocks, chunk headers on the host.
class Writer {
   public:
    using Sink = std::function<void(const uint8_t *, size_t)>;
    Writer(Sink sink, WriterOptions o = WriterOptions()) : sink_(std::move(sink)), o_(o) {
        if (o_.BlockSize > MaxBlockSize || o_.BlockSize < minBlockSize)
            throw Error(Err::InvalidArg, "minlz: block size must be >= 4KB and <= 8MB");
        if (!valid_level(o_.Level)) throw Error(Err::InvalidLevel, "minlz: invalid compression level");
        if (o_.AddIndex && !o_.CreateIndex)
            throw Error(Err::InvalidArg, "WriterAddIndex: WriterCreateIndex has been called with false parameter");
        if (o_.Concurrency <= 0) o_.Concurrency = std::max(1, (256 << 20) / o_.BlockSize);
        index_.reset(o_.BlockSize);
    }
    // writer.go:276 Write: buffers; full batches of blocks are encoded as they fill
    size_t Write(const uint8_t *p, size_t n) {
        if (closed_) throw Error(Err::InvalidArg, "minlz: writer closed");
        ibuf_.insert(ibuf_.end(), p, p + n);
        const size_t batch = size_t(o_.Concurrency) * size_t(o_.BlockSize);
        if (ibuf_.size() >= batch) {
            const size_t take = (ibuf_.size() / batch) * batch;
            encodeBlocks(ibuf_.data(), take);
            ibuf_.erase(ibuf_.begin(), ibuf_.begin() + long(take));
        }
        return n;
    }
    // writer.go:441 EncodeBuffer: encodes buf directly (pending data first)
    void EncodeBuffer(const uint8_t *p, size_t n) {
        Flush();
        encodeBlocks(p, n);
    }
    void Flush() {  // writer.go:1006
        if (!ibuf_.empty()) {
            Bytes data;
            data.swap(ibuf_);
            encodeBlocks(data.data(), data.size());
        }
    }
    void Close() { closeIndex(o_.AddIndex); }          // writer.go:1033
    Bytes CloseIndex() { return closeIndex(true); }    // writer.go:1047
    std::pair<int64_t, int64_t> Written() const { return {uncompWritten_, written_}; }  // writer.go:1041

   private:
    void out(const uint8_t *p, size_t n) {
        if (sink_ && n) sink_(p, n);
        written_ += int64_t(n);
    }
    void encodeBlocks(const uint8_t *data, size_t n) {
        if (n == 0) return;
        if (!wroteHeader_) {
            wroteHeader_ = true;
            if (o_.CreateIndex) index_.add(written_, 0);  // writer.go:241: the header item is indexed too
            const Bytes h = makeHeader(o_.BlockSize);
            out(h.data(), h.size());
        }
        const size_t bs = size_t(o_.BlockSize), per = size_t(o_.Concurrency) * bs;
        for (size_t base = 0; base < n; base += per) {
            const uint8_t *part = data + base;
            const size_t psz = std::min(per, n - base);
            const size_t nblk = (psz + bs - 1) / bs;
            std::vector<uint64_t> soff(nblk + 1), doff;
            for (size_t i = 0; i <= nblk; i++) soff[i] = std::min(i * bs, psz);
            std::vector<uint32_t> crc(nblk, 0);
            Bytes comp;
            if (o_.Level == LevelUncompressed) {
                detail::check(mzcu_crc32c_blocks(o_.Device, int(nblk), part, soff.data(), crc.data()));
                doff.assign(nblk + 1, 0);
            } else {
                comp = EncodeBlocks(o_.Level, part, soff, doff, &crc, o_.Device);
            }
            Bytes chunk;
            for (size_t i = 0; i < nblk; i++) {
                const size_t a = size_t(soff[i]), b = size_t(soff[i + 1]);
                const size_t c0 = size_t(doff[i]), c1 = size_t(doff[i + 1]);
                if (o_.CreateIndex) index_.add(written_ + int64_t(chunk.size()), uncompWritten_ + int64_t(a));
                auto header = [&](uint8_t type, size_t len) {
                    chunk.insert(chunk.end(), {type, uint8_t(len), uint8_t(len >> 8), uint8_t(len >> 16)});
                    for (int k = 0; k < 4; k++) chunk.push_back(uint8_t(crc[i] >> (8 * k)));
                };
                if (c1 > c0) {  // writer.go:680-696: crc + uvarint(len) + tokens
                    Bytes lenhdr;
                    detail::put_uvarint(lenhdr, b - a);
                    header(chunkTypeMinLZCompressedData, 4 + lenhdr.size() + (c1 - c0));
                    chunk.insert(chunk.end(), lenhdr.begin(), lenhdr.end());
                    chunk.insert(chunk.end(), comp.begin() + long(c0), comp.begin() + long(c1));
                } else {        // n2 == 0: uncompressed chunk
                    header(chunkTypeUncompressedData, 4 + (b - a));
                    chunk.insert(chunk.end(), part + a, part + b);
                }
            }
            uncompWritten_ += int64_t(psz);
            out(chunk.data(), chunk.size());
        }
    }
    Bytes closeIndex(bool want) {  // writer.go:1051-1127
        if (closed_) return {};
        if (want && !o_.CreateIndex) throw Error(Err::InvalidArg, "index requested, but was asked to not generate one");
        Flush();
        Bytes eof = {chunkTypeEOF, 0, 0, 0};
        detail::put_uvarint(eof, uint64_t(uncompWritten_));
        eof[1] = uint8_t(eof.size() - 4);
        out(eof.data(), eof.size());
        Bytes idx;
        if (want) {
            idx = index_.appendTo({}, uncompWritten_, written_);
            if (o_.AddIndex) out(idx.data(), idx.size());
        }
        closed_ = true;
        return idx;
    }
    Sink sink_;
    WriterOptions o_;
    Bytes ibuf_;
    Index index_;
    bool wroteHeader_ = false, closed_ = false;
    int64_t uncompWritten_ = 0, written_ = 0;
};

struct ReaderOptions {
    int MaxBlockSize = minlz::MaxBlockSize;   // ReaderMaxBlockSize
    bool IgnoreCRC = false;                   // ReaderIgnoreCRC
    bool IgnoreStreamIdentifier = false;      // ReaderIgnoreStreamIdentifier
    int Concurrency = 128;                    // blocks per GPU call
    int Device = -1;
};

// reader.go Reader: parses chunks on the host, decodes + checksums blocks in GPU batches.
// The source is a read callback (returns bytes read, 0 at the end) and, for Seek / ReadAt,
// an optional absolute-seek callback.
class Reader {
   public:
    using Source = std::function<size_t(uint8_t *, size_t)>;
    using Seeker = std::function<void(int64_t)>;
    Reader(Source src, ReaderOptions o = ReaderOptions(), Seeker seek = nullptr)
        : src_(std::move(src)), seek_(std::move(seek)), o_(o), maxBlock_(o.MaxBlockSize), readHeader_(o.IgnoreStreamIdentifier) {
        if (o_.MaxBlockSize > minlz::MaxBlockSize || o_.MaxBlockSize <= 0)
            throw Error(Err::InvalidArg, "minlz: block size too large. Must be <= 8MB and > 0");
    }
    // reader.go:248 Read: up to n bytes; 0 at the end of the stream.  A stream error is thrown once
    // the data decoded before it has been delivered.
    size_t Read(uint8_t *p, size_t n) {
        while (out_.size() - outPos_ < n && !failed_ && !done_) fill(n - (out_.size() - outPos_));
        const size_t have = out_.size() - outPos_;
        if (have == 0 && failed_) throw err_;
        const size_t k = std::min(n, have);
        std::memcpy(p, out_.data() + outPos_, k);
        outPos_ += k;
        compact();
        return k;
    }
    Bytes ReadAll() {  // io.ReadAll(reader)
        while (!failed_ && !done_) fill(0);
        Bytes r(out_.begin() + long(outPos_), out_.end());
        out_.clear();
        outPos_ = 0;
        if (failed_) {
            err_.partial = r;
            throw err_;
        }
        return r;
    }
    // reader.go:548 WriteTo / :575 DecodeConcurrent
    int64_t WriteTo(const std::function<void(const uint8_t *, size_t)> &w) {
        int64_t total = 0;
        for (;;) {
            if (out_.size() == outPos_ && !done_ && !failed_) fill(0);
            if (out_.size() > outPos_) {
                w(out_.data() + outPos_, out_.size() - outPos_);
                total += int64_t(out_.size() - outPos_);
                out_.clear();
                outPos_ = 0;
                continue;
            }
            if (failed_) throw err_;
            if (done_) return total;
        }
    }
    // reader.go:1034 Skip: blocks lying entirely inside the skipped range are not decoded
    void Skip(int64_t n) {
        if (n < 0) throw Error(Err::InvalidArg, "attempted negative skip");
        if (failed_) throw err_;
        const int64_t take = std::min<int64_t>(n, int64_t(out_.size() - outPos_));
        outPos_ += size_t(take);
        n -= take;
        compact();
        if (n == 0) return;
        skipLeft_ += n;
        while (skipLeft_ && !failed_ && !done_) fill(1);
        if (failed_) throw err_;
        if (skipLeft_) {
            skipLeft_ = 0;
            fail(Error(Err::UnexpectedEOF, "unexpected EOF"));
            throw err_;
        }
    }
    // reader.go:1322 ReadSeeker: hand over the index (from CloseIndex / IndexStream) for Seek / ReadAt
    void LoadIndex(const Bytes &idx) {
        if (!seek_) throw Error(Err::CantSeek, "minlz: Can't seek because input stream isn't seekable");
        try {
            index_.Load(idx);
        } catch (const Error &e) {
            throw Error(Err::CantSeek, std::string("minlz: Can't seek because loading index returned: ") + e.what());
        }
        haveIndex_ = true;
    }
    const Index &GetIndex() const { return index_; }
    // reader.go:1373 Seek (absolute uncompressed offset; negative = from the end)
    int64_t Seek(int64_t offset) {
        if (!haveIndex_) throw Error(Err::CantSeek, "minlz: Can't seek because no index was loaded");
        if (failed_) throw err_;
        if (offset < 0) offset += index_.TotalUncompressed;
        if (offset < 0) throw Error(Err::InvalidArg, "seek before start of file");
        const int64_t lo = blockStart_ - int64_t(out_.size() - outPos_);
        if (skipLeft_ == 0 && lo <= offset && offset < blockStart_) {  // inside what is already decoded
            outPos_ += size_t(offset - lo);
            return offset;
        }
        const OffsetPair e = index_.Find(offset);
        seek_(e.CompressedOffset);
        out_.clear();
        outPos_ = 0;
        blockStart_ = e.UncompressedOffset;
        skipLeft_ = 0;
        done_ = false;
        wantEOF_ = false;
        readHeader_ = true;  // chunks are self-delimiting: parsing may start at any indexed chunk
        if (offset > e.UncompressedOffset) Skip(offset - e.UncompressedOffset);
        return offset;
    }
    // reader.go:1469 ReadAt: short only at the end of the stream
    Bytes ReadAt(size_t n, int64_t offset) {
        Seek(offset);
        Bytes r(n);
        size_t got = 0;
        while (got < n) {
            const size_t k = Read(r.data() + got, n - got);
            if (k == 0) break;
            got += k;
        }
        r.resize(got);
        return r;
    }

   private:
    struct Item {
        uint8_t type;
        uint32_t crc;
        Bytes body;
        size_t dlen;
    };
    void compact() {
        if (outPos_ > (1u << 20) && outPos_ * 2 > out_.size()) {
            out_.erase(out_.begin(), out_.begin() + long(outPos_));
            outPos_ = 0;
        }
    }
    void fail(const Error &e) {
        if (!failed_) {
            failed_ = true;
            err_ = e;
        }
    }
    // reads exactly n bytes; false on a clean end of input when allowed
    bool readFull(Bytes &b, size_t n, bool allowEOF) {
        b.resize(n);
        size_t got = 0;
        while (got < n) {
            const size_t k = src_(b.data() + got, n - got);
            if (k == 0) break;
            got += k;
        }
        if (got < n) {
            if (got == 0 && allowEOF) return false;
            throw Error(Err::Corrupt, "minlz: corrupt input (unexpected end of stream)");
        }
        return true;
    }
    void flushBatch(std::vector<Item> &batch) {
        if (batch.empty()) return;
        std::vector<Bytes> results(batch.size());
        std::vector<int> bad(batch.size(), 0);  // 1 corrupt, 2 crc
        std::vector<size_t> comp, raw;
        for (size_t i = 0; i < batch.size(); i++) (batch[i].type == chunkTypeUncompressedData ? raw : comp).push_back(i);
        if (!comp.empty()) {
            Bytes src;
            std::vector<uint64_t> soff(1, 0), doff(1, 0);
            for (size_t i : comp) {
                src.insert(src.end(), batch[i].body.begin(), batch[i].body.end());
                soff.push_back(src.size());
                doff.push_back(doff.back() + batch[i].dlen);
            }
            Bytes dst(size_t(std::max<uint64_t>(doff.back(), 1)));
            std::vector<uint32_t> crc, ccrc;
            const std::vector<int32_t> status = DecodeBlocks(src.data(), soff, dst.data(), doff, &crc, o_.Device);
            bool anyComp = false;
            for (size_t i : comp) anyComp |= batch[i].type == chunkTypeMinLZCompressedDataCompCRC;
            if (anyComp) {  // 0x03: the checksum covers the compressed bytes
                ccrc.assign(comp.size(), 0);
                detail::check(mzcu_crc32c_blocks(o_.Device, int(comp.size()), src.data(), soff.data(), ccrc.data()));
            }
            for (size_t k = 0; k < comp.size(); k++) {
                const size_t i = comp[k];
                if (status[k] != 0) {
                    bad[i] = 1;
                    continue;
                }
                const uint32_t got = batch[i].type == chunkTypeMinLZCompressedDataCompCRC ? ccrc[k] : crc[k];
                if (!o_.IgnoreCRC && got != batch[i].crc) {
                    bad[i] = 2;
                    continue;
                }
                results[i].assign(dst.begin() + long(doff[k]), dst.begin() + long(doff[k + 1]));
            }
        }
        if (!raw.empty()) {
            Bytes src;
            std::vector<uint64_t> soff(1, 0);
            for (size_t i : raw) {
                src.insert(src.end(), batch[i].body.begin(), batch[i].body.end());
                soff.push_back(src.size());
            }
            std::vector<uint32_t> crc(raw.size(), 0xa282ead8u);  // crc of the empty block
            if (!o_.IgnoreCRC && !src.empty())
                detail::check(mzcu_crc32c_blocks(o_.Device, int(raw.size()), src.data(), soff.data(), crc.data()));
            for (size_t k = 0; k < raw.size(); k++) {
                const size_t i = raw[k];
                if (!o_.IgnoreCRC && crc[k] != batch[i].crc) bad[i] = 2;
                else results[i] = std::move(batch[i].body);
            }
        }
        for (size_t i = 0; i < batch.size(); i++) {
            if (bad[i]) {
                fail(bad[i] == 1 ? Error(Err::Corrupt, "minlz: corrupt input") : Error(Err::CRC, "minlz: corrupt input, crc mismatch"));
                break;
            }
            blockStart_ += int64_t(results[i].size());
            size_t drop = 0;
            if (skipLeft_) {
                drop = size_t(std::min<int64_t>(skipLeft_, int64_t(results[i].size())));
                skipLeft_ -= int64_t(drop);
            }
            out_.insert(out_.end(), results[i].begin() + long(drop), results[i].end());
        }
        batch.clear();
    }
    // Parses chunks until a batch is complete (or the stream ends, or `budget` uncompressed bytes are
    // covered; 0 = no budget), then decodes it.
    void fill(size_t budget) {
        std::vector<Item> batch;
        size_t covered = 0;
        Bytes hdr, buf;
        try {
            while (int(batch.size()) < o_.Concurrency && (budget == 0 || covered < budget || batch.empty())) {
                if (!readFull(hdr, 4, !wantEOF_)) {
                    done_ = true;
                    break;
                }
                const uint8_t ctype = hdr[0];
                const size_t clen = size_t(hdr[1]) | size_t(hdr[2]) << 8 | size_t(hdr[3]) << 16;
                if (!readHeader_) {  // reader.go:273-284
                    if (ctype == ChunkTypeStreamIdentifier) readHeader_ = true;
                    else if (ctype <= maxNonSkippableChunk && ctype != chunkTypeEOF) throw Error(Err::Corrupt, "minlz: corrupt input");
                }
                if (ctype == chunkTypeMinLZCompressedData || ctype == chunkTypeMinLZCompressedDataCompCRC) {
                    if (clen < 4) throw Error(Err::Corrupt, "minlz: corrupt input");
                    readFull(buf, clen, false);
                    const uint32_t crc = uint32_t(buf[0]) | uint32_t(buf[1]) << 8 | uint32_t(buf[2]) << 16 | uint32_t(buf[3]) << 24;
                    uint64_t n;
                    const int hl = detail::uvarint(buf.data() + 4, buf.size() - 4, &n);
                    if (hl <= 0 || n > 0xffffffffull) throw Error(Err::Corrupt, "minlz: corrupt input");
                    if (n > uint64_t(maxBlock_)) throw Error(Err::TooLarge, "minlz: decoded block is too large");
                    const size_t body = buf.size() - 4 - size_t(hl);
                    if (n == 0 || n < body) throw Error(Err::Corrupt, "minlz: corrupt input");  // reader.go:327-333
                    if (batch.empty() && skipLeft_ >= int64_t(n)) {
                        skipLeft_ -= int64_t(n);
                        blockStart_ += int64_t(n);
                        continue;
                    }
                    batch.push_back({ctype, crc, Bytes(buf.begin() + 4 + hl, buf.end()), size_t(n)});
                    covered += size_t(n);
                } else if (ctype == chunkTypeUncompressedData) {
                    if (clen < 4) throw Error(Err::Corrupt, "minlz: corrupt input");
                    const size_t n = clen - 4;
                    if (n > size_t(maxBlock_)) throw Error(Err::TooLarge, "minlz: decoded block is too large");
                    readFull(buf, clen, false);
                    if (batch.empty() && skipLeft_ >= int64_t(n)) {
                        skipLeft_ -= int64_t(n);
                        blockStart_ += int64_t(n);
                        continue;
                    }
                    const uint32_t crc = uint32_t(buf[0]) | uint32_t(buf[1]) << 8 | uint32_t(buf[2]) << 16 | uint32_t(buf[3]) << 24;
                    batch.push_back({ctype, crc, Bytes(buf.begin() + 4, buf.end()), n});
                    covered += n;
                } else if (ctype == chunkTypeLegacy) {
                    throw Error(Err::Unsupported, "minlz: unsupported input");  // Snappy/S2 stays in host Go
                } else if (ctype == chunkTypeEOF) {
                    if (clen > 10) throw Error(Err::Corrupt, "minlz: corrupt input");
                    flushBatch(batch);  // the size check needs everything before it decoded
                    if (failed_) return;
                    if (clen != 0) {
                        readFull(buf, clen, false);
                        if (!o_.IgnoreStreamIdentifier) {
                            uint64_t want;
                            const int k = detail::uvarint(buf.data(), buf.size(), &want);
                            if (k != int(clen) || int64_t(want) != blockStart_) throw Error(Err::Corrupt, "minlz: corrupt input");
                        }
                    }
                    wantEOF_ = false;
                    readHeader_ = o_.IgnoreStreamIdentifier;
                } else if (ctype == ChunkTypeStreamIdentifier) {
                    if (clen != 6) throw Error(Err::Corrupt, "minlz: corrupt input");
                    readFull(buf, clen, false);
                    flushBatch(batch);
                    if (failed_) return;
                    blockStart_ = 0;
                    if (std::memcmp(buf.data(), "MinLz", 5) != 0) throw Error(Err::Unsupported, "minlz: unsupported input");
                    const uint8_t bs = buf[5];  // reader.go:994-1030 minLzHeader
                    if ((bs & 0xc0) || (bs & 15) > 13) throw Error(Err::Corrupt, "minlz: corrupt input");
                    const int blk = 1 << ((bs & 15) + 10);
                    if (blk > o_.MaxBlockSize) throw Error(Err::TooLarge, "minlz: decoded block is too large");
                    maxBlock_ = blk;
                    wantEOF_ = true;
                } else if (ctype <= maxNonSkippableChunk) {
                    throw Error(Err::Unsupported, "minlz: unsupported input");  // reserved unskippable chunk
                } else {
                    readFull(buf, clen, false);  // padding, index and other skippable chunks
                }
            }
        } catch (const Error &e) {
            flushBatch(batch);  // data before the failure is still delivered first
            fail(e);
            return;
        }
        flushBatch(batch);
    }

    Source src_;
    Seeker seek_;
    ReaderOptions o_;
    int maxBlock_;
    bool readHeader_, wantEOF_ = false, done_ = false, failed_ = false, haveIndex_ = false;
    Error err_{Err::Corrupt, ""};
    Bytes out_;
    size_t outPos_ = 0;
    int64_t blockStart_ = 0, skipLeft_ = 0;
    Index index_;
};

// index.go:455-550 IndexStream: index of an existing stream (structure checked, block data not)
inline Bytes IndexStream(const Reader::Source &read) {
    Index idx;
    idx.TotalCompressed = idx.TotalUncompressed = 0;
    bool readHeader = false;
    Bytes buf;
    auto readFull = [&](size_t n) -> size_t {
        buf.resize(n);
        size_t got = 0;
        while (got < n) {
            const size_t k = read(buf.data() + got, n - got);
            if (k == 0) break;
            got += k;
        }
        return got;
    };
    for (;;) {
        const size_t got = readFull(4);
        if (got == 0) return idx.appendTo({}, idx.TotalUncompressed, idx.TotalCompressed);
        if (got < 4) throw Error(Err::UnexpectedEOF, "unexpected EOF");
        const int64_t startChunk = idx.TotalCompressed;
        idx.TotalCompressed += 4;
        const uint8_t ctype = buf[0];
        if (!readHeader) {
            if (ctype != ChunkTypeStreamIdentifier && ctype != chunkTypeEOF) throw Error(Err::Corrupt, "minlz: corrupt input");
            readHeader = true;
        }
        const size_t clen = size_t(buf[1]) | size_t(buf[2]) << 8 | size_t(buf[3]) << 16;
        if (clen < 4) throw Error(Err::Corrupt, "minlz: corrupt input");
        idx.TotalCompressed += int64_t(clen);
        if (readFull(clen) != clen) throw Error(Err::UnexpectedEOF, "unexpected EOF");
        int64_t n2;
        if (ctype == chunkTypeMinLZCompressedData || ctype == chunkTypeMinLZCompressedDataCompCRC) {
            uint64_t v;
            if (detail::uvarint(buf.data() + 4, buf.size() - 4, &v) <= 0 || v > uint64_t(MaxBlockSize))
                throw Error(Err::Corrupt, "minlz: corrupt input");
            n2 = int64_t(v);
        } else if (ctype == chunkTypeUncompressedData) {
            n2 = int64_t(clen) - 4;
            if (n2 > MaxBlockSize) throw Error(Err::Corrupt, "minlz: corrupt input");
        } else if (ctype == ChunkTypeStreamIdentifier) {
            if (clen != 6) throw Error(Err::Corrupt, "minlz: corrupt input");
            continue;
        } else if (ctype == chunkTypeEOF) {
            continue;
        } else if (ctype <= maxNonSkippableChunk) {
            throw Error(Err::Unsupported, "minlz: unsupported input");
        } else {
            continue;  // user chunks and padding
        }
        if (idx.estBlockUncomp == 0) idx.estBlockUncomp = n2;
        idx.add(startChunk, idx.TotalUncompressed);
        idx.TotalUncompressed += n2;
    }
}

}  // namespace minlz
#endif  // MINLZ_HPP
