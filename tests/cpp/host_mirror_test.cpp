// Driver for the C++ host mirror (include/minlz.hpp).  Two modes:
//   host_mirror_test index <out>            no GPU: Index appendTo / Load / Find / reduce (compared with index.py by the test)
//   host_mirror_test gpu <input> <outdir>   block API, Writer / Reader / Skip / Seek / ReadAt, IndexStream on the device
// Exit code 0 and a final "OK" line mean every internal check held; the files it
// writes are compared byte for byte with the Python mirror / oracle by tests/test_cpp_host.py.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>

#include "minlz.hpp"

using minlz::Bytes;

#define CHECK(c)                                                                 \
    do {                                                                         \
        if (!(c)) {                                                              \
            std::fprintf(stderr, "CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #c); \
            std::exit(1);                                                        \
        }                                                                        \
    } while (0)

static Bytes slurp(const std::string &p) {
    std::ifstream f(p, std::ios::binary);
    return Bytes((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}
static void dump(const std::string &p, const Bytes &b) {
    std::ofstream f(p, std::ios::binary);
    f.write(reinterpret_cast<const char *>(b.data()), long(b.size()));
}
template <class F>
static bool throws(minlz::Err want, F f) {
    try {
        f();
    } catch (const minlz::Error &e) {
        return e.code() == want;
    }
    return false;
}

static int index_mode(const std::string &out) {
    // varints (encoding/binary examples)
    Bytes v;
    minlz::detail::put_varint(v, -1);
    CHECK(v == Bytes({0x01}));
    v.clear();
    minlz::detail::put_varint(v, 64);
    CHECK(v == Bytes({0x80, 0x01}));
    // regular spacing, then irregular
    minlz::Index ix;
    ix.estBlockUncomp = 1 << 20;
    for (int i = 0; i < 50; i++) ix.Offsets.push_back({10 + i * 400000 + (i * i) % 977, int64_t(i) << 20});
    const Bytes a = ix.appendTo({}, 50ll << 20, 20000000);
    dump(out + "/index_regular.bin", a);
    minlz::Index back;
    CHECK(back.Load(a) == a.size());
    CHECK(back.Offsets == ix.Offsets && back.TotalUncompressed == (50ll << 20) && back.TotalCompressed == 20000000);
    for (size_t i = 0; i < ix.Offsets.size(); i++) ix.Offsets[i].UncompressedOffset += int64_t(i % 3) * 1000 + (i ? 0 : 5);
    const Bytes b = ix.appendTo({}, 60ll << 20, -1);
    dump(out + "/index_irregular.bin", b);
    CHECK(back.Load(b) == b.size() && back.Offsets == ix.Offsets && back.TotalCompressed == -1);
    // Find
    minlz::Index f;
    CHECK(throws(minlz::Err::Corrupt, [&] { f.Find(0); }));
    f.TotalUncompressed = 10 << 20;
    for (int i = 0; i < 10; i++) f.Offsets.push_back({i * 300000, int64_t(i) << 20});
    CHECK(f.Find((3 << 20) - 1) == (minlz::OffsetPair{600000, 2 << 20}));
    CHECK(f.Find(-1) == (minlz::OffsetPair{2700000, 9 << 20}));
    CHECK(throws(minlz::Err::UnexpectedEOF, [&] { f.Find((10 << 20) + 1); }));
    // add spacing, reduceLight quirk
    minlz::Index s;
    s.reset(64 << 10);
    CHECK(s.estBlockUncomp == 1 << 20);
    for (int i = 0; i < 100; i++) s.add(i * 30000, int64_t(i) * (64 << 10));
    CHECK(s.Offsets.size() == 7 && s.Offsets[6].UncompressedOffset == 6 << 20);
    minlz::Index l;
    l.estBlockUncomp = 1 << 20;
    for (int i = 0; i < 10; i++) l.Offsets.push_back({i * 10, int64_t(i) << 20});
    l.reduceLight();
    CHECK(l.Offsets.size() == 4 && l.Offsets[1].UncompressedOffset == 3 << 20 && l.Offsets[3].UncompressedOffset == 9 << 20);
    // Load rejects
    CHECK(throws(minlz::Err::UnexpectedEOF, [&] { minlz::Index().Load(a.data(), 12); }));
    Bytes bad = a;
    bad[0] = 0x41;
    CHECK(throws(minlz::Err::Corrupt, [&] { minlz::Index().Load(bad); }));
    bad = a;
    bad.back() = 'X';
    CHECK(throws(minlz::Err::Corrupt, [&] { minlz::Index().Load(bad); }));
    CHECK(minlz::makeHeader(2 << 20).back() == 11 && minlz::makeHeader(4 << 10).back() == 2);
    std::puts("OK");
    return 0;
}

static int gpu_mode(const std::string &in, const std::string &out) {
    const Bytes data = slurp(in);
    CHECK(minlz::DeviceCount() >= 1);
    CHECK(minlz::MaxEncodedLen(0) == 1 && minlz::MaxEncodedLen(int64_t(data.size())) == int64_t(data.size()) + 2);
    // block API at every level
    for (int level : {minlz::LevelSuperFast, minlz::LevelUncompressed, minlz::LevelFastest, minlz::LevelBalanced}) {
        const Bytes enc = minlz::Encode(data, level);
        dump(out + "/block_" + std::to_string(level) + ".mzb", enc);
        CHECK(int64_t(enc.size()) <= minlz::MaxEncodedLen(int64_t(data.size())));
        CHECK(minlz::DecodedLen(enc) == int64_t(data.size()) && minlz::IsMinLZ(enc).first);
        CHECK(minlz::Decode(enc) == data);
        Bytes t;
        if (minlz::TryEncode(t, data, level)) CHECK(t.size() < data.size() && minlz::Decode(t) == data);
    }
    // the other flavour: different bytes, same content
    minlz::SetEncoderFlavor(minlz::FlavorAMD64);
    const Bytes asm1 = minlz::Encode(data, minlz::LevelFastest);
    dump(out + "/block_1_amd64.mzb", asm1);
    CHECK(minlz::Decode(asm1) == data);
    minlz::SetEncoderFlavor(minlz::FlavorGo);
    // wrapper edge cases (encode.go:83-85,223-229; decode.go:74-76)
    CHECK(minlz::Encode(Bytes(), 1) == Bytes({0}));
    CHECK(minlz::Encode(Bytes({'a', 'b', 'c'}), 1) == Bytes({0, 0, 'a', 'b', 'c'}));
    CHECK(throws(minlz::Err::InvalidLevel, [&] { minlz::Encode(data, 7); }));
    {
        Bytes broken = minlz::Encode(data, 1);
        broken.resize(broken.size() - 3);
        bool got = false;
        try {
            minlz::Decode(broken);
        } catch (const minlz::Error &e) {
            got = e.code() == minlz::Err::Corrupt && e.partial.size() == data.size();
        }
        CHECK(got);
    }
    // stream: Writer with an appended index
    Bytes stream;
    minlz::WriterOptions wo;
    wo.Level = minlz::LevelFastest;
    wo.BlockSize = 64 << 10;
    wo.AddIndex = true;
    minlz::Writer w([&](const uint8_t *p, size_t n) { stream.insert(stream.end(), p, p + n); }, wo);
    w.Write(data.data(), data.size() / 3);
    w.Write(data.data() + data.size() / 3, data.size() - data.size() / 3);
    const Bytes idx = w.CloseIndex();
    CHECK(w.Written().first == int64_t(data.size()) && w.Written().second == int64_t(stream.size()));
    dump(out + "/stream.mz", stream);
    dump(out + "/index.bin", idx);
    // Reader: everything, then Skip over a non-seekable source, then Seek / ReadAt
    auto source = [](const Bytes &b, size_t *pos) {
        return [&b, pos](uint8_t *p, size_t n) {
            const size_t k = std::min(n, b.size() - *pos);
            std::memcpy(p, b.data() + *pos, k);
            *pos += k;
            return k;
        };
    };
    size_t pos = 0;
    CHECK(minlz::Reader(source(stream, &pos)).ReadAll() == data);
    minlz::Index ix;
    ix.Load(idx);
    for (int64_t want = 0; want < int64_t(data.size()); want += 555555) {  // index_test.go:30 ExampleIndex_Load
        const minlz::OffsetPair e = ix.Find(want);
        size_t p2 = size_t(e.CompressedOffset);
        minlz::ReaderOptions ro;
        ro.IgnoreStreamIdentifier = true;
        minlz::Reader r(source(stream, &p2), ro);
        r.Skip(want - e.UncompressedOffset);
        CHECK(r.ReadAll() == Bytes(data.begin() + long(want), data.end()));
    }
    pos = 0;
    minlz::Reader rs(source(stream, &pos), minlz::ReaderOptions(), [&](int64_t o) { pos = size_t(o); });
    rs.LoadIndex(idx);
    for (int64_t off : {int64_t(0), int64_t(1), int64_t(65535), int64_t(65536), int64_t(1 << 20) + 1, int64_t(data.size()) - 1}) {
        const Bytes got = rs.ReadAt(1000, off);
        const size_t n = std::min<size_t>(1000, data.size() - size_t(off));
        CHECK(got == Bytes(data.begin() + long(off), data.begin() + long(off) + long(n)));
    }
    CHECK(rs.Seek(-100) == int64_t(data.size()) - 100);
    // index of an existing stream
    pos = 0;
    const Bytes idx2 = minlz::IndexStream(source(stream, &pos));
    dump(out + "/index_stream.bin", idx2);
    minlz::Index i2;
    i2.Load(idx2);
    CHECK(i2.TotalUncompressed == int64_t(data.size()));
    // stream errors: flipped payload byte -> CRC or corrupt, after the data before it
    Bytes hurt = stream;
    hurt[hurt.size() / 2] ^= 0x55;
    pos = 0;
    bool failed = false;
    try {
        minlz::Reader(source(hurt, &pos)).ReadAll();
    } catch (const minlz::Error &e) {
        failed = (e.code() == minlz::Err::CRC || e.code() == minlz::Err::Corrupt) && !e.partial.empty();
    }
    CHECK(failed);
    std::puts("OK");
    return 0;
}

int main(int argc, char **argv) {
    if (argc >= 3 && std::string(argv[1]) == "index") return index_mode(argv[2]);
    if (argc >= 4 && std::string(argv[1]) == "gpu") return gpu_mode(argv[2], argv[3]);
    std::fprintf(stderr, "usage: %s index <outdir> | gpu <input> <outdir>\n", argv[0]);
    return 2;
}
