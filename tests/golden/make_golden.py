"""Regenerates tests/golden/ from the reference checkout (run in the build
container only; /root/reference does not exist on the GPU box).

The fixtures are the reference's own test DATA (no code):
  Mark.Twain-Tom.Sawyer.txt(.mzb)   minlz_test.go:632-660 TestDecodeGoldenInput
  enc_regressions.zip               minlz_test.go:1538 TestDataRoundtrips
  dec-block-regressions.zip         fuzz_test.go:120 FuzzDecodeBlock seeds
  block-corpus-dec.zip              fuzz_test.go:120 (adversarial decode inputs)
  block-corpus-raw-sample.zip       deterministic sample (every 12th input,
                                    <= 256 KiB each) of fuzz/block-corpus-raw.zip
  block-corpus-enc-sample.zip       same sampling of fuzz/block-corpus-enc.zip
"""
import os
import shutil
import sys
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import corpus  # noqa: E402

REF = "/root/reference/testdata"


def sample(src_zip, dst_zip, every, max_size):
    items = sorted(corpus.load_zip(src_zip))
    with zipfile.ZipFile(dst_zip, "w", zipfile.ZIP_DEFLATED, compresslevel=9) as z:
        k = 0
        for i, (name, data) in enumerate(items):
            if i % every or len(data) > max_size:
                continue
            zi = zipfile.ZipInfo("s%04d" % k, date_time=(2026, 1, 1, 0, 0, 0))
            zi.compress_type = zipfile.ZIP_DEFLATED
            z.writestr(zi, data)
            k += 1
    return k


def main():
    for f in ("Mark.Twain-Tom.Sawyer.txt", "Mark.Twain-Tom.Sawyer.txt.mzb", "enc_regressions.zip",
              "dec-block-regressions.zip"):
        shutil.copyfile(os.path.join(REF, f), os.path.join(HERE, f))
    shutil.copyfile(os.path.join(REF, "fuzz/block-corpus-dec.zip"), os.path.join(HERE, "block-corpus-dec.zip"))
    n = sample(os.path.join(REF, "fuzz/block-corpus-raw.zip"), os.path.join(HERE, "block-corpus-raw-sample.zip"),
               12, 256 << 10)
    m = sample(os.path.join(REF, "fuzz/block-corpus-enc.zip"), os.path.join(HERE, "block-corpus-enc-sample.zip"),
               6, 256 << 10)
    print("raw sample:", n, "enc sample:", m)


if __name__ == "__main__":
    main()
