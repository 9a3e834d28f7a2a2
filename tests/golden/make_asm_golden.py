"""Generates tests/golden/asm_digests.json: SHA-256 digests of what the reference's OWN
amd64 assembly (asm_amd64.s, run through oracle/_ref) emits for a fixed set of inputs at
LevelSuperFast / LevelFastest / LevelBalanced, plus its decode status for mutated streams.

Run in the build container only (needs /root/reference to build oracle/_ref):
    python tests/golden/make_asm_golden.py
The digests travel with the repository, so the restated amd64 flavour of the oracle stays
pinned to real reference output on machines where oracle/_ref cannot be built
(tests/test_oracle_golden.py::test_asm_flavour_matches_recorded_reference_output).
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

import corpus  # noqa: E402
import patterns  # noqa: E402
import synth  # noqa: E402


def inputs():
    """(name, bytes): deterministic, reproducible without the reference checkout."""
    out = []
    for kind in ("json", "log", "text", "binary", "random"):
        big = synth.make_blocks(kind, 1, 3 << 20).numpy()[0]
        for n in (17, 33, 100, 1000, 1025, 4097, 16385, 65536, 65537, 300000, 524289, 1 << 20, (2 << 20) + 1, 3 << 20):
            out.append(("%s-%d" % (kind, n), big[:n].tobytes()))
    out.append(("twain", open(corpus.golden_path("Mark.Twain-Tom.Sawyer.txt"), "rb").read()))
    out.append(("text-8MiB", synth.make_blocks("text", 1, 8 << 20).numpy()[0].tobytes()))
    out.append(("large-offset-8MiB", patterns.large_offset(8 << 20, (2 << 20) + 70000)))
    for name in ("enc_regressions.zip", "block-corpus-raw-sample.zip"):
        for tag, data in corpus.load_zip(corpus.golden_path(name)):
            if len(data) >= 17:
                out.append((name + ":" + tag, data))
    return out


def main():
    from oracle import refasm
    assert refasm.build() is not None, "oracle/_ref is needed (build container with /root/reference)"
    rec = {}
    for name, data in inputs():
        rec[name] = {str(level): hashlib.sha256(refasm.encode_block(data, level)).hexdigest() for level in (-1, 1, 2)}
        rec[name]["n"] = len(data)
    path = os.path.join(HERE, "asm_digests.json")
    json.dump({"what": "sha256 of the token streams asm_amd64.s emits (encodeFastBlockAsm* / encodeBlockAsm* / "
                       "encodeBetterBlockAsm*, dispatch of encode_amd64.go), via oracle/_ref",
               "digests": rec}, open(path, "w"), indent=0, sort_keys=True)
    print("wrote", path, len(rec), "inputs")


if __name__ == "__main__":
    main()
