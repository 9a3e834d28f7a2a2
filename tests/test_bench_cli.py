"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`) runs the CPU
implementation of the path on this host and prints ONE JSON line with the keys the driver reads,
for every workload; the workload defaults are the BASELINE.json configurations."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(args), capture_output=True, text=True,
                         timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, out.stdout
    return json.loads(lines[0])


@pytest.mark.parametrize("workload,extra", [
    ("blocks", ["--blocks", "16", "--block-size", "131072"]),
    ("stream", ["--blocks", "8", "--block-size", "262144", "--level", "2"]),
    ("sweep", ["--blocks", "4", "--block-size", "1048576"]),
])
def test_reference_arm_line(workload, extra):
    line = _run("--impl", "reference", "--workload", workload, "--steps", "2", "--warmup", "1", "--cpu-bytes", str(8 << 20), *extra)
    assert line["impl"] == "reference"
    assert line["unit"] == "GB/s" and line["higher_is_better"] is True and line["value"] > 0
    assert line["steps"] == 2 and line["n_gpus"] == 1 and line["dtype"] == "u8" and line["vs_baseline"] is None
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"]
    assert "touched once" in cb["sample"]
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    assert line["config"]["workload"].startswith({"blocks": "configs[1]+[2]", "stream": "configs[3]", "sweep": "configs[4]"}[workload])
    if workload == "sweep":
        assert set(line["sweep"]) == {"text", "binary", "random"}
        assert line["sweep"]["random"]["stored_blocks"] > 0      # incompressible blocks come back as "stored"
        assert line["sweep"]["text"]["stored_blocks"] == 0


def test_workload_defaults_are_the_baseline_configs():
    sys.path.insert(0, ROOT)
    import argparse
    import bench
    for wl, want in (("blocks", ("json", 1 << 20, 4096, 1)), ("stream", ("log", 2 << 20, 2048, 1)),
                     ("sweep", ("text+binary+random", 8 << 20, 512, 2))):
        a = argparse.Namespace(workload=wl, kind=None, block_size=None, blocks=None, level=None, flavor="auto")
        legs = bench.resolve_workload(a)
        assert (a.kind, a.block_size, a.blocks, a.level) == want
        assert a.flavor == "amd64"
        assert len(legs) == (3 if wl == "sweep" else 1)
