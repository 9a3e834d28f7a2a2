"""The encode kernels keep every block of the benchmark batch in flight at once: 28 one-warp CTAs per
SM x 148 SMs >= 4096 blocks.  That needs <= 72 registers per thread and 28 CTAs' shared memory per SM,
and the kernels are only as fast as measured while nothing spills (the four-warp build
spilled 172 bytes per thread and was 15 % slower).  Checked on the built library with cuobjdump; no GPU."""
import re
import shutil
import subprocess

import pytest

from minlz_b200 import build

SM_SHARED_BYTES = 227 * 1024   # usable shared memory per SM on sm_100
BLOCKS_PER_SM = 28


@pytest.fixture(scope="module")
def usage():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    try:
        out = subprocess.run([exe, "-res-usage", build.build()], capture_output=True, text=True, check=True).stdout
    except (OSError, subprocess.CalledProcessError) as e:
        pytest.skip("cuobjdump not usable here: %s" % e)
    res = {}
    for name, line in re.findall(r"Function (\S+):\n\s*(REG:.*)", out):
        res[name] = {k: int(v) for k, v in re.findall(r"([A-Z]+):(\d+)", line)}
    assert res, out[:400]
    return res


@pytest.mark.parametrize("kernel", ["encode_l1_kernel", "encode_l1_asm_kernel", "encode_l2_kernel", "encode_l2_asm_kernel"])
def test_encode_kernels_fit_28_blocks_per_sm_without_spills(usage, kernel):
    hits = {n: r for n, r in usage.items() if re.search(r"\d+%sI?" % kernel, n)}
    assert hits, sorted(usage)
    for name, r in hits.items():
        assert r["REG"] <= 72, (name, r)                      # 65536 / (32 threads x 28 CTAs) = 73
        assert r["STACK"] <= 8 and r["LOCAL"] == 0, (name, r)  # no spills in the walk
        assert BLOCKS_PER_SM * r["SHARED"] <= SM_SHARED_BYTES, (name, r)


def test_library_is_sm_100a_only():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    try:
        out = subprocess.run([exe, "-lelf", build.build()], capture_output=True, text=True, check=True).stdout
    except (OSError, subprocess.CalledProcessError) as e:
        pytest.skip("cuobjdump not usable here: %s" % e)
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs
