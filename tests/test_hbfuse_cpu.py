"""CPU: index algebra of the fused last-two-stages half-band pass (csrc/r8b_hbfuse.cuh).

The header is plain C++ over an accessor, so the very code the kernel runs is compiled with g++ here and
compared bit for bit with two plain CDSPHBUpsampler stages in a row (same summation order), including the
"negative stream indices are zeros" rule at the stream start."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def test_fused_last_two_stages_equal_two_plain_stages(tmp_path):
    exe = str(tmp_path / "hbfuse_check")
    src = os.path.join(HERE, "cpp", "hbfuse_check.cpp")
    subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-o", exe, src], check=True)
    r = subprocess.run([exe], stdout=subprocess.PIPE, text=True)
    sys.stdout.write(r.stdout)
    assert r.returncode == 0 and "hbfuse ok" in r.stdout
