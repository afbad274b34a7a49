"""The stream-K bookkeeping of the tensor-core kernel (csrc/sched.cuh: CTA unit ranges, segment iterator, owner /
contributors of a split tile, grid planning) is plain integer code shared by host and device; tests/sched_check.cpp
checks its invariants on the CPU over the BASELINE.json shapes and 60 000 random launch plans."""
import os
import subprocess


def test_stream_k_bookkeeping_invariants(tmp_path):
    here = os.path.dirname(os.path.abspath(__file__))
    exe = str(tmp_path / "sched_check")
    subprocess.run(["g++", "-O1", "-std=c++17", "-Wall", "-Werror", "-o", exe, os.path.join(here, "sched_check.cpp")], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.startswith("sched ok"), out.stdout + out.stderr
