"""The 32-byte bin record (csrc/brg_bins.cuh: count / coverage / nearest-pair bins, wide and packed slot formats) checked on the
CPU: the header is plain C++, so g++ builds tests/harness/bins_record_harness.cpp around it -- records made the way
cb_fill_kernel makes them, both rank queries answered the way bins_pair_ranks answers them, every coordinate around the key's
range against a brute-force count, for every bin width of both formats."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bin_record_formats(tmp_path):
    exe = str(tmp_path / "bins_record_harness")
    src = os.path.join(ROOT, "tests", "harness", "bins_record_harness.cpp")
    inc = os.path.join(ROOT, "datafusion-bio-functions_b200", "csrc")
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Werror", "-I", inc, "-o", exe, src], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    sys.stdout.write(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("OK ")
