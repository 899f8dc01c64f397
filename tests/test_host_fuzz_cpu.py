"""Mutation fuzzing of the host-side parsers (WBRO, CSV, VSRI text, BRO stream layout) under
AddressSanitizer + UBSan: tools/host_fuzz.cpp, built from the shipped sources.  No GPU needed."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_parsers_survive_mutated_inputs(tmp_path):
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("no g++")
    exe = str(tmp_path / "host_fuzz")
    src = [os.path.join(ROOT, "tools", "host_fuzz.cpp")] + [os.path.join(ROOT, "atsc_b200", "csrc", f)
                                                            for f in ("ingest.cpp", "vsri.cpp", "stream.cpp")]
    r = subprocess.run([gxx, "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=all", *src,
                        "-o", exe, "-lpthread"], capture_output=True, text=True, timeout=300)
    if r.returncode != 0 and "sanitize" in r.stderr:
        pytest.skip("sanitizer runtime not available")
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([exe, "40000"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "no sanitizer report" in r.stdout, (r.stdout + r.stderr)[-3000:]
