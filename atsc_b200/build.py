"""In-tree build of libatsc_gpu.so (CUDA kernels + C ABI + stream layer) for sm_100a.

    python -m atsc_b200.build            # build if stale
    python -m atsc_b200.build --force

nvcc cross-compiles without a GPU; the .so stays in-tree (git-ignored) so it travels
to the GPU box with the repo snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libatsc_gpu.so")
CLI = os.path.join(HERE, "atsc")
CSV_CLI = os.path.join(HERE, "csv-compressor")
SOURCES = ["kernels.cu", "api.cu", "stream.cpp", "ingest.cpp", "vsri.cpp"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--expt-extended-lambda", "-Xcompiler", "-fPIC",

    # f32 FFT may contract to FMA; every f64 value computation that must be bit-exact uses
    # explicit __d*_rn intrinsics, which are never contracted.
]


def _deps():
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    files += [os.path.join(HERE, "host", f) for f in os.listdir(os.path.join(HERE, "host"))]
    files.append(os.path.join(os.path.dirname(HERE), "include", "atsc_gpu.h"))
    return files


def is_stale():
    if not all(os.path.exists(p) for p in (OUT, CLI, CSV_CLI)):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(f) > t for f in _deps())


def build(force=False, verbose=False):
    if not force and not is_stale():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    procs = []
    bdir = os.path.join(HERE, "build")
    os.makedirs(bdir, exist_ok=True)
    for s in SOURCES:
        o = os.path.join(bdir, s.rsplit(".", 1)[0] + ".o")
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-x", "cu", "-c", os.path.join(CSRC, s), "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    for s, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError(f"nvcc failed on {s}")
    cmd = [nvcc, "-shared", "-o", OUT] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    subprocess.check_call(cmd)
    # the `atsc` command line (reference option surface) on top of the library
    # and the `csv-compressor` tool (csv-compressor/src/main.rs)
    for src, exe in (("atsc_cli.cpp", CLI), ("csv_compressor_cli.cpp", CSV_CLI)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", os.path.join(HERE, "host", src), "-o", exe,
                               "-L" + HERE, "-latsc_gpu", "-Wl,-rpath,$ORIGIN"])
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
