"""Build the C-ABI CUDA library (libimw_b200.so) in-tree with nvcc for sm_100a."""
import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "lib" / "libimw_b200.so"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O3", "--expt-relaxed-constexpr",
]


def sources():
    return sorted(CSRC.glob("*.cu"))


def is_stale():
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*")) + [PKG.parent / "include" / "imw_b200.h"]
    return any(p.stat().st_mtime > t for p in deps)


def build(force=False, verbose=False):
    if not force and not is_stale():
        return LIB
    LIB.parent.mkdir(exist_ok=True)
    objs = []
    procs = []
    for src in sources():
        obj = LIB.parent / (src.stem + ".o")
        cmd = [NVCC, *FLAGS, "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if out.strip():
            print(f"--- {src.name}\n{out}")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    subprocess.check_call([NVCC, "-arch=sm_100a", "-shared", "-o", str(LIB), *map(str, objs)])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
