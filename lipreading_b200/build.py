"""In-tree build of liblr_b200.so (the C-ABI library) for sm_100a.

    python -m lipreading_b200.build [--force]

nvcc cross-compiles without a GPU; the .so stays in-tree (git-ignored) so it travels to the GPU
box with the gpurun snapshot.
"""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB = os.path.join(HERE, "liblr_b200.so")
STAMP = os.path.join(HERE, ".liblr_b200.stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--shared", "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest():
    h = hashlib.sha256()
    files = sources() + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    files.append(os.path.join(INCLUDE, "lr_b200.h"))
    for f in files:
        h.update(f.encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ into liblr_b200.so; returns the library path."""
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as fh:
            if fh.read().strip() == dig:
                return LIB
    objs = []
    build_dir = os.path.join(HERE, "build")
    os.makedirs(build_dir, exist_ok=True)
    procs = []
    for src in sources():
        obj = os.path.join(build_dir, os.path.basename(src)[:-3] + ".o")
        cmd = [_nvcc(), "-c", src, "-o", obj, "-I", INCLUDE] + [f for f in NVCC_FLAGS if f != "--shared"]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, obj, p in procs:
        out, _ = p.communicate()
        log.append("== %s\n%s" % (os.path.basename(src), out))
        if p.returncode != 0:
            sys.stderr.write("\n".join(log))
            raise RuntimeError("nvcc failed on %s" % src)
        objs.append(obj)
    cmd = [_nvcc(), "--shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    subprocess.run(cmd, check=True)
    with open(os.path.join(build_dir, "ptxas.log"), "w") as fh:
        fh.write("\n".join(log))
    with open(STAMP, "w") as fh:
        fh.write(dig)
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
