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
DIAG_LIB = os.path.join(HERE, "liblr_b200_diag.so")      # product objects + csrc/diag/*.cu, hooks enabled (-DLR_DIAG)
DIAG_HOOKED = ("conv3d_sm100.cu",)                         # sources that carry #ifdef LR_DIAG measurement hooks
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


def diag_sources():
    d = os.path.join(CSRC, "diag")
    return sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cu")) if os.path.isdir(d) else []


def _digest():
    h = hashlib.sha256()
    files = sources() + diag_sources() + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    files.append(os.path.join(INCLUDE, "lr_b200.h"))
    files.append(os.path.join(INCLUDE, "lr_b200_diag.h"))
    for f in files:
        h.update(f.encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ into liblr_b200.so; returns the library path."""
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(DIAG_LIB) and os.path.exists(STAMP):
        with open(STAMP) as fh:
            if fh.read().strip() == dig:
                return LIB
    objs = []
    build_dir = os.path.join(HERE, "build")
    os.makedirs(build_dir, exist_ok=True)
    procs = []
    flags = [f for f in NVCC_FLAGS if f != "--shared"]
    jobs = [(src, os.path.join(build_dir, os.path.basename(src)[:-3] + ".o"), [], "product") for src in sources()]
    jobs += [(src, os.path.join(build_dir, "diag_" + os.path.basename(src)[:-3] + ".o"), ["-DLR_DIAG"], "diag")
             for src in sources() if os.path.basename(src) in DIAG_HOOKED] + \
            [(src, os.path.join(build_dir, "diag_" + os.path.basename(src)[:-3] + ".o"), ["-DLR_DIAG"], "diag")
             for src in diag_sources()]
    for src, obj, extra, kind in jobs:
        cmd = [_nvcc(), "-c", src, "-o", obj, "-I", INCLUDE] + flags + extra
        procs.append((src, obj, kind, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    diag_objs = []
    for src, obj, kind, p in procs:
        out, _ = p.communicate()
        log.append("== %s (%s)\n%s" % (os.path.basename(src), kind, out))
        if p.returncode != 0:
            sys.stderr.write("\n".join(log))
            raise RuntimeError("nvcc failed on %s" % src)
        (objs if kind == "product" else diag_objs).append(obj)
    arch = ["-gencode", "arch=compute_100a,code=sm_100a"]
    subprocess.run([_nvcc(), "--shared", "-o", LIB] + objs + arch, check=True)
    # the diagnostics library: the hooked sources recompiled with -DLR_DIAG replace their product objects
    hooked = {os.path.join(build_dir, n[:-3] + ".o") for n in DIAG_HOOKED}
    subprocess.run([_nvcc(), "--shared", "-o", DIAG_LIB] + [o for o in objs if o not in hooked] + diag_objs + arch, check=True)
    with open(os.path.join(build_dir, "ptxas.log"), "w") as fh:
        fh.write("\n".join(log))
    with open(STAMP, "w") as fh:
        fh.write(dig)
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
