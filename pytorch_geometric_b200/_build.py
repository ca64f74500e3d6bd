"""In-tree build of libb200mp.so (the C-ABI library) with nvcc for sm_100a.

No torch involvement: the library only depends on the CUDA runtime (statically linked), so the
same .so serves Python (ctypes), C, C++ or any FFI.  The .so lands in pytorch_geometric_b200/lib/ next to
a stamp file holding the SHA-256 of every source, header and compiler flag it was built from, and travels
to the GPU box with the repo snapshot (both are git-ignored, not gpurun-ignored).  Staleness is decided by
that fingerprint, NOT by file times: a snapshot copy does not keep mtimes in any useful order, and a
spurious rebuild costs minutes of nvcc on the GPU box.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.environ.get("B200MP_OBJ_DIR", os.path.join("/tmp", f"b200mp_build_{os.getuid()}"))  # objects stay out of tree
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libb200mp.so")
STAMP = os.path.join(LIBDIR, "libb200mp.sha256")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC",
    "-Wno-deprecated-declarations", "-DB200MP_BUILD",
]


def _nvcc() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found: the b200mp CUDA library cannot be built")
    return cand


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(INCLUDE, "b200mp.h"))
    return sorted(hs)


def _headers_mtime() -> float:
    return max(os.path.getmtime(h) for h in _headers())


def fingerprint() -> str:
    """SHA-256 over the compiler flags and the contents of every source and header."""
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for path in sources() + _headers():
        h.update(os.path.basename(path).encode())
        with open(path, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    try:
        with open(STAMP) as fh:
            return fh.read().strip() != fingerprint()
    except OSError:
        return True


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    nvcc = _nvcc()
    hm = _headers_mtime()

    def compile_one(src):
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        if (not force and os.path.exists(obj) and os.path.getmtime(obj) > os.path.getmtime(src)
                and os.path.getmtime(obj) > hm):
            return obj
        cmd = [nvcc, *NVCC_FLAGS, "-I", INCLUDE, "-c", src, "-o", obj]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 4)) as ex:
        objs = list(ex.map(compile_one, sources()))
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs,
           "-Xlinker", "--exclude-libs,ALL"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(STAMP, "w") as fh:
        fh.write(fingerprint() + "\n")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
