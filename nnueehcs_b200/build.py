"""In-tree build of the native library (``nnueehcs_b200/_native/libnnueehcs_b200.so``).

nvcc cross-compiles for sm_100a without a GPU; the built ``.so`` is git-ignored but travels to
the GPU box with the repo snapshot.  ``python -m nnueehcs_b200.build`` or
``__graft_entry__.build()`` runs it; ``nnueehcs_b200._lib`` only *loads* the result and fails
loudly when it is missing.
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
OUT_DIR = os.path.join(HERE, "_native")
LIB_PATH = os.path.join(OUT_DIR, "libnnueehcs_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build the nnueehcs_b200 native library")


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _fingerprint() -> str:
    h = hashlib.sha256()
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))
             if os.path.isfile(os.path.join(CSRC, f))]
    files.append(os.path.join(INCLUDE, "nnueehcs_b200.h"))
    for p in files:
        with open(p, "rb") as f:
            h.update(os.path.basename(p).encode())   # not the checkout root: the stamp travels
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _object_fingerprint(src, extra_flags) -> str:
    """One translation unit: its source, every header of csrc/ + the public header, the flags."""
    h = hashlib.sha256()
    deps = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))
            if not f.endswith(".cu") and os.path.isfile(os.path.join(CSRC, f))]
    deps += [os.path.join(INCLUDE, "nnueehcs_b200.h"), src]
    for p in deps:
        with open(p, "rb") as f:
            h.update(os.path.basename(p).encode())
            h.update(f.read())
    h.update(" ".join([*NVCC_FLAGS, *extra_flags]).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False, trace: bool = False) -> str:
    """Compile every ``csrc/*.cu`` for sm_100a and link the shared library.  Returns its path.

    ``trace=True`` builds the bring-up variant ``libnnueehcs_b200_trace.so`` (-DUQ_TC_TRACE: the
    fused kernel logs a per-role event timeline of CTA 0); select it with NNUEEHCS_B200_LIB."""
    os.makedirs(OUT_DIR, exist_ok=True)
    if trace:
        return _build_variant("_trace", ["-DUQ_TC_TRACE"], verbose)
    stamp = os.path.join(OUT_DIR, "build.stamp")
    fp = _fingerprint()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(stamp):
        with open(stamp) as f:
            if f.read().strip() == fp:
                return LIB_PATH
    _build_variant("", [], verbose)
    with open(stamp, "w") as f:
        f.write(fp)
    return LIB_PATH


def _build_variant(suffix, extra_flags, verbose) -> str:
    nvcc = _nvcc()
    lib_path = LIB_PATH[:-3] + suffix + ".so"
    objs = []

    def compile_one(src):
        obj = os.path.join(OUT_DIR, os.path.basename(src)[:-3] + suffix + ".o")
        fp, fp_path = _object_fingerprint(src, extra_flags), obj[:-2] + ".stamp"
        if os.path.exists(obj) and os.path.exists(fp_path):
            with open(fp_path) as f:
                if f.read().strip() == fp:
                    return obj   # unchanged translation unit
        cmd = [nvcc, *NVCC_FLAGS, *extra_flags, "-I", INCLUDE, "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        with open(obj[:-2] + ".ptxas.log", "w") as f:
            f.write(r.stderr)
        if verbose:
            sys.stderr.write(r.stderr)
        with open(fp_path, "w") as f:
            f.write(fp)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, _sources()))
    cmd = [nvcc, "-shared", "-o", lib_path, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
           "-Xcompiler", "-fPIC"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return lib_path


if __name__ == "__main__":
    # bring-up variants: `--variant NAME -DFLAG ...` builds libnnueehcs_b200_NAME.so
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        os.makedirs(OUT_DIR, exist_ok=True)
        print(_build_variant("_" + sys.argv[i + 1], sys.argv[i + 2:], False))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv,
                    trace="--trace" in sys.argv))
