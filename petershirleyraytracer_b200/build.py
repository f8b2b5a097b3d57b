"""In-tree build of the CUDA library (sm_100a) and of the test-only oracle libraries.

    python -m petershirleyraytracer_b200.build            # librt_b200.so (+ oracle libs)

The product library is petershirleyraytracer_b200/librt_b200.so (git-ignored, travels to the GPU box).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(PKG_DIR)
LIB_PATH = os.path.join(PKG_DIR, "librt_b200.so")
HOST_MAIN = os.path.join(PKG_DIR, "rt_main")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",            # FP64 chain must not contract; FP32 cull uses explicit fmaf()
    "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources if os.path.exists(s))


def _digest(sources: list[str], flags: list[str]) -> str:
    """Content hash of the sources and the compiler flags: the library is rebuilt when THIS changes (not when a
    timestamp does -- a checkout or a copied tree can carry a stale .so that is newer than every source)."""
    import hashlib
    h = hashlib.sha256(" ".join(flags).encode())
    for s in sorted(sources):
        h.update(os.path.basename(s).encode())
        with open(s, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def build_cuda(force: bool = False, verbose: bool = False) -> str:
    src_dir = os.path.join(PKG_DIR, "csrc")
    sources = [os.path.join(src_dir, f) for f in sorted(os.listdir(src_dir))] + [os.path.join(REPO, "include", "rt.h")]
    stamp, digest = LIB_PATH + ".sha256", _digest(sources, NVCC_FLAGS)
    if not force and not verbose and os.path.exists(LIB_PATH) and os.path.exists(stamp) and open(stamp).read().strip() == digest:
        return LIB_PATH
    cmd = [_nvcc(), *NVCC_FLAGS, "-shared", os.path.join(src_dir, "rt_api.cu"), "-o", LIB_PATH]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.run(cmd, check=True, cwd=PKG_DIR)
    with open(stamp, "w") as f:
        f.write(digest + "\n")
    return LIB_PATH


def build_host_main(force: bool = False) -> str | None:
    """The reference-style main() (include/rt_host.hpp API) linked against the C ABI."""
    src = os.path.join(PKG_DIR, "host", "main.cc")
    if not os.path.exists(src):
        return None
    deps = [src] + [os.path.join(REPO, "include", f) for f in os.listdir(os.path.join(REPO, "include"))
                    if os.path.isfile(os.path.join(REPO, "include", f))]
    if not force and _newer(HOST_MAIN, deps + [LIB_PATH]):
        return HOST_MAIN
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-O2", "-std=c++17", "-I", os.path.join(REPO, "include"), "-I", os.path.join(REPO, "include", "compat"),
           src, "-o", HOST_MAIN,
           "-L", PKG_DIR, "-lrt_b200", "-Wl,-rpath,$ORIGIN"]
    subprocess.run(cmd, check=True, cwd=PKG_DIR)
    return HOST_MAIN


def build_oracle() -> None:
    """Test infrastructure: the C restatement always; oracle/_ref only where /root/reference exists."""
    odir = os.path.join(REPO, "oracle")
    subprocess.run(["make", "-s", "-C", odir, "oracle"], check=True)
    if os.path.isdir("/root/reference/programs"):
        subprocess.run(["make", "-s", "-C", odir, "ref", "ref_O0"], check=True)


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_cuda(force=force, verbose=verbose)
    build_host_main(force=force)
    build_oracle()


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB_PATH)
