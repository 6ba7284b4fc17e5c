"""Builds the in-tree CUDA library libkzg.so for sm_100a (nvcc cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libkzg.so")
SOURCES = ["kzg_kernels.cu", "kzg_api.cu", "kzg_stream.cu", "kzg_enc.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-O3,-Wall", "--expt-relaxed-constexpr"]


def nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "kzg.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    extra = os.environ.get("KZG_NVCC_EXTRA", "").split()       # development only, e.g. -DKZG_FSE_WIDE_LIMIT=6
    cmd = [nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-shared", "-o", LIB] + srcs + ["-lcudart"]
    subprocess.check_call(cmd)
    return LIB


def find_jni_include() -> list[str]:
    """Include directories of a JDK's <jni.h>, or [] (this image has no JDK)."""
    cands = [os.environ.get("JAVA_HOME", "")] + sorted(
        os.path.join("/usr/lib/jvm", d) for d in (os.listdir("/usr/lib/jvm") if os.path.isdir("/usr/lib/jvm") else []))
    for c in cands:
        inc = os.path.join(c, "include")
        if c and os.path.exists(os.path.join(inc, "jni.h")):
            return [inc, os.path.join(inc, "linux")]
    return []


def build_jni(force: bool = False) -> str | None:
    """Layer C: libzstd-jni-kzg.so with the Java_com_ensody_kompressor_zstd_ZstdWrapper_* decompress symbols
    (csrc/kzg_jni.cpp), linked against libkzg.so.  Returns None where no JDK headers exist."""
    inc = find_jni_include()
    if not inc:
        return None
    out = os.path.join(HERE, "libzstd-jni-kzg.so")
    src = os.path.join(CSRC, "kzg_jni.cpp")
    if force or not os.path.exists(out) or os.path.getmtime(src) > os.path.getmtime(out):
        build()
        cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", out, src] + [f"-I{i}" for i in inc] + \
              [f"-L{HERE}", "-lkzg", "-Wl,-rpath,$ORIGIN"]
        subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
