"""Build recipes (in-tree, explicit nvcc/gcc; no JIT cache).

    python -m simdutf_b200.build            # product library only
    python -m simdutf_b200.build --all      # + oracle/ checkers + host test binary

The product is ONE shared library, simdutf_b200/libsimdutf_b200.so, holding the sm_100a kernels
(csrc/k_*.cu) and the C ABI of include/simdutf_b200.h (csrc/capi.cu).  It links the CUDA runtime
statically and nothing from oracle/.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "build")
LIB = os.path.join(PKG, "libsimdutf_b200.so")
SOURCES = ["k_utf8.cu", "k_utf16.cu", "k_base64.cu", "capi.cu"]
HEADERS = ["swar.h", "device_common.cuh", "launch.h", os.path.join("..", "..", "include", "simdutf_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "-Xptxas", "-v",
]


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps if os.path.exists(d))


def _run(cmd: list[str], log: str | None = None) -> None:
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if log:
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + p.stdout)
    if p.returncode != 0:
        sys.stderr.write(p.stdout)
        raise RuntimeError("command failed: " + " ".join(cmd))


def build_library(force: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    objs = []
    jobs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s.replace(".cu", ".o"))
        objs.append(obj)
        if force or not _newer(obj, [src] + hdrs):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        _run([NVCC] + NVCC_FLAGS + ["-c", src, "-o", obj], log=obj + ".log")

    if jobs:
        with ThreadPoolExecutor(max_workers=len(jobs)) as ex:
            list(ex.map(compile_one, jobs))
    if force or jobs or not _newer(LIB, objs):
        _run([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs + ["-cudart", "static", "-Xlinker", "--no-undefined"])
    return LIB


def build_oracle() -> None:
    _run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "all"])


def build_host_tests() -> str:
    out = os.path.join(ROOT, "tests", "host", "swar_host_test")
    src = os.path.join(ROOT, "tests", "host", "swar_host_test.cpp")
    deps = [src, os.path.join(CSRC, "swar.h"), os.path.join(ROOT, "oracle", "oracle.c")]
    if not _newer(out, deps):
        obj = os.path.join(OBJ, "oracle_c.o")
        os.makedirs(OBJ, exist_ok=True)
        _run(["gcc", "-O2", "-c", os.path.join(ROOT, "oracle", "oracle.c"), "-o", obj])
        _run(["g++", "-O2", "-std=c++17", "-o", out, src, obj])
    return out


if __name__ == "__main__":
    build_library(force="--force" in sys.argv)
    if "--all" in sys.argv:
        build_oracle()
        build_host_tests()
    print(LIB)
