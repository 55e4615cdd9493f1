"""In-tree builds: the sm_100a CUDA library (product) and the CPU oracle (test infrastructure).

`python -m wepp_b200.build` builds everything; __graft_entry__.build() calls build_all().
Outputs: wepp_b200/libwepp_b200.so, oracle/libwepp_oracle.so and — only when the reference
tree is mounted at /root/reference — oracle/_ref/libwepp_ref.so (see oracle/Makefile).
"""
from __future__ import annotations

import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "wepp_b200", "csrc")
LIB = os.path.join(ROOT, "wepp_b200", "libwepp_b200.so")
ORACLE_SRC = os.path.join(ROOT, "oracle", "wepp_oracle.cpp")
ORACLE_LIB = os.path.join(ROOT, "oracle", "libwepp_oracle.so")
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libwepp_ref.so")

NVCC_OBJ_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
                  "-Xcompiler", "-fPIC,-O3,-pthread"]


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _run(cmd: list[str]) -> None:
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("build failed: " + " ".join(cmd))


HOST_SRCS = ["host_prep.cpp", "host_arena.cpp", "host_io.cpp", "wepp_abi_io.cpp", "pipeline.cpp", "writers.cpp", "sam2pb.cpp"]
CLI = os.path.join(ROOT, "build", "wepp")   # where the Snakemake rules expect the binary (workflow/rules/filter.smk:3)
OBJ_DIR = os.path.join(ROOT, "wepp_b200", "build")


def build_cuda(force: bool = False) -> str:
    """wepp_abi.cu (all kernels) is compiled by nvcc for sm_100a, the host translation units by g++,
    each into its own object (rebuilt only when it or a header changed), then linked into one library."""
    headers = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(ROOT, "include", "wepp_b200.h"))
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    cu, cu_o = os.path.join(CSRC, "wepp_abi.cu"), os.path.join(OBJ_DIR, "wepp_abi.o")
    if force or not _newer(cu_o, [cu] + headers):
        _run([nvcc, *NVCC_OBJ_FLAGS, *os.environ.get("WEPP_NVCC_EXTRA", "").split(), "-c", cu, "-o", cu_o])
    objs.append(cu_o)
    for f in HOST_SRCS:
        src, obj = os.path.join(CSRC, f), os.path.join(OBJ_DIR, f[:-4] + ".o")
        if force or not _newer(obj, [src] + headers):
            _run(["g++", "-O3", "-std=c++17", "-fPIC", "-pthread", "-c", src, "-o", obj])
        objs.append(obj)
    if force or not _newer(LIB, objs):
        _run([nvcc, "-shared", "-Xcompiler", "-pthread", *objs, "-lz", "-o", LIB])
    main_src = os.path.join(CSRC, "wepp_main.cpp")
    if force or not _newer(CLI, [main_src, LIB]):
        os.makedirs(os.path.dirname(CLI), exist_ok=True)
        _run(["g++", "-O2", "-std=c++17", main_src, "-o", CLI, "-L" + os.path.dirname(LIB), "-lwepp_b200",
              "-Wl,-rpath,$ORIGIN/../wepp_b200", "-pthread"])
    return LIB


def build_oracle(force: bool = False) -> str:
    if not force and _newer(ORACLE_LIB, [ORACLE_SRC]):
        return ORACLE_LIB
    _run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-pthread", ORACLE_SRC, "-o", ORACLE_LIB])
    return ORACLE_LIB


def build_ref(force: bool = False) -> str | None:
    """Shim-compile the reference's own placement translation units (only where the reference
    tree is mounted; the GPU box uses the prebuilt oracle/_ref/libwepp_ref.so)."""
    mk = os.path.join(ROOT, "oracle", "Makefile")
    if not os.path.isdir("/root/reference/src/WEPP") or not os.path.exists(mk):
        return REF_LIB if os.path.exists(REF_LIB) else None
    _run(["make", "-s", "-j4", "-C", os.path.join(ROOT, "oracle"), "ref", "refcli"] + (["-B"] if force else []))
    return REF_LIB


def build_all(force: bool = False) -> None:
    build_cuda(force)
    build_oracle(force)
    build_ref(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
    print("built", LIB, ORACLE_LIB, REF_LIB if os.path.exists(REF_LIB) else "(no oracle/_ref)")
