"""Build libkzg_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libkzg_b200.so")
SOURCES = ["abi.cu", "engine_one.cu", "host_helpers.cu", "poly.cu", "setup.cu", "setup_g2.cu", "g2_import.cu", "pairing.cu", "msm.cu", "msm_tail.cu", "msm_tree.cu", "msm_tree_r1.cu", "batch.cu", "batch_tree.cu", "probe.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v",
]


def _newer(src: str, dst: str) -> bool:
    if not os.path.exists(dst):
        return True
    deps = [src] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(HERE, "..", "include", "kzg_b200.h"))
    return any(os.path.getmtime(d) > os.path.getmtime(dst) for d in deps)


def _compile(src: str) -> str:
    obj = os.path.join(CSRC, src.replace(".cu", ".o"))
    path = os.path.join(CSRC, src)
    if not _newer(path, obj):
        return obj
    log = obj + ".log"
    cmd = ["nvcc", *NVCC_FLAGS, "-c", path, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj


def build(force: bool = False) -> str:
    if force:
        for s in SOURCES:
            o = os.path.join(CSRC, s.replace(".cu", ".o"))
            if os.path.exists(o):
                os.remove(o)
    with cf.ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(_compile, SOURCES))
    if force or not os.path.exists(OUT) or any(os.path.getmtime(o) > os.path.getmtime(OUT) for o in objs):
        cmd = ["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static",
               "-Xlinker", "-soname=libkzg_b200.so", "-o", OUT, *objs]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return OUT


CLI_SRC = os.path.join(HERE, "..", "cli", "kzg_cli.cpp")
CLI_OUT = os.path.join(HERE, "..", "cli", "kzg-cli")


def build_cli() -> str:
    """The reference's command-line tool (src/main.rs) over the C ABI: plain g++, links libkzg_b200.so."""
    deps = [CLI_SRC, OUT, os.path.join(HERE, "..", "include", "kzg_b200.hpp"), os.path.join(HERE, "..", "include", "kzg_b200.h")]
    if os.path.exists(CLI_OUT) and all(os.path.getmtime(d) <= os.path.getmtime(CLI_OUT) for d in deps):
        return CLI_OUT
    cmd = ["g++", "-std=c++17", "-O2", "-I", os.path.join(HERE, "..", "include"), CLI_SRC, "-o", CLI_OUT,
           "-L", HERE, "-l:libkzg_b200.so", "-Wl,-rpath,$ORIGIN/../kzg_poly_commit_exploration_b200"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"g++ failed for the CLI:\n{r.stdout}\n{r.stderr}")
    return CLI_OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
    print(build_cli())
