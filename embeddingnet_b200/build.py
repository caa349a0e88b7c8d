"""Builds ``libembeddingnet_b200.so`` in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m embeddingnet_b200.build [--force]

The shared library sits next to this file so that it travels with the source tree (it is git-ignored).
"""
from __future__ import annotations

import concurrent.futures
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libembeddingnet_b200.so")
OBJ_DIR = os.path.join(HERE, "_build")
SOURCES = ["core.cu", "rowwise.cu", "pairwise.cu", "batch_losses.cu", "pair_tc.cu", "knn.cu", "head.cu", "mine_bank.cu", "comm.cu", "host_pipe.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    # IEEE sqrt / div, no flush-to-zero, no fast-math: mined indices depend on exact float32 comparisons
    "-prec-sqrt=true", "-prec-div=true", "-ftz=false",
]


LAST_MODE = "not run"  # what the last build() call did (reported by __graft_entry__.build)


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _digest() -> str:
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for name in sorted(os.listdir(root)):
            if name.endswith((".cu", ".cuh", ".h")):
                with open(os.path.join(root, name), "rb") as f:
                    h.update(name.encode())
                    h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build_variant(out_path: str, extra_flags, only=("batch_losses.cu",)) -> str:
    """Developer aid (tools/ab_bh.py): a second library with extra -D switches on some sources, linked with the
    objects of the regular build.  Never loaded by the package itself."""
    build()
    nvcc = _nvcc()
    vdir = os.path.join(OBJ_DIR, "variant_" + hashlib.sha256(" ".join(extra_flags).encode()).hexdigest()[:8])
    os.makedirs(vdir, exist_ok=True)
    objs = []
    for src in SOURCES:
        if src in only:
            obj = os.path.join(vdir, src.replace(".cu", ".o"))
            subprocess.run([nvcc, *NVCC_FLAGS, *extra_flags, "-c", os.path.join(CSRC, src), "-o", obj], check=True)
        else:
            obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        objs.append(obj)
    subprocess.run([nvcc, "-shared", "-o", out_path, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-ldl"],
                   check=True)
    return out_path


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ_DIR, exist_ok=True)
    stamp = os.path.join(OBJ_DIR, "digest.txt")
    digest = _digest()
    global LAST_MODE
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
        LAST_MODE = "reused: the in-tree .so matches the digest of csrc/ + include/ + flags"
        return LIB
    LAST_MODE = "compiled %d sources with nvcc (sm_100a)" % len(SOURCES)
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose and r.stderr:
            print(r.stderr, file=sys.stderr)
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose=True)
    print(path)
