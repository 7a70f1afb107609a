"""Builds libgenmap_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m genmap_b200._build [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libgenmap_b200.so")
SOURCES = ["capi.cu", "map_kernel.cu", "exact_kernel.cu", "block_kernel.cu", "locate_kernel.cu", "rle_kernel.cu", "jump_table.cu", "index_build_gpu.cu", "gmb_host.cpp", "seqan_export.cpp"]
HEADERS = ["gmb_layout.h", "gmb_core.h", "gmb_host.h", "sais.hpp", "map_kernel.cuh", "map_kernel_impl.cuh", "locate.cuh", "rle.cuh", "jump_table.cuh", "index_build_gpu.cuh",
           os.path.join("..", "..", "include", "genmap_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Wno-deprecated-declarations", "-shared"]


def nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


BIN_DIR = os.path.join(HERE, "bin")
CLI = os.path.join(BIN_DIR, "genmap")
CLI_SOURCES = [os.path.join("cli", "genmap_cli.cpp"), os.path.join("cli", "writers.hpp")]


def build_cli(force=False, verbose=False):
    """The `genmap` command line (host C++), linked against the library with an $ORIGIN rpath."""
    srcs = [os.path.join(CSRC, s) for s in CLI_SOURCES]
    if not force and os.path.exists(CLI) and all(os.path.getmtime(s) <= os.path.getmtime(CLI) for s in srcs + [LIB]):
        return CLI
    os.makedirs(BIN_DIR, exist_ok=True)
    cmd = ["/usr/bin/g++", "-O2", "-std=c++17", "-Wall", "-o", CLI + ".tmp", srcs[0], "-L" + LIB_DIR, "-lgenmap_b200",
           "-Wl,-rpath,$ORIGIN/../lib", "-pthread"]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    os.replace(CLI + ".tmp", CLI)
    return CLI


def build(force=False, verbose=False):
    """Compile every CUDA/C++ source of the product into genmap_b200/lib/libgenmap_b200.so."""
    if not force and not is_stale():
        build_cli(force, verbose)
        return LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    obj_dir = os.path.join(HERE, "..", "build", "obj")
    os.makedirs(obj_dir, exist_ok=True)
    env = dict(os.environ, CC="", CXX="")
    flags = [f for f in NVCC_FLAGS if f != "-shared"]

    def compile_one(src):
        obj = os.path.join(obj_dir, src.replace(".", "_") + ".o")
        cmd = [nvcc()] + flags + ["-c", "-o", obj, os.path.join(CSRC, src)]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True, env=env)
        return obj

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as pool:
        objs = list(pool.map(compile_one, SOURCES))
    cmd = [nvcc()] + NVCC_FLAGS + ["-o", LIB + ".tmp"] + objs
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.run(cmd, check=True, env=env)
    os.replace(LIB + ".tmp", LIB)
    build_cli(True, verbose)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
