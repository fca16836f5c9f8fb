"""Build veritas_b200/libveritas_b200.so in-tree with nvcc for sm_100a (no JIT cache, no CPU fallback).

    python -m veritas_b200.build [--force] [--verbose]

vrt_fields.cu and vrt_split.cu are compiled with -fmad=false (reference operation order);
vrt_fused.cu allows FMA contraction except in the speed/gamma chain, which uses explicit intrinsics.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
BUILD = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libveritas_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fno-strict-aliasing"]
UNITS = [
    ("vrt_abi.cu", []),
    ("vrt_fields.cu", ["-fmad=false"]),
    ("vrt_split.cu", ["-fmad=false"]),
    ("vrt_amr.cu", ["-fmad=false"]),
    ("vrt_fused.cu", []),
    ("vrt_init.cu", ["-fmad=false"]),
    ("vrt_comm.cu", []),
    ("vrt_checkpoint.cu", []),
]
HOST_UNITS = ["case_abi.cpp"]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers += [os.path.join(HERE, "..", "include", "veritas_b200.h"), os.path.join(HOST, "laser_plasma_case.hpp")]
    objs = []
    procs = []
    for src, extra in UNITS:
        path = os.path.join(CSRC, src)
        obj = os.path.join(BUILD, src + ".o")
        objs.append(obj)
        if force or _newer(obj, [path] + headers):
            cmd = ["nvcc"] + ARCH + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj]
            procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src in HOST_UNITS:
        path = os.path.join(HOST, src)
        obj = os.path.join(BUILD, src + ".o")
        objs.append(obj)
        if force or _newer(obj, [path] + headers):
            cmd = ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-c", path, "-o", obj]
            procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(" ".join(cmd) + "\n" + out)
        if p.returncode:
            raise RuntimeError("compile failed: " + " ".join(cmd))
    if force or procs or _newer(LIB, objs):
        cmd = ["nvcc"] + ARCH + ["-shared", "-o", LIB] + objs + ["-ldl"]
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
