"""Builds libipb200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

Three translation units (the C ABI and one instantiation of the kernels per real kind) are compiled
in parallel and linked into one shared library next to this file.
"""
import hashlib
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
OUT = os.path.join(HERE, "libipb200.so")
UNITS = ["ipb_api.cu", "ipb_kind_f32.cu", "ipb_kind_f64.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",           # the reference is built without FMA contraction; rounding is part of the contract
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fno-fast-math", "-Xcompiler", "-ffp-contract=off",
    "-Xcudafe", "--diag_suppress=177",
]


def sources():
    root = os.path.dirname(HERE)
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [os.path.join(root, "shims", "ip_shim.c"), os.path.join(root, "include", "iplib.h"),
                                                                        os.path.join(root, "include", "ipb200.h")]


STAMP = os.path.join(OBJ, "libipb200.sha256")


def source_digest(extra=()):
    """Content hash of every source and of the flags: file times do not survive the copy to a GPU box."""
    h = hashlib.sha256(" ".join(NVCC_FLAGS + list(extra)).encode())
    for s in sources():
        h.update(os.path.basename(s).encode())
        h.update(open(s, "rb").read())
    return h.hexdigest()


def up_to_date(extra=()):
    if not (os.path.exists(OUT) and os.path.exists(STAMP) and all(os.path.exists(os.path.join(HERE, f"libip_{k}.so")) for k in "4d8")):
        return False
    return open(STAMP).read().strip() == source_digest(extra)


def build_library(force=False, verbose=False, extra=()):
    if not force and up_to_date(extra):
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ, exist_ok=True)

    def compile_one(u):
        o = os.path.join(OBJ, u.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + list(extra) + ["-c", "-o", o, os.path.join(CSRC, u)]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0 or (verbose and r.stderr.strip()):
            print(r.stdout + r.stderr, flush=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {u}")
        return o

    with ThreadPoolExecutor(len(UNITS)) as ex:
        objs = list(ex.map(compile_one, UNITS))
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", OUT] + objs
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    build_shims(verbose)
    with open(STAMP, "w") as f:
        f.write(source_digest(extra) + "\n")
    return OUT


def build_shims(verbose=False):
    """libip_4.so / libip_d.so / libip_8.so: the reference's unsuffixed C symbols forwarding to libipb200.so (shims/ip_shim.c)."""
    root = os.path.dirname(HERE)
    src = os.path.join(root, "shims", "ip_shim.c")
    outs = []
    for lsize, sfx in (("4", "4"), ("D", "d"), ("8", "8")):
        out = os.path.join(HERE, f"libip_{sfx}.so")
        cmd = ["gcc", "-O2", "-shared", "-fPIC", "-Wall", "-Werror", f"-DLSIZE={lsize}", "-I", os.path.join(root, "include"), src, "-o", out,
               "-L", HERE, "-lipb200", "-Wl,-rpath,$ORIGIN"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
        outs.append(out)
    return outs
