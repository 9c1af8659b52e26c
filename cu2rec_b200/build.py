"""Builds every native artefact of the repository in-tree (no JIT cache, no site-packages).

  cu2rec_b200/lib/libcu2b.so   the product: sm_100a kernels + C ABI + host IO   (nvcc)
  bin/mf, bin/predict, bin/prep  drop-in CLIs over the C ABI                     (g++)

(The test oracle under oracle/ is built by __graft_entry__.build(), not from here: the package
never touches it.)
nvcc cross-compiles for sm_100a without a GPU, so this runs on the CPU-only builder too.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "cu2rec_b200", "csrc")
LIBDIR = os.path.join(ROOT, "cu2rec_b200", "lib")
BINDIR = os.path.join(ROOT, "bin")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _run(cmd, cwd=None):
    proc = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + proc.stdout + "\n")
        raise RuntimeError("build step failed: " + " ".join(cmd[:3]) + " ...")
    return proc.stdout


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def lib_sources():
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cpp")) and not f.startswith("cli_")]
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "cu2b.h")]
    return srcs, deps


def build_lib(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    out = os.path.join(LIBDIR, "libcu2b.so")
    srcs, deps = lib_sources()
    if force or _newer(out, deps):
        cmd = [NVCC, *ARCH, "-lineinfo", "-O3", "-std=c++17", "-I" + os.path.join(ROOT, "include"), "-I" + CSRC,
               "-Xcompiler", "-fPIC,-fopenmp,-O3", "-shared", "-o", out, *srcs]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        log = _run(cmd)
        if verbose:
            print(log)
    return out


def build_cli(force=False):
    os.makedirs(BINDIR, exist_ok=True)
    outs = []
    for name in ("mf", "predict", "prep"):
        src = os.path.join(CSRC, "cli_%s.cpp" % name)
        if not os.path.exists(src):
            continue
        out = os.path.join(BINDIR, name)
        deps = [src, os.path.join(ROOT, "include", "cu2b.h"), os.path.join(LIBDIR, "libcu2b.so")]
        shim = os.path.join(CSRC, "cu2rec_shim.h")
        if os.path.exists(shim):
            deps.append(shim)
        if force or _newer(out, deps):
            _run(["g++", "-O2", "-std=c++17", "-I" + os.path.join(ROOT, "include"), "-I" + CSRC, "-o", out, src,
                  "-L" + LIBDIR, "-lcu2b", "-Wl,-rpath,$ORIGIN/../cu2rec_b200/lib"])
        outs.append(out)
    return outs


def build_micro(force=False):
    """tools/micro/l2_rows: the L2 row-traffic micro-benchmark bench.py takes its roofline denominator from."""
    src = os.path.join(ROOT, "tools", "micro", "l2_rows.cu")
    out = os.path.join(ROOT, "tools", "micro", "l2_rows")
    if os.path.exists(src) and (force or _newer(out, [src])):
        _run([NVCC, *ARCH, "-lineinfo", "-O3", "-Xcompiler", "-fopenmp", "-o", out, src])
    return out


def build_all(force=False, verbose=False):
    build_lib(force, verbose)
    build_cli(force)
    build_micro(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built:", os.path.join(LIBDIR, "libcu2b.so"))
