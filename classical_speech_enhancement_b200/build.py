"""Builds the native libraries in-tree.

* ``build_cuda()``  -> ``classical_speech_enhancement_b200/libcse_sm100a.so`` (fp32) and,
  with ``fp64=True``, ``libcse_sm100a_fp64.so`` (same symbols, ``-DCSE_FP64``): nvcc,
  ``-gencode arch=compute_100a,code=sm_100a -lineinfo``.  nvcc cross-compiles without a GPU.
* ``build_emu()``   -> ``tests/emu/libcse_emu.so``: the same sources compiled by g++ against a
  CPU thread-emulation of the CUDA subset they use.  TEST-ONLY; the package never loads it.
"""
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
SOURCES = [os.path.join(CSRC, "cse_lib.cu")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def _deps():
    out = []
    for d in (CSRC, os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "emu")):
        if os.path.isdir(d):
            out += [os.path.join(d, f) for f in os.listdir(d) if f.endswith((".cu", ".cuh", ".h", ".inl"))]
    return out


def _stale(target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in _deps())


def _run(cmd, log=None):
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if log:
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + p.stdout)
    if p.returncode != 0:
        sys.stderr.write(p.stdout)
        raise RuntimeError("build failed: " + " ".join(cmd))
    return p.stdout


def lib_path(fp64=False):
    return os.path.join(PKG, "libcse_sm100a_fp64.so" if fp64 else "libcse_sm100a.so")


def build_cuda(fp64=False, force=False, verbose=False):
    target = lib_path(fp64)
    if not force and not _stale(target):
        return target
    cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-DCSE_NO_FAST_MATH", "-Xptxas", "-v",
           "-Xcompiler", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include")]
    if fp64:
        cmd.append("-DCSE_FP64")
    tmp = f"{target}.part.{os.getpid()}"   # built aside and renamed: a reader (or a repository snapshot) never sees half a library
    cmd += SOURCES + ["-o", tmp]
    out = _run(cmd, log=os.path.join(PKG, "build_fp64.log" if fp64 else "build.log"))
    os.replace(tmp, target)
    if verbose:
        print(out)
    return target


def emu_path(fp64=False):
    return os.path.join(ROOT, "tests", "emu", "libcse_emu_fp64.so" if fp64 else "libcse_emu.so")


def build_emu(fp64=False, force=False):
    target = emu_path(fp64)
    if not force and not _stale(target):
        return target
    cmd = ["g++", "-O2", "-std=c++17", "-DCSE_EMU", "-x", "c++", "-fPIC", "-shared", "-pthread",
           "-ffp-contract=fast", "-Wno-unknown-pragmas",
           "-I", os.path.join(ROOT, "tests", "emu"), "-I", os.path.join(ROOT, "include")]
    if fp64:
        cmd.append("-DCSE_FP64")
    tmp = f"{target}.part.{os.getpid()}"       # per process: two builders (the ranks of a multi-process test) never share a file
    cmd += SOURCES + ["-o", tmp]
    _run(cmd)
    os.replace(tmp, target)
    return target


if __name__ == "__main__":
    what = sys.argv[1:] or ["cuda"]
    if "cuda" in what:
        print(build_cuda(force=True))
    if "fp64" in what:
        print(build_cuda(fp64=True, force=True))
    if "emu" in what:
        print(build_emu(force=True))
    if "emu64" in what:
        print(build_emu(fp64=True, force=True))
