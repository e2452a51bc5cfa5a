"""Build the CUDA library in-tree: wgdi_b200/libicl_b200.so (sm_100a only).

    python -m wgdi_b200.build            # nvcc, cross-compiles without a GPU
    python -m wgdi_b200.build --emu DIR  # test-suite only: host emulation of the kernel bodies (g++)

The emulation library is written outside the package (tests/_emu/) and is never loaded by wgdi_b200.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "icl_b200.cu")
DEPS = [SRC, os.path.join(HERE, "csrc", "icl_platform.h"), os.path.join(os.path.dirname(HERE), "include", "icl_b200.h")]
OUT = os.path.join(HERE, "libicl_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "--extended-lambda", "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v",
]


def _stale(out, deps):
    return not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale(OUT, DEPS):
        return OUT
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc] + NVCC_FLAGS + ["-ccbin", "/usr/bin/g++", "-o", OUT, SRC]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode:
        raise RuntimeError("nvcc failed: " + " ".join(cmd))
    return OUT


def build_bi(force=False, verbose=False):
    """Device part of `-bi` (csrc/icl_blockinfo.cu -> libicl_bi_b200.so, include/icl_blockinfo.h)."""
    src = os.path.join(HERE, "csrc", "icl_blockinfo.cu")
    out = os.path.join(HERE, "libicl_bi_b200.so")
    if not force and not _stale(out, [src, os.path.join(os.path.dirname(HERE), "include", "icl_blockinfo.h")]):
        return out
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc] + NVCC_FLAGS + ["-ccbin", "/usr/bin/g++", "-o", out, src]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode:
        raise RuntimeError("nvcc failed: " + " ".join(cmd))
    return out


def build_npsort(force=False):
    """CPython helper that calls numpy's own argsort routine from a thread pool (csrc/npsort.c)."""
    import sysconfig
    import numpy
    src = os.path.join(HERE, "csrc", "npsort.c")
    out = os.path.join(HERE, "_npsort" + sysconfig.get_config_var("EXT_SUFFIX"))
    if not force and not _stale(out, [src]):
        return out
    import platform
    x86 = platform.machine().lower() in ("x86_64", "amd64", "i686")
    cmd = ["/usr/bin/gcc", "-O2"] + (["-mclflushopt"] if x86 else []) + ["-fPIC", "-shared", "-pthread", "-I",
           sysconfig.get_paths()["include"], "-I", numpy.get_include(), "-o", out, src]
    try:
        subprocess.check_call(cmd)
    except (subprocess.CalledProcessError, OSError) as e:
        # optional helper (GIL-free host rank): without it the engine calls numpy once per segment — same results
        sys.stderr.write(f"wgdi_b200.build: _npsort not built ({e}); the host rank falls back to per-segment numpy calls\n")
        return None
    return out


def build_host(force=False):
    """Native host helpers (m8 ingest, grading, writer): g++, no CUDA."""
    src = os.path.join(HERE, "csrc", "icl_host.cpp")
    out = os.path.join(HERE, "libicl_host.so")
    if not force and not _stale(out, [src, os.path.join(os.path.dirname(HERE), "include", "icl_host.h")]):
        return out
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-pthread", "-o", out, src])
    return out


def build_emu(outdir, force=False):
    os.makedirs(outdir, exist_ok=True)
    out = os.path.join(outdir, "libicl_b200_emu.so")
    if not force and not _stale(out, DEPS):
        return out
    cmd = ["/usr/bin/g++", "-x", "c++", "-std=c++17", "-O2", "-DICL_EMU", "-fPIC", "-shared", "-o", out, SRC]
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--emu":
        print(build_emu(sys.argv[2], force=True))
    else:
        print(build(force=True, verbose=True))
        print(build_bi(force=True, verbose=True))
        print(build_npsort(force=True))
        print(build_host(force=True))
