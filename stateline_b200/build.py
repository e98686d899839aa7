"""In-tree nvcc build of the engine (libslgpu.so) and the built-in likelihood plugin
(libslgpu_targets.so) for sm_100a.  The .so files are git-ignored but travel to the GPU box."""
import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
INCLUDE = os.path.join(ROOT, "include")
CSRC = os.path.join(PKG, "csrc")

ENGINE_SO = os.path.join(PKG, "libslgpu.so")
TARGETS_SO = os.path.join(PKG, "libslgpu_targets.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
# -fmad=false: products and sums round separately unless fma() is written out (bit parity with the oracle)
COMMON = ["-std=c++17", "-lineinfo", "-fmad=false", "-shared", "-Xcompiler", "-fPIC", "-I" + INCLUDE, "-I" + CSRC]


def nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the engine cannot be built (there is no CPU fallback)")
    return exe


def _stale(out, deps):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in deps)


def _headers():
    hs = [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE)]
    hs += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".hpp", ".cuh", ".h"))]
    return hs


def build_engine(force=False, verbose=False):
    src = os.path.join(CSRC, "engine.cu")
    if force or _stale(ENGINE_SO, [src] + _headers()):
        cmd = [nvcc()] + ARCH + COMMON + ["-O2", "-o", ENGINE_SO, src, "-ldl"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return ENGINE_SO


def build_plugin(src, out, force=False, verbose=False, extra=()):
    """Compile a likelihood plugin translation unit (it includes slgpu_device.cuh) into a .so."""
    if force or _stale(out, [src] + _headers()):
        cmd = [nvcc()] + ARCH + COMMON + ["-O3", "-o", out, src] + list(extra)
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return out


def build_targets(force=False, verbose=False):
    return build_plugin(os.path.join(PKG, "plugins", "builtin_targets.cu"), TARGETS_SO, force, verbose)


def build_all(force=False, verbose=False):
    return build_engine(force, verbose), build_targets(force, verbose)


if __name__ == "__main__":
    import sys

    print(build_all(force="--force" in sys.argv, verbose=True))
