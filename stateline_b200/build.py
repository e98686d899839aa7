"""In-tree nvcc build of the engine (libslgpu.so) and the built-in likelihood plugin
(libslgpu_targets.so) for sm_100a.  The .so files are git-ignored but travel to the GPU box."""
import fcntl
import hashlib
import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
INCLUDE = os.path.join(ROOT, "include")
CSRC = os.path.join(PKG, "csrc")

ENGINE_SO = os.path.join(PKG, "libslgpu.so")
TARGETS_SO = os.path.join(PKG, "libslgpu_targets.so")
TARGETS_TILE_SO = os.path.join(PKG, "libslgpu_targets_tile.so")  # the same targets on the opt-in tile kernel
CLI_EXE = os.path.join(PKG, "stateline-gpu")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
# -fmad=false: products and sums round separately unless fma() is written out (bit parity with the oracle)
COMMON = ["-std=c++17", "-lineinfo", "-fmad=false", "-shared", "-Xcompiler", "-fPIC", "-I" + INCLUDE, "-I" + CSRC]


def nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the engine cannot be built (there is no CPU fallback)")
    return exe


def _digest(deps, flags):
    """Content hash of the sources and flags: file times do not survive the copy to the GPU box."""
    h = hashlib.sha256(" ".join(flags).replace(ROOT, ".").encode())  # the checkout's own path is not part of the build
    for d in sorted(deps):
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _stale(out, digest):
    try:
        return not os.path.exists(out) or open(out + ".srchash").read().strip() != digest
    except OSError:
        return True


def _compile(cmd, out, digest, verbose):
    """Build under an exclusive lock (ranks of one node may race here) and record the source hash."""
    with open(out + ".lock", "w") as lk:
        fcntl.flock(lk, fcntl.LOCK_EX)
        if _stale(out, digest):
            tmp = out + ".tmp.%d" % os.getpid()
            full = [c if c != out else tmp for c in cmd]
            if verbose:
                print(" ".join(full))
            subprocess.check_call(full)
            os.replace(tmp, out)
            with open(out + ".srchash", "w") as f:
                f.write(digest)


def _headers():
    hs = [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE)]
    hs += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".hpp", ".cuh", ".h"))]
    return hs


def _plugin_headers():
    """What a likelihood plugin compiles: the device headers (slgpu.h is the host API and no part of a plugin build)."""
    return [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE) if f != "slgpu.h"]


def build_engine(force=False, verbose=False):
    src = os.path.join(CSRC, "engine.cu")
    flags = ARCH + COMMON + ["-O2", "-ldl"]
    digest = _digest([src] + _headers(), flags)
    if force and os.path.exists(ENGINE_SO + ".srchash"):
        os.remove(ENGINE_SO + ".srchash")
    if _stale(ENGINE_SO, digest):
        _compile([nvcc()] + ARCH + COMMON + ["-O2", "-o", ENGINE_SO, src, "-ldl"], ENGINE_SO, digest, verbose)
    return ENGINE_SO


def build_plugin(src, out, force=False, verbose=False, extra=()):
    """Compile a likelihood plugin translation unit (it includes slgpu_device.cuh) into a .so."""
    flags = ARCH + COMMON + ["-O3"] + list(extra)
    digest = _digest([src] + _plugin_headers(), flags)
    if force and os.path.exists(out + ".srchash"):
        os.remove(out + ".srchash")
    if _stale(out, digest):
        _compile([nvcc()] + ARCH + COMMON + ["-O3", "-o", out, src] + list(extra), out, digest, verbose)
    return out


def build_targets(force=False, verbose=False):
    return build_plugin(os.path.join(PKG, "plugins", "builtin_targets.cu"), TARGETS_SO, force, verbose)


def build_targets_tile(force=False, verbose=False):
    """The built-in targets with -DSLGPU_TILE_KERNEL: the step kernel that keeps a stack group's factors in shared memory
    for the launch (include/slgpu_step_tile.cuh).  Opt-in (slower than the default); built so that its tests travel."""
    return build_plugin(os.path.join(PKG, "plugins", "builtin_targets.cu"), TARGETS_TILE_SO, force, verbose, extra=("-DSLGPU_TILE_KERNEL",))


def build_cli(force=False, verbose=False, out=CLI_EXE):
    """stateline-gpu: the `stateline -c config.json` server binary for the GPU path (plain C++ over the C-ABI)."""
    build_engine(force, verbose)
    src = os.path.join(PKG, "cli", "stateline_gpu.cpp")
    cxx = shutil.which("g++") or "g++"
    flags = ["-std=c++17", "-O2", "-I" + INCLUDE, "-I" + CSRC]
    digest = _digest([src, os.path.join(INCLUDE, "slgpu.h"), os.path.join(CSRC, "json_min.hpp")], flags)
    if force and os.path.exists(out + ".srchash"):
        os.remove(out + ".srchash")
    if _stale(out, digest):
        _compile([cxx] + flags + [src, "-L" + PKG, "-lslgpu", "-Wl,-rpath," + PKG, "-Wl,-rpath,$ORIGIN", "-o", out], out, digest, verbose)
    return out


def build_all(force=False, verbose=False):
    build_targets_tile(force, verbose)
    return build_engine(force, verbose), build_targets(force, verbose), build_cli(force, verbose)


if __name__ == "__main__":
    import sys

    print(build_all(force="--force" in sys.argv, verbose=True))
