"""Builds libsnapb200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libsnapb200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-fmad=false", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)) + [os.path.join(os.path.dirname(HERE), "include", "snapb200.h")]


def stale():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(s) > t for s in sources())


SO_PROFILE = os.path.join(HERE, "libsnapb200_prof.so")  # same library with the cycle accounting compiled in (scripts/profrun.py)


def build(force=False, verbose=False, profile=False):
    if profile:
        out, extra = SO_PROFILE, ["-DSNAPB200_PROFILE"]
        if not force and os.path.exists(out) and not any(os.path.getmtime(s) > os.path.getmtime(out) for s in sources()):
            return out
    else:
        out, extra = SO, []
        if not force and not stale():
            return SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", out, os.path.join(CSRC, "snapb200.cu")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("nvcc failed building libsnapb200.so")
    if verbose:
        print(r.stdout)
    return out


VARIANT_DIR = os.path.join(HERE, "variants")  # experiment builds (scripts/ab_variants.py); *.so is git-ignored and travels with gpurun


def build_variant(name, flags, verbose=False):
    """The same library compiled with extra flags (e.g. -DLANE_K_EXTRA=3) into variants/libsnapb200_<name>.so."""
    os.makedirs(VARIANT_DIR, exist_ok=True)
    out = os.path.join(VARIANT_DIR, f"libsnapb200_{name}.so")
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + list(flags) + (["-Xptxas", "-v"] if verbose else []) + ["-o", out, os.path.join(CSRC, "snapb200.cu")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError(f"nvcc failed building variant {name}")
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, profile="--profile" in sys.argv))
