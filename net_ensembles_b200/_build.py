"""Builds the native artefacts in-tree (no JIT cache): the CUDA library with the C ABI and,
for tests/bench only, the CPU oracle.  Called by __graft_entry__.build(); nvcc cross-compiles
sm_100a without a GPU."""
import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
LIB_DIR = os.path.join(PKG, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libnet_ensembles_cuda.so")
HOST_LIB_PATH = os.path.join(LIB_DIR, "libnec_host.so")     # graph / ensemble mirror: no CUDA, not linked to the CUDA library
ORACLE_PATH = os.path.join(ROOT, "oracle", "build", "liboracle.so")

HOST_HEADERS = [os.path.join(PKG, "csrc", "host", h) for h in ("nec_graphs.hpp", "nec_rand.hpp", "nec_handle.hpp")]
HOST_SOURCES = [os.path.join(PKG, "csrc", "host", "host_api.cpp")]
CUDA_SOURCES = [os.path.join(PKG, "csrc", "nec_api.cu"), os.path.join(PKG, "csrc", "host", "host_bridge.cpp")]
CUDA_DEPS = CUDA_SOURCES + HOST_HEADERS + [
    os.path.join(PKG, "csrc", "nec_kernels.cuh"), os.path.join(PKG, "csrc", "nec_small.cuh"),
    os.path.join(PKG, "csrc", "nec_mid.cuh"), os.path.join(PKG, "csrc", "nec_wide.cuh"), os.path.join(PKG, "csrc", "nec_sample.cuh"), os.path.join(PKG, "csrc", "nec_chain.cuh"), os.path.join(PKG, "csrc", "nec_multi.hpp"),
    os.path.join(ROOT, "include", "net_ensembles_cuda.h"),
]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fopenmp", "-shared"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; net_ensembles_b200 cannot be built (there is no CPU fallback)")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_cuda(force=False, verbose=False):
    if not force and not _stale(LIB_PATH, CUDA_DEPS):
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + CUDA_SOURCES + ["-o", LIB_PATH]
    env = dict(os.environ)
    env.pop("CC", None), env.pop("CXX", None)   # the image's $CC wrapper is not a valid nvcc host compiler
    res = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), res.stderr))
    if verbose:
        print(res.stderr)
    return LIB_PATH


def build_variant(tag, defines):
    """A/B build of the CUDA library with extra -D flags (scripts/exp_r2.py --lib <tag>); lives next to the product library."""
    out = os.path.join(LIB_DIR, tag + ".so")
    cmd = [_nvcc()] + NVCC_FLAGS + ["-D" + d for d in defines] + CUDA_SOURCES + ["-o", out]
    env = dict(os.environ)
    env.pop("CC", None), env.pop("CXX", None)
    res = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), res.stderr))
    return out


def build_host(force=False):
    if not force and not _stale(HOST_LIB_PATH, HOST_SOURCES + HOST_HEADERS):
        return HOST_LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    cxx = shutil.which("g++") or "g++"
    cmd = [cxx, "-O3", "-std=c++17", "-fPIC", "-fopenmp", "-shared"] + HOST_SOURCES + ["-o", HOST_LIB_PATH]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("host library build failed:\n%s\n%s" % (" ".join(cmd), res.stderr))
    return HOST_LIB_PATH


def build_oracle(force=False):
    src = os.path.join(ROOT, "oracle", "ref_vertex_load.c")
    if not force and not _stale(ORACLE_PATH, [src, os.path.join(ROOT, "oracle", "ref_measures.c"), os.path.join(ROOT, "oracle", "Makefile")]):
        return ORACLE_PATH
    res = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")] + (["-B"] if force else []),
                         capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + res.stdout + res.stderr)
    return ORACLE_PATH


def build_all(force=False, verbose=False):
    build_host(force)
    return build_cuda(force, verbose), build_oracle(force)


if __name__ == "__main__":
    print(build_all(force=True, verbose=True))
