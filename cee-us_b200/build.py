"""Builds the CUDA libraries (sm_100a only) in-tree with nvcc.  Used by __graft_entry__.build().

  libceeus_b200.so        the product: C ABI of include/ceeus_b200.h
  libceeus_b200_probe.so  test infrastructure: self-tests of the tcgen05 / TMEM building blocks (csrc/tc_probe.cu,
                          csrc/debug_tools.h), loaded only by tests/test_gpu_tc_probe.py and the *_diag.py scripts
  libceeus_b200_dbg.so    (python build.py --debug) the product sources with -DCEEUS_DEBUG_TOOLS: pipeline timeline probes of
                          the rollout kernel + epilogue micro-benchmark; never loaded by the package
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libceeus_b200.so")
PROBE_LIB = os.path.join(HERE, "libceeus_b200_probe.so")
DEBUG_LIB = os.path.join(HERE, "libceeus_b200_dbg.so")
SOURCES = ["ceeus_api.cu", "rollout_simt.cu", "icem_kernels.cu", "rollout_tc.cu", "rollout_mlp_tc.cu"]
PROBE_SOURCES = ["tc_probe.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ARCH + ["-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2"]


def _stale(lib):
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "ceeus_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def _build_lib(lib, sources, defines, tag, verbose):
    from concurrent.futures import ThreadPoolExecutor

    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", f"{tag}.o"))
        cmd = [nvcc] + CFLAGS + defines + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        return obj, subprocess.run(cmd, capture_output=True, text=True)

    with ThreadPoolExecutor(max_workers=len(sources)) as pool:
        results = list(pool.map(compile_one, sources))
    for obj, res in results:
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError(f"nvcc failed building {os.path.basename(lib)}")
        if verbose:
            print(res.stderr)
    link = subprocess.run([nvcc] + ARCH + ["-shared", "-Xcompiler", "-fPIC"] + [obj for obj, _ in results] + ["-o", lib],
                          capture_output=True, text=True)
    if link.returncode != 0:
        sys.stderr.write(link.stdout + link.stderr)
        raise RuntimeError(f"linking {os.path.basename(lib)} failed")
    return lib


def build(force=False, verbose=False, debug=False):
    """Compiles every translation unit for sm_100a in parallel (one nvcc per source) and links the shared libraries."""
    if force or _stale(LIB):
        _build_lib(LIB, SOURCES, [], "", verbose)
    if force or _stale(PROBE_LIB):
        _build_lib(PROBE_LIB, PROBE_SOURCES, ["-DCEEUS_DEBUG_TOOLS"], "_probe", verbose)
    if debug:
        _build_lib(DEBUG_LIB, SOURCES + PROBE_SOURCES, ["-DCEEUS_DEBUG_TOOLS"], "_dbg", verbose)
    return LIB


def build_variant(name, defines, verbose=False):
    """Development aid: the product sources with extra -D flags -> libceeus_b200_<name>.so (use with CEEUS_LIB=...)."""
    return _build_lib(os.path.join(HERE, f"libceeus_b200_{name}.so"), SOURCES, list(defines), f"_{name}", verbose)


if __name__ == "__main__":
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2:]))
        sys.exit(0)
    print(build(force="--no-force" not in sys.argv, verbose="-v" in sys.argv, debug="--debug" in sys.argv))
