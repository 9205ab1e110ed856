"""In-tree build of the native libraries (nvcc cross-compiles sm_100a without a GPU).

    polar_b200/lib/libpolar_b200.so   CUDA kernels + the C ABI of include/polar_b200.h
    polar_b200/lib/libpolar_host.so   the C++ drop-in `PolarCode` class + ctypes wrappers

`python -m polar_b200.build` rebuilds what is stale; `--force` rebuilds everything.
The .so files are git-ignored but travel to the GPU box with the tree.
"""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "lib")
INC = os.path.join(ROOT, "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xfatbin", "-compress-all", "-I" + INC, "-I" + CSRC,
]
FAST_PARTS = 4          # fast_parts.cu is compiled once per quarter of the variant table (fast_variants.cuh)


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA library cannot be built (there is no CPU fallback)")


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(force=False, verbose=False, out_dir=None, extra_flags=()):
    """Compile the CUDA translation units in parallel (polar_b200.cu + fast_parts.cu x FAST_PARTS), link
    libpolar_b200.so, then libpolar_host.so. `out_dir` / `extra_flags` build a variant of the libraries elsewhere
    (A/B runs: POLAR_B200_LIB_DIR)."""
    lib = out_dir or LIB
    obj = os.path.join(lib, "obj")
    os.makedirs(obj, exist_ok=True)
    hdrs = [os.path.join(INC, "polar_b200.h")] + [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))
                                                   if f.endswith((".h", ".cuh"))]
    dev_so = os.path.join(lib, "libpolar_b200.so")
    dev_src = [os.path.join(CSRC, "polar_b200.cu"), os.path.join(CSRC, "fast_parts.cu")]
    if force or _stale(dev_so, dev_src + hdrs):
        flags = NVCC_FLAGS + list(extra_flags) + (["-Xptxas", "-v"] if verbose else [])
        jobs = [([_nvcc()] + flags + ["-c", dev_src[0], "-o", os.path.join(obj, "polar_b200.o")], "polar_b200.cu")]
        for k in range(FAST_PARTS):
            jobs.append(([_nvcc()] + flags + ["-DPOLAR_PART=%d" % k, "-c", dev_src[1], "-o", os.path.join(obj, "fast_part%d.o" % k)],
                         "fast_parts.cu part %d" % k))
        procs = [(subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True), name) for cmd, name in jobs]
        # the opt-in min-sum build of a few variants (its own namespace and table, see fast_parts.cu)
        jobs.append(([_nvcc()] + flags + ["-DPOLAR_MINSUM=1", "-DPOLAR_FAST_NS=fastms", "-c", dev_src[1], "-o",
                                          os.path.join(obj, "fast_ms.o")], "fast_parts.cu min-sum build"))
        procs.append((subprocess.Popen(jobs[-1][0], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True), jobs[-1][1]))
        failed = []
        for pr, name in procs:
            out, _ = pr.communicate()
            if verbose or pr.returncode != 0:
                sys.stdout.write(out)
            if pr.returncode != 0:
                failed.append(name)
        if failed:
            raise RuntimeError("nvcc failed for: " + ", ".join(failed))
        objs = [os.path.join(obj, "polar_b200.o")] + [os.path.join(obj, "fast_part%d.o" % k) for k in range(FAST_PARTS)]
        objs.append(os.path.join(obj, "fast_ms.o"))
        subprocess.check_call([_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a"] + objs + ["-o", dev_so])
    host_so = os.path.join(lib, "libpolar_host.so")
    host_src = [os.path.join(CSRC, "PolarCode.cpp")]
    if force or _stale(host_so, host_src + hdrs + [dev_so]):
        cmd = [_nvcc(), "-O2", "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-I" + INC, "-I" + CSRC] + host_src + [
            "-L" + lib, "-lpolar_b200", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN", "-o", host_so]
        subprocess.check_call(cmd)
    return dev_so, host_so


def build_acceptance(ref_dir="/root/reference/PolarC"):
    """The reference's UNMODIFIED main.cpp compiled against this repo's PolarCode.h (the drop-in
    test of SURVEY.md section 4 item 3). Output goes next to the compiled reference in oracle/_ref/
    (git-ignored; the source is compiled where it lies, never copied)."""
    main_cpp = os.path.join(ref_dir, "main.cpp")
    if not os.path.exists(main_cpp):
        return None
    out_dir = os.path.join(ROOT, "oracle", "_ref")
    os.makedirs(out_dir, exist_ok=True)
    exe = os.path.join(out_dir, "polar_b200_main")
    srcs = [main_cpp, os.path.join(CSRC, "PolarCode.h"), os.path.join(LIB, "libpolar_host.so")]
    if _stale(exe, srcs):
        # -I CSRC first so that `#include "PolarCode.h"` resolves to the drop-in header; nvcc's
        # host compiler looks in the including file's own directory first for quoted includes, so
        # the file is fed through stdin-less indirection: compile a one-line TU that includes it.
        shim = os.path.join(out_dir, "main_shim.cpp")
        with open(shim, "w") as f:
            f.write('#include "PolarCode.h"\n#define POLARC_POLARCODE_H\n#include "%s"\n' % main_cpp)
        rpath = os.path.relpath(LIB, out_dir)
        cmd = ["g++", "-std=c++11", "-O2", "-I" + CSRC, shim, "-L" + LIB, "-lpolar_host", "-lpolar_b200",
               "-Wl,-rpath,$ORIGIN/" + rpath, "-o", exe]
        subprocess.check_call(cmd)
    return exe


if __name__ == "__main__":
    # python -m polar_b200.build [--force] [-v] [--dir <name under polar_b200/>] [-- extra nvcc flags]
    argv = sys.argv[1:]
    extra = []
    if "--" in argv:
        extra = argv[argv.index("--") + 1:]
        argv = argv[:argv.index("--")]
    out_dir = os.path.join(PKG, argv[argv.index("--dir") + 1]) if "--dir" in argv else None
    build(force="--force" in argv, verbose="-v" in argv, out_dir=out_dir, extra_flags=extra)
    if out_dir is None:
        build_acceptance()
    print("built:", sorted(f for f in os.listdir(out_dir or LIB) if f.endswith(".so")))
