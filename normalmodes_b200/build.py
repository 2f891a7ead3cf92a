"""Build libnm_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m normalmodes_b200.build [--force]
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libnm_b200.so")
SOURCES = ["nm_runtime.cu", "nm_parcsr.cu", "nm_slab.cu", "nm_chebiter.cu", "nm_ops.cu", "nm_lanczos.cu", "nm_hostmath.cpp",
           "nm_pevsl_f90.cpp", "nm_assembly.cu", "nm_pattern.cpp"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nccl_paths():
    """Prefer the NCCL torch itself loads (nvidia-nccl-cu12 wheel) so one process holds one libnccl."""
    site = sysconfig.get_paths()["purelib"]
    inc = os.path.join(site, "nvidia", "nccl", "include")
    lib = os.path.join(site, "nvidia", "nccl", "lib")
    if os.path.exists(os.path.join(lib, "libnccl.so.2")):
        return inc, lib
    return "/usr/include", "/usr/lib/x86_64-linux-gnu"


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.abspath(__file__)]
    deps += [os.path.join(HERE, "..", "include", f) for f in os.listdir(os.path.join(HERE, "..", "include"))]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    inc, libdir = _nccl_paths()
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for s in srcs:
        o = os.path.join(HERE, "build", os.path.basename(s) + ".o")
        cmd = ["nvcc", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-x", "cu", *ARCH,
               "-I", inc, "-I", os.path.join(HERE, "..", "include"), "-c", s, "-o", o]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    ok = True
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write("---- %s\n%s\n" % (os.path.basename(s), out))
        ok = ok and p.returncode == 0
    if not ok:
        raise RuntimeError("nvcc failed building libnm_b200.so")
    link = ["nvcc", "-shared", *ARCH, "-o", LIB, *objs, "-L", libdir, "-l:libnccl.so.2", "-lcudart",
            "-Xlinker", "-rpath", "-Xlinker", libdir, "-Xlinker", "-rpath", "-Xlinker", "/usr/lib/x86_64-linux-gnu"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link failed for libnm_b200.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
