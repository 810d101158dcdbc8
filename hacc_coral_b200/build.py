"""Build libhaccsr.so (and the C++ facade library) in-tree with nvcc for sm_100a."""
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
HOST = os.path.join(_HERE, "host")
SOURCES = ["api.cu", "tree_build.cu", "walk.cu", "force.cu", "force_v0.cu", "force_v1.cu", "force_v2.cu", "force_v3.cu", "force_v4.cu",
           "refresh.cu", "exchange.cu", "cic.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; libhaccsr.so cannot be built")
    return nvcc


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile csrc/*.cu -> csrc/libhaccsr.so and host/RCBForceTree.cxx -> host/libhaccsr_facade.so."""
    out = os.path.join(CSRC, "libhaccsr.so")
    common = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "force_kernels.cuh"), os.path.join(_HERE, "..", "include", "haccsr.h")]
    objdir = os.path.join(CSRC, "build")
    os.makedirs(objdir, exist_ok=True)
    # one object per source, compiled in parallel and only when stale (force.cu alone takes minutes: 5 arithmetic variants)
    jobs, objs = [], []
    for src in SOURCES:
        obj = os.path.join(objdir, src[:-3] + ".o")
        objs.append(obj)
        if force or _stale(obj, [os.path.join(CSRC, src)] + common):
            cmd = [_nvcc()] + [f for f in NVCC_FLAGS if f != "-shared"] + ["-c", os.path.join(CSRC, src), "-o", obj]
            if verbose:
                cmd += ["-Xptxas", "-v"]
            jobs.append((src, subprocess.Popen(cmd)))
    failed = [src for src, pr in jobs if pr.wait() != 0]
    if failed:
        raise RuntimeError("nvcc failed on " + ", ".join(failed))
    if jobs or force or _stale(out, objs):
        subprocess.check_call([_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out] + objs + ["-ldl"])
    facade_src = os.path.join(HOST, "RCBForceTree.cxx")
    if os.path.exists(facade_src):
        fout = os.path.join(HOST, "libhaccsr_facade.so")
        fdeps = [facade_src, os.path.join(HOST, "RCBForceTree.h"), os.path.join(HOST, "ForceLaw.h"), out]
        if force or _stale(fout, [d for d in fdeps if os.path.exists(d)]):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I", HOST,
                                   "-I", os.path.join(_HERE, "..", "include"), "-o", fout, facade_src,
                                   "-L", CSRC, "-lhaccsr", "-Wl,-rpath,$ORIGIN/../csrc"])
    test_src = os.path.join(_HERE, "..", "tests", "cxx", "facade_test.cxx")
    fout = os.path.join(HOST, "libhaccsr_facade.so")
    if os.path.exists(test_src) and os.path.exists(fout):
        texe = os.path.join(_HERE, "..", "tests", "cxx", "facade_test")
        if force or _stale(texe, [test_src, fout]):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", HOST, "-I", os.path.join(_HERE, "..", "include"),
                                   "-o", texe, test_src, "-L", HOST, "-lhaccsr_facade", "-L", CSRC, "-lhaccsr",
                                   "-Wl,-rpath,$ORIGIN/../../hacc_coral_b200/host:$ORIGIN/../../hacc_coral_b200/csrc"])
    return out
