#!/usr/bin/env python3
"""Build the reference's own CPU backend (torchsparse 1.4.0) into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing under taseg_b200/ may import this.

The reference ships its native backend as source inside
/root/reference/package/torchsparse.zip (member torchsparse/torchsparse/backend/**)
and its only native dependency (Google sparsehash, header-only) inside
/root/reference/package/sparsehash.zip.  We do NOT run the reference's build system
(setup.py / autoconf): this recipe extracts the handful of *_cpu.cpp files to a
scratch directory, writes the tiny platform config header sparsehash's autoconf
would generate, and calls g++ directly.  Only the resulting shared object is kept:

    oracle/_ref/backend.cpython-*.so   pybind11 module exposing the reference's
                                       10 CPU entry points (pybind_cpu.cpp:18-29)

oracle/_ref/ is git-ignored (never commit reference code) but NOT gpurun-ignored,
so the built .so travels to the GPU box, where /root/reference does not exist.
"""
import glob
import os
import shutil
import subprocess
import sys
import sysconfig
import tempfile
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF_PKG = "/root/reference/package"

SPARSECONFIG = """\
// platform config for Google sparsehash (what ./configure would emit on linux/gcc)
#define GOOGLE_NAMESPACE ::google
#define HASH_FUN_H <functional>
#define HASH_NAMESPACE std
#define HAVE_INTTYPES_H 1
#define HAVE_LONG_LONG 1
#define HAVE_MEMCPY 1
#define HAVE_STDINT_H 1
#define HAVE_SYS_TYPES_H 1
#define HAVE_UINT16_T 1
#define HAVE_U_INT16_T 1
#define SPARSEHASH_HASH HASH_NAMESPACE::hash
#define _END_GOOGLE_NAMESPACE_ }
#define _START_GOOGLE_NAMESPACE_ namespace google {
"""


def so_name() -> str:
    return "backend" + sysconfig.get_config_var("EXT_SUFFIX")


def built() -> bool:
    return os.path.exists(os.path.join(OUT, so_name()))


def build(force: bool = False) -> str:
    target = os.path.join(OUT, so_name())
    if built() and not force:
        return target
    if not os.path.isdir(REF_PKG):
        raise RuntimeError("reference sources not present (only the prebuilt .so travels)")
    import torch
    from torch.utils import cpp_extension

    os.makedirs(OUT, exist_ok=True)
    scratch = tempfile.mkdtemp(prefix="taseg_ref_build_")
    try:
        zipfile.ZipFile(os.path.join(REF_PKG, "torchsparse.zip")).extractall(scratch)
        zipfile.ZipFile(os.path.join(REF_PKG, "sparsehash.zip")).extractall(scratch)
        sh_src = os.path.join(scratch, "sparsehash-master", "src")
        with open(os.path.join(sh_src, "sparsehash", "internal", "sparseconfig.h"), "w") as f:
            f.write(SPARSECONFIG)
        be = os.path.join(scratch, "torchsparse", "torchsparse", "backend")
        srcs = sorted(set(glob.glob(os.path.join(be, "**", "*_cpu.cpp"), recursive=True)))
        incs = cpp_extension.include_paths() + [sysconfig.get_paths()["include"], sh_src]
        libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
        cc = ["g++", "-O3", "-fopenmp", "-fPIC", "-std=c++17", "-w",
              "-DTORCH_EXTENSION_NAME=backend", "-DTORCH_API_INCLUDE_EXTENSION_H",
              f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"]
        cc += [f"-I{i}" for i in incs]
        objs = [os.path.join(scratch, f"o{i}.o") for i in range(len(srcs))]
        procs = [subprocess.Popen(cc + ["-c", s, "-o", o]) for s, o in zip(srcs, objs)]
        if any(p.wait() != 0 for p in procs):
            raise RuntimeError("reference CPU backend failed to compile")
        subprocess.check_call(["g++", "-shared", "-fopenmp"] + objs + [
            f"-L{libdir}", "-ltorch", "-ltorch_cpu", "-lc10", "-ltorch_python",
            f"-Wl,-rpath,{libdir}", "-lgomp", "-o", target])
    finally:
        shutil.rmtree(scratch, ignore_errors=True)
    return target


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
