#!/usr/bin/env python
"""Build the REFERENCE's own pytrimal Cython extension with the CUDA compute platform.

What is compiled (all reference files are read where they lie under $REF; patched
copies and objects live only under integration/_build/, which is git-ignored but
travels to the GPU box):

  vendored trimAl      every source libtrimal is made of (vendor/trimal/scripts/CMake/
                       OBJ-LIB-creator.cmake, src/trimal/CMakeLists.txt), with pytrimal's
                       own four patches (patches/*.patch, src/CMakeLists.txt:5-21) and its
                       replacement reportsystem.cpp (Python exceptions / warnings)
  + CUDA platform      integration/patches/Manager.{h,cpp}.patch, pytrimal_b200/csrc/shim/
  pystreambuf          src/pystreambuf/*.cpp
  _trimal.pyx          src/pytrimal/_trimal.pyx + integration/patches/_trimal.pyx.patch,
                       include/trimal/statistics.pxd + statistics.pxd.patch, cythonized with
                       the compile-time constants of src/scripts/cmake/CythonExtension.cmake
                       plus CUDA_BUILD_SUPPORT=True
  scoring_matrices     the offline stand-in in integration/shims/ (the real dependency is
                       not installable here)

Result: integration/_build/pkg/{pytrimal,scoring_matrices}/ -- `import pytrimal` with
PYTHONPATH=integration/_build/pkg, `AutomaticTrimmer(platform="cuda")` etc.

This is not the reference's build system (no CMake, no scikit-build): a flat list of
g++ invocations with the flags that build uses for a Release configuration.
"""
from __future__ import annotations

import concurrent.futures as cf
import glob
import io
import os
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("REF", "/root/reference")
TRIMAL = os.path.join(REF, "vendor", "trimal")
CPUFEAT = os.path.join(TRIMAL, "vendor", "cpu_features")
OUT = os.path.join(HERE, "_build", "pytrimal_ext")
PKG = os.path.join(HERE, "_build", "pkg")
OBJ = os.path.join(OUT, "obj")
PATCHED = os.path.join(OUT, "trimal")       # patched trimAl headers / sources
CYINC = os.path.join(OUT, "cython_include")  # patched .pxd tree
CXX = os.environ.get("CXX", "g++")
CC = os.environ.get("CC", "gcc")
PYINC = sysconfig.get_paths()["include"]
EXT_SUFFIX = sysconfig.get_config_var("EXT_SUFFIX")


def run(cmd, **kw):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, **kw)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + "\n")
        raise SystemExit(f"build step failed: {cmd[0]} ... {cmd[-1]}")
    return r.stdout


def apply_patch(src, patch, dst):
    """The reference's own patch applier (src/scripts/apply_patch.py), used as a tool."""
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    run([sys.executable, os.path.join(REF, "src", "scripts", "apply_patch.py"),
         "--input", src, "--patch", patch, "--output", dst])


def unified_patch(src, patch, dst, strip_dir):
    """Apply one of OUR unified diffs (paths a/<rel>) with patch(1) on a private copy."""
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    shutil.copyfile(src, dst)
    run(["patch", "-s", "-p1", "-d", strip_dir, "-i", patch])


def prepare_sources():
    for d in (OBJ, PATCHED, PKG):
        os.makedirs(d, exist_ok=True)
    # pytrimal's own patches to vendored trimAl (src/CMakeLists.txt:5-21)
    for rel in ["source/Cleaner.cpp", "include/Statistics/similarityMatrix.h",
                "include/trimalManager.h", "include/Alignment/Alignment.h"]:
        apply_patch(os.path.join(TRIMAL, rel),
                    os.path.join(REF, "patches", os.path.basename(rel) + ".patch"),
                    os.path.join(PATCHED, rel))
    # pytrimal's replacement report system
    os.makedirs(os.path.join(PATCHED, "source"), exist_ok=True)
    shutil.copyfile(os.path.join(REF, "src", "trimal", "source", "reportsystem.cpp"),
                    os.path.join(PATCHED, "source", "reportsystem.cpp"))
    # the CUDA platform: Manager patches + shim header
    unified_patch(os.path.join(TRIMAL, "include/Statistics/Manager.h"),
                  os.path.join(HERE, "patches", "Manager.h.patch"),
                  os.path.join(PATCHED, "include/Statistics/Manager.h"), PATCHED)
    unified_patch(os.path.join(TRIMAL, "source/Statistics/Manager.cpp"),
                  os.path.join(HERE, "patches", "Manager.cpp.patch"),
                  os.path.join(PATCHED, "source/Statistics/Manager.cpp"), PATCHED)
    unified_patch(os.path.join(TRIMAL, "source/Alignment/Alignment.cpp"),
                  os.path.join(HERE, "patches", "Alignment.cpp.patch"),
                  os.path.join(PATCHED, "source/Alignment/Alignment.cpp"), PATCHED)
    # Cleaner.cpp: our patch goes on top of pytrimal's own (applied above)
    run(["patch", "-s", "-p1", "-d", PATCHED, "-i", os.path.join(HERE, "patches", "Cleaner.cpp.patch")])
    os.makedirs(os.path.join(PATCHED, "include/Platform/CUDA"), exist_ok=True)
    shutil.copyfile(os.path.join(ROOT, "pytrimal_b200/csrc/shim/CUDA.h"),
                    os.path.join(PATCHED, "include/Platform/CUDA/CUDA.h"))
    # Cython side: .pxd tree with the enumerator, .pyx with the platform plumbing
    if os.path.isdir(CYINC):
        shutil.rmtree(CYINC)
    shutil.copytree(os.path.join(REF, "include"), CYINC)
    os.makedirs(os.path.join(CYINC, "include", "trimal"), exist_ok=True)
    unified_patch(os.path.join(REF, "include/trimal/statistics.pxd"),
                  os.path.join(HERE, "patches", "statistics.pxd.patch"),
                  os.path.join(CYINC, "include/trimal/statistics.pxd"), CYINC)
    shutil.copyfile(os.path.join(CYINC, "include/trimal/statistics.pxd"),
                    os.path.join(CYINC, "trimal/statistics.pxd"))
    shutil.rmtree(os.path.join(CYINC, "include"))
    pyxdir = os.path.join(OUT, "src", "pytrimal")
    os.makedirs(pyxdir, exist_ok=True)
    unified_patch(os.path.join(REF, "src/pytrimal/_trimal.pyx"),
                  os.path.join(HERE, "patches", "_trimal.pyx.patch"),
                  os.path.join(pyxdir, "_trimal.pyx"), OUT)
    shutil.copyfile(os.path.join(REF, "src/pytrimal/_trimal.pxd"), os.path.join(pyxdir, "_trimal.pxd"))
    # pystreambuf's .pxd is cimported as `pystreambuf`
    os.makedirs(os.path.join(CYINC, "pystreambuf"), exist_ok=True)
    shutil.copyfile(os.path.join(REF, "src/pystreambuf/__init__.pxd"),
                    os.path.join(CYINC, "pystreambuf", "__init__.pxd"))


INCLUDES = None


def cxx_flags():
    return ["-O3", "-DNDEBUG", "-std=gnu++11", "-fPIC", "-w",
            "-DHAVE_AVX2=1", "-DHAVE_SSE2=1", "-DHAVE_CUDA=1",
            "-DFormatHandlerOverwrites=true", "-DFormatHandlerOverwritesOriginal=true",
            "-I" + os.path.join(PATCHED, "include"), "-I" + os.path.join(PATCHED, "include/Statistics"),
            "-I" + os.path.join(TRIMAL, "include"), "-I" + os.path.join(TRIMAL, "include/Statistics"),
            "-I" + os.path.join(CPUFEAT, "include"), "-I" + os.path.join(ROOT, "include"),
            "-I" + PYINC]


def compile_all():
    jobs = []

    def add(src, obj, extra=(), cc=False):
        obj = os.path.join(OBJ, obj)
        if os.path.exists(obj) and os.path.getmtime(obj) > os.path.getmtime(src):
            return obj
        if cc:
            cmd = [CC, "-O2", "-fPIC", "-w", "-DSTACK_LINE_READER_BUFFER_SIZE=1024",
                   "-DHAVE_STRONG_GETAUXVAL", "-DHAVE_DLFCN_H", "-I" + os.path.join(CPUFEAT, "include"),
                   "-I" + os.path.join(CPUFEAT, "include/internal"), "-c", src, "-o", obj]
        else:
            cmd = [CXX, *cxx_flags(), *extra, "-c", src, "-o", obj]
        jobs.append(cmd)
        return obj

    objs = []
    patched_src = {"source/Cleaner.cpp", "source/reportsystem.cpp", "source/Statistics/Manager.cpp",
                   "source/Alignment/Alignment.cpp"}
    core = ["source/Cleaner.cpp", "source/Alignment/Alignment.cpp", "source/Alignment/sequencesMatrix.cpp",
            "source/Statistics/similarityMatrix.cpp", "source/Statistics/Mold.cpp",
            "source/Statistics/Gaps.cpp", "source/Statistics/Manager.cpp",
            "source/Statistics/Similarity.cpp", "source/Statistics/Identity.cpp",
            "source/Statistics/Overlap.cpp", "source/Statistics/Consistency.cpp",
            "source/reportsystem.cpp", "source/reportMessages/infoMessages.cpp",
            "source/reportMessages/errorMessages.cpp", "source/reportMessages/warningMessages.cpp",
            "source/utils.cpp", "source/InternalBenchmarker.cpp", "source/trimalManager.cpp",
            "source/VCFHandler.cpp", "source/FormatHandling/BaseFormatHandler.cpp"]
    core += [os.path.relpath(p, TRIMAL)
             for p in sorted(glob.glob(os.path.join(TRIMAL, "source/FormatHandling/*_state.cpp")))]
    for rel in core:
        src = os.path.join(PATCHED if rel in patched_src else TRIMAL, rel)
        objs.append(add(src, rel.replace("/", "_")[:-4] + ".o"))
    objs.append(add(os.path.join(TRIMAL, "source/Platform/x86/AVX2.cpp"), "AVX2.o", ["-mavx2"]))
    objs.append(add(os.path.join(TRIMAL, "source/Platform/x86/SSE2.cpp"), "SSE2.o", ["-msse2"]))
    objs.append(add(os.path.join(ROOT, "pytrimal_b200/csrc/shim/CUDA.cpp"), "CUDA.o"))
    for f in ["impl_x86_linux_or_android", "filesystem", "stack_line_reader", "string_view", "hwcaps"]:
        objs.append(add(os.path.join(CPUFEAT, "src", f + ".c"), "cf_" + f + ".o", cc=True))
    for f in ["pyreadbuf", "pyreadintobuf", "pywritebuf"]:
        objs.append(add(os.path.join(REF, "src/pystreambuf", f + ".cpp"), "psb_" + f + ".o",
                        ["-I" + os.path.join(REF, "src/pystreambuf")]))
    with cf.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        list(ex.map(run, jobs))
    return objs


def cythonize_and_link(objs):
    version = "0.8.5"
    for line in open(os.path.join(REF, "pyproject.toml")):
        if line.startswith("version"):
            version = line.split("=")[1].strip().strip('"')
            break
    consts = {
        "SSE2_BUILD_SUPPORT": "True", "AVX2_BUILD_SUPPORT": "True", "NEON_BUILD_SUPPORT": "False",
        "CUDA_BUILD_SUPPORT": "True",
        "SYS_IMPLEMENTATION_NAME": sys.implementation.name,
        "SYS_VERSION_INFO_MAJOR": str(sys.version_info.major),
        "SYS_VERSION_INFO_MINOR": str(sys.version_info.minor),
        "TARGET_CPU": "x86_64", "TARGET_SYSTEM": "Linux", "SYS_BYTEORDER": sys.byteorder,
        "PYPY": "False", "PROJECT_VERSION": version, "DEFAULT_BUFFER_SIZE": str(io.DEFAULT_BUFFER_SIZE),
    }
    directives = ["-X", "cdivision=True", "-X", "nonecheck=False", "-X", "boundscheck=False",
                  "-X", "wraparound=False"]
    for k, v in consts.items():
        directives += ["-E", f"{k}={v}"]

    # the scoring_matrices stand-in
    shim_src = os.path.join(HERE, "shims", "scoring_matrices")
    shim_pkg = os.path.join(PKG, "scoring_matrices")
    os.makedirs(shim_pkg, exist_ok=True)
    for f in ["__init__.py", "lib.pxd"]:
        shutil.copyfile(os.path.join(shim_src, f), os.path.join(shim_pkg, f))
    shim_c = os.path.join(OUT, "scoring_matrices_lib.c")
    run([sys.executable, "-m", "cython", os.path.join(shim_src, "lib.pyx"), "-3",
         "--output-file", shim_c, "-I", os.path.join(HERE, "shims")], cwd=os.path.join(HERE, "shims"))
    run([CC, "-O2", "-fPIC", "-w", "-shared", "-I" + PYINC, shim_c, "-o",
         os.path.join(shim_pkg, "lib" + EXT_SUFFIX)])

    # _trimal
    pyx = os.path.join(OUT, "src", "pytrimal", "_trimal.pyx")
    cpp = os.path.join(OUT, "_trimal.cpp")
    run([sys.executable, "-m", "cython", pyx, "--output-file", cpp, "--cplus",
         "-I", CYINC, "-I", os.path.join(HERE, "shims"), *directives])
    flags = [f for f in cxx_flags() if f != "-std=gnu++11"] + ["-std=gnu++17"]
    ext_obj = os.path.join(OBJ, "_trimal.o")
    run([CXX, *flags, "-DCYTHON_WITHOUT_ASSERTIONS=1", "-DHAVE_PYINTERPRETERSTATE_GETID",
         "-include", os.path.join(REF, "src/scripts/cmake/pystate_patch.h"),
         "-I" + os.path.join(REF, "src/pystreambuf"), "-I" + os.path.join(OUT, "src", "pytrimal"),
         "-c", cpp, "-o", ext_obj])
    pkg = os.path.join(PKG, "pytrimal")
    os.makedirs(pkg, exist_ok=True)
    run([CXX, "-shared", "-o", os.path.join(pkg, "_trimal" + EXT_SUFFIX), ext_obj, *objs,
         "-L" + os.path.join(ROOT, "pytrimal_b200"), "-ltrimal_cuda",
         "-Wl,-rpath,$ORIGIN/../../../../pytrimal_b200", "-lm"])
    # the pure-Python side of the package and its test data (symlinks dereferenced)
    for f in ["__init__.py", "py.typed", "_trimal.pyi"]:
        shutil.copyfile(os.path.join(REF, "src/pytrimal", f), os.path.join(pkg, f))
    # the type stub learns the new platform literal
    run(["patch", "-s", "-p3", "-d", pkg, "-i", os.path.join(HERE, "patches", "_trimal.pyi.patch")])
    tests = os.path.join(pkg, "tests")
    if os.path.isdir(tests):
        shutil.rmtree(tests)
    shutil.copytree(os.path.join(REF, "src/pytrimal/tests"), tests, symlinks=False)


def main():
    if not os.path.isdir(REF):
        print("integration/build_pytrimal.py: no reference tree at", REF, "- nothing to build")
        return
    if not os.path.exists(os.path.join(ROOT, "pytrimal_b200", "libtrimal_cuda.so")):
        raise SystemExit("build pytrimal_b200/libtrimal_cuda.so first (python -m pytrimal_b200.build)")
    prepare_sources()
    objs = compile_all()
    cythonize_and_link(objs)
    print(os.path.join(PKG, "pytrimal"))


if __name__ == "__main__":
    main()
