"""Link-level drop-in check of libapm_refcompat.so against the reference's OWN host sources.

DESIGN.md claims that the reference's .c files link against libapm_refcompat.so unchanged, in place of its three .cu
objects (Makefile:45-56).  MPI is absent from this image, so the files are compiled to objects against a
declarations-only tests/stub_mpi/mpi.h and inspected with nm:

  * the only undefined symbols of main.o + patterns_over_ranks.o + database_over_ranks.o + utils.o that are not libc /
    libm / OpenMP / MPI are the six GPU entry points (patterns_over_ranks.c:33-36, database_over_ranks.c:18-22,
    main.c:18-19);
  * libapm_refcompat.so defines each of them as a C symbol;
  * the four objects link into an executable against the library with --no-undefined (MPI supplied by an aborting
    stand-in, tests/stub_mpi/mpi_stub.c), and its dynamic symbol table binds the six names to the library;
  * the reference's own prototypes and include/apm_refcompat.h agree (one translation unit sees both).

Nothing of the reference is copied into the repo; the sources are read where they lie.  Skipped when /root/reference
is absent (the GPU box)."""
import os
import re
import subprocess

import pytest

import apm_b200
from apm_b200 import refcompat

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUB = os.path.join(ROOT, "tests", "stub_mpi")
GPU_ENTRY_POINTS = {"invoke_kernel", "write_kernel_result", "initializeGPU", "getGPUResult", "getDeviceCount", "setDevice"}

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src")), reason="reference sources not present")


def _nm(args, path):
    out = subprocess.run(["nm"] + args + [path], capture_output=True, text=True, check=True).stdout
    return {line.split()[-1].split("@")[0] for line in out.splitlines() if line.strip()}


@pytest.fixture(scope="module")
def objects(tmp_path_factory):
    d = tmp_path_factory.mktemp("reflink")
    objs = {}
    for name in ("main", "patterns_over_ranks", "database_over_ranks", "utils"):
        o = str(d / f"{name}.o")
        subprocess.run(["gcc", "-c", "-O1", "-w", "-fopenmp", "-DUSE_GPU_FLAG", f"-I{REF}/include", f"-I{STUB}",
                        f"{REF}/src/{name}.c", "-o", o], check=True)
        objs[name] = o
    return d, objs


def test_undefined_symbols_are_exactly_the_gpu_entry_points(objects):
    _, objs = objects
    defined, undefined = set(), set()
    for o in objs.values():
        defined |= _nm(["--defined-only"], o)
        undefined |= _nm(["-u"], o)
    undefined -= defined
    system = _system_symbols()
    assert "printf" in system and "GOMP_parallel" in system
    other = {s for s in undefined if not (s.startswith(("MPI_", "ompi_")) or s in system)}
    assert other == GPU_ENTRY_POINTS, other ^ GPU_ENTRY_POINTS


def _system_symbols():
    """everything libc, libm and libgomp export"""
    syms = set()
    for lib in ("libc.so.6", "libm.so.6", "libgomp.so.1"):
        path = subprocess.run(["gcc", f"-print-file-name={lib}"], capture_output=True, text=True).stdout.strip()
        if os.path.isabs(path) and os.path.exists(path):
            syms |= _nm(["-D", "--defined-only"], path)
    return syms


def test_library_defines_every_entry_point():
    exported = _nm(["-D", "--defined-only"], refcompat.LIB_PATH)
    assert GPU_ENTRY_POINTS <= exported, GPU_ENTRY_POINTS - exported


def test_reference_objects_link_against_the_library(objects):
    """main.o + patterns_over_ranks.o + database_over_ranks.o + utils.o + libapm_refcompat.so + an MPI stand-in close
    the link with --no-undefined: the library supplies everything the reference's three .cu objects supplied."""
    d, objs = objects
    exe = str(d / "apm_parallel_linked")
    pkg = os.path.dirname(refcompat.LIB_PATH)
    stub = str(d / "mpi_stub.o")
    subprocess.run(["gcc", "-c", "-O1", f"-I{STUB}", os.path.join(STUB, "mpi_stub.c"), "-o", stub], check=True)
    subprocess.run(["gcc", "-fopenmp", "-o", exe] + list(objs.values()) + [stub] +
                   [f"-L{pkg}", "-lapm_refcompat", "-lapm_b200", "-lm", f"-Wl,-rpath,{pkg}", "-Wl,--no-undefined"], check=True)
    dyn_undefined = _nm(["-D", "-u"], exe)
    assert GPU_ENTRY_POINTS <= dyn_undefined  # bound at load time ...
    needed = subprocess.run(["readelf", "-d", exe], capture_output=True, text=True, check=True).stdout
    assert "libapm_refcompat.so" in needed      # ... by this library


def test_prototypes_agree_with_the_reference(objects, tmp_path):
    """one translation unit with the reference's own declarations (read from its sources) and ours"""
    decls = []
    for fname, names in (("patterns_over_ranks.c", ("invoke_kernel", "write_kernel_result")),
                         ("database_over_ranks.c", ("initializeGPU", "getGPUResult")),
                         ("main.c", ("getDeviceCount", "setDevice"))):
        src = open(os.path.join(REF, "src", fname)).read()
        for n in names:
            m = re.search(r"^[a-z][\w \*]*?\b" + n + r"\s*\([^;{]*\)\s*;", src, re.M | re.S)
            assert m, (fname, n)
            decls.append(m.group(0))
    tu = tmp_path / "protos.c"
    tu.write_text('#include "approaches.h"\n' + "\n".join(decls) + '\n#include "apm_refcompat.h"\n'
                  "#define patterns_over_ranks_hybrid_check 1\n"
                  "int (*p1)(int, char **, int, int, int) = apm_patterns_over_ranks_hybrid;\n"
                  "int (*p2)(int, char **, int, int, int) = patterns_over_ranks_hybrid;\n"
                  "int (*p3)(int, char **, int, int, int) = apm_database_over_ranks;\n"
                  "int (*p4)(int, char **, int, int, int) = database_over_ranks;\n")
    subprocess.run(["gcc", "-c", "-Wall", "-Werror", f"-I{REF}/include", f"-I{ROOT}/include", str(tu), "-o",
                    str(tmp_path / "protos.o")], check=True)
