"""The C++ twin of the reference seam (include/aha_b200.hpp: InferenceModel, GenerationContext, generate_generic,
B200Model) compiles against the C ABI and behaves like generate.rs:115-159 on a scripted model; the C header is plain
C11.  No GPU needed (the program only constructs a B200Model to check that the library's error comes back)."""
import os
import shutil
import subprocess

import pytest

from conftest import ROOT, has_gpu


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_cpp_host_mirror(lib_built, tmp_path):
    exe = tmp_path / "host_mirror_test"
    libdir = os.path.dirname(lib_built)
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp"), "-L", libdir, "-laha_b200", f"-Wl,-rpath,{libdir}", "-o", str(exe)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)] + ([] if has_gpu() else ["nogpu"]), capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "host mirror OK" in r.stdout


@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not available")
def test_c_header_is_plain_c11():
    r = subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c",
                        os.path.join(ROOT, "include", "aha_b200.h")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
