"""The C++ host mirror (include/biogo_align.hpp) against the reference's example tests, compiled
with g++ and linked to the CPU-emulated build of the C ABI here (CUDA build on the GPU box)."""
import os
import subprocess

import pytest

import helpers

SRC = os.path.join(helpers.ROOT, "tests", "cpp", "mirror_test.cpp")


def _build_and_run(libdir, libname, tmp_path):
    exe = str(tmp_path / "mirror_test")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I" + os.path.join(helpers.ROOT, "include"), SRC, "-o", exe,
                           "-L" + libdir, "-l" + libname, "-Wl,-rpath," + libdir])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "all ok" in out.stdout


def test_cpp_mirror_on_emulated_kernels(built, tmp_path):
    _build_and_run(os.path.dirname(helpers.EMU_LIB), "biogoalign_emu", tmp_path)


@pytest.mark.gpu
def test_cpp_mirror_on_gpu(tmp_path):
    _build_and_run(os.path.join(helpers.ROOT, "biogo_b200", "lib"), "biogoalign", tmp_path)
