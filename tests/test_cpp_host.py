"""The C++ host mirror of the reference interface (bbcat-dsp_b200/host/*.h) compiles against the C ABI without a
GPU, and (with -m gpu) a client program written like reference client code produces the reference's known answers."""
import os
import shutil
import subprocess

import pytest

from conftest import ROOT

SRC = os.path.join(ROOT, "tests", "cpp", "host_api_test.cpp")
LIBDIR = os.path.join(ROOT, "bbcat-dsp_b200")


def build(tmp_path):
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    exe = str(tmp_path / "host_api_test")
    subprocess.check_call([cxx, "-std=c++11", "-Wall", "-Werror", "-I" + os.path.join(LIBDIR, "host"), SRC, "-o", exe,
                           "-L" + LIBDIR, "-lbbx", "-Wl,-rpath," + LIBDIR])
    return exe


def test_host_headers_compile_and_link(tmp_path):
    assert os.path.exists(build(tmp_path))


@pytest.mark.gpu
def test_host_client_program(tmp_path, bbx):
    out = subprocess.run([build(tmp_path)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "PASS" in out.stdout
