"""The C++ facade include/teaser/registration.h: the reference driver's call sequence
(examples/teaser_cpp_ply/PSULVSB.cc:291-331) compiles and links against libpsulvsb_b200.so; on a GPU
box it also has to register the synthetic pair correctly."""
import os
import subprocess

import pytest

import psulvsb_b200  # noqa: F401
from psulvsb_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_snippet(tmp_path):
    exe = str(tmp_path / "driver_snippet")
    libdir = os.path.dirname(capi.LIB_PATH)
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"),
           "-I", os.path.join(ROOT, "tests", "shim"), os.path.join(ROOT, "tests", "cpp", "driver_snippet.cc"),
           "-o", exe, "-L", libdir, "-l:libpsulvsb_b200.so", "-Wl,-rpath," + libdir]
    subprocess.check_call(cmd)
    return exe


def test_facade_compiles_links_and_fails_loudly_without_gpu(tmp_path):
    exe = build_snippet(tmp_path)
    if capi.lib().psulvsb_device_count() > 0:
        pytest.skip("GPU present: covered by the gpu-marked test")
    r = subprocess.run([exe, "300"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 3, r.stdout + r.stderr          # PSULVSB_ERR_NO_DEVICE, no CPU fallback
    assert "valid=0" in r.stdout and "no CUDA device" in r.stdout


@pytest.mark.gpu
def test_facade_registers_synthetic_pair(tmp_path):
    exe = build_snippet(tmp_path)
    r = subprocess.run([exe, "1500"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "valid=1" in r.stdout
