"""Compiles and runs the C++ tests of the host-side mirror of the reference interface."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(src, exe, tmp_path):
    import ssme_b200 as sb
    sb.load_library()
    libdir = os.path.dirname(sb.library_path())
    cmd = ["g++", "-O1", "-std=c++17", "-Wall", "-Wextra", "-I" + os.path.join(ROOT, "include"), "-o", str(tmp_path / exe),
           os.path.join(ROOT, "tests", "cpp", src), "-L" + libdir, "-lssme_b200", "-Wl,-rpath," + libdir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "warning" not in r.stderr, r.stderr
    return str(tmp_path / exe)


def test_host_layer_known_answers(tmp_path):
    exe = _build("test_host.cpp", "test_host", tmp_path)
    r = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert " 0 failed" in r.stdout


def test_example_program_compiles(tmp_path):
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "examples"), "-B"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    usage = subprocess.run([os.path.join(ROOT, "examples", "ssme_example")], capture_output=True, text=True)
    assert "number of pfilters" in usage.stderr  # same usage text as example/main.cpp:22-27


@pytest.mark.gpu
def test_gpu_host_layer(tmp_path):
    exe = _build("test_gpu_host.cpp", "test_gpu_host", tmp_path)
    r = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert " 0 failed" in r.stdout
