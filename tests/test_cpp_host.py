"""Compiles and runs the C++ tests of the host-side mirror of the reference interface."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(src, exe, tmp_path, extra=()):
    import ssme_b200 as sb
    sb.load_library()
    libdir = os.path.dirname(sb.library_path())
    cmd = ["g++", "-O1", "-std=c++17", "-Wall", "-Wextra", "-I" + os.path.join(ROOT, "include"), "-o", str(tmp_path / exe),
           os.path.join(ROOT, "tests", "cpp", src), "-L" + libdir, "-lssme_b200", "-Wl,-rpath," + libdir, *extra]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "warning" not in r.stderr, r.stderr
    return str(tmp_path / exe)


def test_host_layer_known_answers(tmp_path):
    exe = _build("test_host.cpp", "test_host", tmp_path)
    r = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert " 0 failed" in r.stdout


def test_reference_pmmh_loop_and_pack_equal_the_mirror(tmp_path):
    """The reference's own ada_pmmh_mvn::commence_sampling / update_moments_and_Ct / q_samp and param::pack, compiled
    unmodified (oracle/_ref/libssme_refhdr.so), against pmmh_multichain.hpp / parameters.hpp on identical streams."""
    from oracle import refhdr_binding as rb
    if not rb.available():
        pytest.skip("oracle/_ref/libssme_refhdr.so not built and /root/reference absent")
    rb.build()
    exe = _build("test_ref_pmmh.cpp", "test_ref_pmmh", tmp_path, extra=["-L" + rb.REF_DIR, "-lssme_refhdr", "-Wl,-rpath," + rb.REF_DIR])
    r = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert " 0 failed" in r.stdout


def test_example_program_compiles(tmp_path):
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "examples"), "-B"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    exe = os.path.join(ROOT, "examples", "ssme_example")
    usage = subprocess.run([exe], capture_output=True, text=True)
    assert usage.returncode == 2 and "FILTERS_PER_PROPOSAL" in usage.stderr  # the five positional arguments of example/main.cpp:17-37
    bad = subprocess.run([exe, "data.csv", "s", "m", "ten", "5"], capture_output=True, text=True)
    assert bad.returncode == 2


@pytest.mark.gpu
def test_example_program_runs_the_reference_example(tmp_path):
    """The reference's README example: SPY returns, 500 particles, start (1, .5, 2e-4) -- here 6 iterations x 20 filters."""
    import numpy as np
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "examples"), "-B"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    y = np.load(os.path.join(ROOT, "tests", "golden", "spy_config1.npz"))["y"]
    data = tmp_path / "returns.csv"
    np.savetxt(data, y, fmt="%.17g")
    run = subprocess.run([os.path.join(ROOT, "examples", "ssme_example"), str(data), str(tmp_path / "samples"), str(tmp_path / "messages"), "6", "20"],
                         capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, run.stdout + run.stderr
    files = sorted(p for p in os.listdir(tmp_path) if p.startswith("samples"))
    assert len(files) == 1
    draws = np.loadtxt(tmp_path / files[0], delimiter=",", ndmin=2)
    assert draws.shape == (6, 3) and np.all(np.isfinite(draws))
    assert np.all(draws[:, 2] > 0) and np.all(np.abs(draws[:, 1]) < 1)   # sigma^2 > 0, |phi| < 1 on the untransformed scale


@pytest.mark.gpu
def test_gpu_host_layer(tmp_path):
    exe = _build("test_gpu_host.cpp", "test_gpu_host", tmp_path)
    r = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert " 0 failed" in r.stdout
