"""The C++ mirror of the reference's operator interface compiles against the C ABI; on a GPU it passes the
reference's own vertex_load known-answer test (tests/integration_test_graph.rs:238-275)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(built, tmp_path):
    exe = str(tmp_path / "mirror_smoke")
    lib_dir = os.path.dirname(built.LIB_PATH)
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [gxx, "-O1", "-std=c++17", os.path.join(ROOT, "tests", "cpp", "mirror_smoke.cpp"), "-o", exe,
           "-L" + lib_dir, "-lnet_ensembles_cuda", "-Wl,-rpath," + lib_dir]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


def test_cpp_mirror_compiles_and_fails_loudly_without_gpu(built, tmp_path):
    import torch
    exe = _build(built, tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; covered by the gpu test")
    res = subprocess.run([exe, "--no-gpu"], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "no CPU fallback" in res.stdout


@pytest.mark.gpu
def test_cpp_mirror_known_answers_on_gpu(built, tmp_path):
    exe = _build(built, tmp_path)
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 0, "exit %d %s %s" % (res.returncode, res.stdout, res.stderr)
