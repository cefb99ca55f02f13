"""The C ABI driven from plain C (examples/c_abi_smoke.c): no Python, no PyTorch in the process."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def _compile(tmp_path):
    exe = str(tmp_path / "c_abi_smoke")
    libdir = os.path.join(ROOT, "hand3d_b200")
    cmd = ["gcc", "-O2", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(CUDA, "include"),
           os.path.join(ROOT, "examples", "c_abi_smoke.c"), "-o", exe, "-L" + libdir, "-lhand3d_b200",
           "-L" + os.path.join(CUDA, "lib64"), "-lcudart", "-lm", "-Wl,-rpath," + libdir]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return exe


@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not available")
def test_c_client_compiles_and_fails_loudly_without_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    import __graft_entry__ as g
    g.build()
    exe = _compile(tmp_path)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert res.returncode == 77, res.stdout + res.stderr          # H3D_ENODEVICE: no CPU fallback
    assert "no CPU fallback" in res.stdout


@pytest.mark.gpu
@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not available")
def test_c_client_known_answers(tmp_path):
    exe = _compile(tmp_path)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "C ABI smoke OK" in res.stdout
