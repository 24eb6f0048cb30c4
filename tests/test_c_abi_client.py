"""The C ABI from a plain C program: tests/c_client/abi_client.c is compiled with gcc against include/scp_b200.h and
linked with libscp_b200.so + the CUDA runtime -- no Python, no torch on that side of the boundary.

  * CPU (not gpu): the client compiles and links without warnings (header is valid C11, every symbol it uses resolves)
    and, with no device present, exits with its "nothing to run" code instead of computing anything on the host.
  * GPU: the client runs S1 (weighted sum fwd+bwd), S2 (table preparation + VQ forward) and S3 (InfoNCE fwd+bwd) on the
    device, checks each against a scalar fp64 restatement of the reference arithmetic, and exercises the status-code
    contract (SCP_ERR_INVALID / _UNSUPPORTED / _WORKSPACE) with bad arguments.
"""
import os
import shutil
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "speechclip_plus_b200")
CUDA_HOME = os.environ.get("CUDA_HOME", "/usr/local/cuda")
EXIT_NO_DEVICE = 77


def _build_client(tmp_path) -> str:
    from speechclip_plus_b200 import build as scp_build
    lib = scp_build.build()          # no-op when the in-tree library is up to date
    assert os.path.exists(lib)
    gcc = shutil.which("gcc")
    assert gcc, "gcc not found"
    exe = str(tmp_path / "abi_client")
    cmd = [gcc, "-std=c11", "-O1", "-Wall", "-Wextra", "-Werror", f"-I{ROOT}/include", f"-I{CUDA_HOME}/include",
           f"{ROOT}/tests/c_client/abi_client.c", f"-L{PKG}", "-lscp_b200", f"-L{CUDA_HOME}/lib64", "-lcudart", "-lm",
           f"-Wl,-rpath,{PKG}", f"-Wl,-rpath,{CUDA_HOME}/lib64", "-o", exe]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


@pytest.mark.skipif(torch.cuda.is_available(), reason="device present: the gpu test runs the client")
def test_c_client_compiles_links_and_refuses_to_run_without_a_device(tmp_path):
    exe = _build_client(tmp_path)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert res.returncode == EXIT_NO_DEVICE, (res.returncode, res.stdout, res.stderr)
    assert "libscp_b200 version" in res.stdout


@pytest.mark.gpu
def test_c_client_runs_the_hot_path_on_the_device(tmp_path):
    exe = _build_client(tmp_path)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, (res.returncode, res.stdout[-2000:], res.stderr[-4000:])
    assert "0 failure(s)" in res.stdout
