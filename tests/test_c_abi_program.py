"""The boundary really is a C ABI: tests/c_abi/abi_check.c, a plain C99 program that includes
include/spamm_b200.h and links libspamm_b200.so, compiled with gcc -Wall -Werror.  Without a device
it checks the ABI version and that plan creation fails loudly; on the GPU box it evaluates a
NuclearContinuum model through spamm_lnpost_batch_host and compares with the closed form in C."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build_program(tmp_path):
    from spamm_b200 import _build
    _build.build()
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not found")
    exe = os.path.join(str(tmp_path), "abi_check")
    libdir = os.path.join(ROOT, "spamm_b200")
    cmd = [gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c_abi", "abi_check.c"), "-o", exe,
           "-L", libdir, "-lspamm_b200", "-Wl,-rpath," + libdir, "-lm"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


def test_c_program_compiles_links_and_fails_loudly_without_device(tmp_path):
    exe = _build_program(tmp_path)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "abi ok, no device" in res.stdout or "lnpost ok" in res.stdout


@pytest.mark.gpu
def test_c_program_lnposterior_matches_closed_form(tmp_path):
    exe = _build_program(tmp_path)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "lnpost ok" in res.stdout, res.stdout
    assert res.stdout.count(" ok") >= 6 and "MISMATCH" not in res.stdout
