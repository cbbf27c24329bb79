"""The header is plain C and the library is usable without Python: a gcc-built client of include/eml_b200.h."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'effective_model_learning_b200')


def _build(tmp_path):
    exe = str(tmp_path / 'c_abi_smoke')
    cmd = ['gcc', '-std=c99', '-pedantic', '-Wall', '-Werror', '-I' + os.path.join(ROOT, 'include'),
           os.path.join(ROOT, 'tests', 'c_abi_smoke.c'), '-o', exe, '-L' + PKG, '-leml_b200', '-lm',
           '-Wl,-rpath,' + PKG]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    return exe


def test_c_client_compiles_and_reports_no_device_on_a_cpu_box(tmp_path):
    from effective_model_learning_b200 import _native
    if _native.device_count() > 0:
        pytest.skip('a GPU is present (covered by the gpu test)')
    out = subprocess.run([_build(tmp_path)], capture_output=True, text=True)
    assert out.returncode == 3, (out.returncode, out.stdout, out.stderr)
    assert 'no' in out.stdout.lower()


@pytest.mark.gpu
def test_c_client_matches_closed_form_on_the_gpu(tmp_path):
    out = subprocess.run([_build(tmp_path)], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert 'c abi ok' in out.stdout
