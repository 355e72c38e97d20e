"""Memory-safety check of the integrator headers (chi_b200/csrc/*.cuh) on the
host: the CPU port is the same template code the kernels are compiled from
(shared-memory slots become a stack array), so AddressSanitizer sees an
out-of-range slot, record or snapshot index that a GPU run would silently
survive.  (compute-sanitizer is not available on the GPU pool.)"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _libasan():
    try:
        path = subprocess.run(
            ['gcc', '-print-file-name=libasan.so'], stdout=subprocess.PIPE,
            text=True, check=True).stdout.strip()
    except Exception:
        return None
    return path if os.path.isabs(path) and os.path.exists(path) else None


def test_cpu_port_is_clean_under_address_sanitizer(tmp_path):
    asan = _libasan()
    if asan is None:
        pytest.skip('libasan not available')
    from oracle.csolver import build
    build.build()          # writes port_models.inc
    lib = str(tmp_path / 'libchi_port_asan.so')
    cmd = ['g++', '-O1', '-g', '-std=c++17', '-fPIC', '-shared', '-fopenmp',
           '-fsanitize=address', '-fno-omit-frame-pointer',
           '-march=x86-64-v3', '-I/usr/local/cuda/include',
           '-I' + os.path.join(ROOT, 'chi_b200', 'csrc'),
           os.path.join(ROOT, 'oracle', 'csolver', 'port.cpp'), '-o', lib]
    subprocess.run(cmd, check=True)
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS='detect_leaks=0',
               PYTHONPATH=ROOT)
    proc = subprocess.run(
        [sys.executable,
         os.path.join(ROOT, 'tests', 'helpers', 'port_asan_run.py'), lib],
        env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
        timeout=900)
    assert proc.returncode == 0 and 'SANITIZER RUN OK' in proc.stdout, \
        proc.stdout[-4000:]
