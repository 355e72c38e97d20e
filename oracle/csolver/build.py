"""Builds oracle/csolver/libchi_port.so (g++ -O3 -fopenmp).

TEST INFRASTRUCTURE ONLY.  Compiles the host port of the per-system solve for
every built-in model listed in chi_b200/csrc/gen/index.json.
"""
import json
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
GEN = os.path.join(ROOT, 'chi_b200', 'csrc', 'gen')
LIB = os.path.join(HERE, 'libchi_port.so')


def build(force=False):
    with open(os.path.join(GEN, 'index.json')) as f:
        index = json.load(f)
    lines = []
    for key in sorted(index):
        lines.append('#include "../../chi_b200/csrc/gen/model_%s.cuh"' % key)
    lines.append('#define CHI_PORT_MODELS \\')
    for key in sorted(index):
        lines.append('  CHI_PORT_MODEL("%s", Model_%s) \\' % (key, key))
    lines.append('')
    text = '\n'.join(lines) + '\n'
    inc = os.path.join(HERE, 'port_models.inc')
    if not os.path.exists(inc) or open(inc).read() != text:
        with open(inc, 'w') as f:
            f.write(text)
    srcs = [os.path.join(HERE, 'port.cpp'), inc,
            os.path.join(ROOT, 'chi_b200', 'csrc', 'integrator.cuh'),
            os.path.join(ROOT, 'chi_b200', 'csrc', 'integrator_stream.cuh'),
            os.path.join(ROOT, 'chi_b200', 'csrc', 'model_api.cuh')]
    srcs += [os.path.join(GEN, 'model_%s.cuh' % k) for k in index]
    if not force and os.path.exists(LIB) and all(
            os.path.getmtime(LIB) >= os.path.getmtime(s) for s in srcs):
        return LIB
    cmd = ['g++', '-O3', '-std=c++17', '-fPIC', '-shared', '-fopenmp',
           '-ffp-contract=fast', '-march=x86-64-v3',
           '-I/usr/local/cuda/include',
           '-I' + os.path.join(ROOT, 'chi_b200', 'csrc'),
           os.path.join(HERE, 'port.cpp'),
           '-o', LIB]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE,
                          stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        raise RuntimeError('g++ failed:\n' + proc.stdout)
    return LIB


def build_models(headers, out_dir):
    """Host port of the solve for models that are not built in: `headers`
    maps model key -> generated header (chi_b200._build.write_model_sources);
    returns the path of the shared object written to `out_dir`."""
    lines = ['#include "%s"' % os.path.abspath(h)
             for _, h in sorted(headers.items())]
    lines.append('#define CHI_PORT_MODELS \\')
    for key in sorted(headers):
        lines.append('  CHI_PORT_MODEL("%s", Model_%s) \\' % (key, key))
    lines.append('')
    inc = os.path.join(out_dir, 'port_models_custom.inc')
    with open(inc, 'w') as f:
        f.write('\n'.join(lines) + '\n')
    lib = os.path.join(out_dir, 'libchi_port_custom.so')
    cmd = ['g++', '-O3', '-std=c++17', '-fPIC', '-shared', '-fopenmp',
           '-ffp-contract=fast', '-march=x86-64-v3',
           '-I/usr/local/cuda/include',
           '-I' + os.path.join(ROOT, 'chi_b200', 'csrc'),
           '-DCHI_PORT_MODELS_FILE="%s"' % inc,
           os.path.join(HERE, 'port.cpp'), '-o', lib]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE,
                          stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        raise RuntimeError('g++ failed:\n' + proc.stdout)
    return lib


if __name__ == '__main__':
    print(build(force=True))
