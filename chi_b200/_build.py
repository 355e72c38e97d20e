"""Builds the native library: core + one translation unit per model.

``build_core`` compiles ``csrc/core.cu`` and the generated sources of every
built-in model (chi_b200/_library.py) for sm_100a into
``chi_b200/libchi_b200.so``.  ``build_model_module`` compiles one additional
model (e.g. a user's SBML file) into ``chi_b200/_modules/`` at run time; the
module registers its kernels with the core library when it is loaded.

Nothing here falls back to a CPU implementation: if nvcc is missing the build
fails loudly.
"""
import concurrent.futures
import hashlib
import json
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
GEN = os.path.join(CSRC, 'gen')
# CHI_B200_BUILD_TAG=<tag>: a side build for kernel experiments (its own
# object directory and library name; load it with CHI_B200_LIB=<path>)
_TAG = os.environ.get('CHI_B200_BUILD_TAG', '')
OBJ = os.path.join(HERE, '_obj' + ('_' + _TAG if _TAG else ''))
MODULES = os.path.join(HERE, '_modules')
LIB = os.path.join(
    HERE, 'libchi_b200%s.so' % ('_' + _TAG if _TAG else ''))

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3',
    '-std=c++17', '-Xcompiler', '-fPIC', '-cudart', 'shared',
    '--expt-relaxed-constexpr', '-Xcudafe', '--diag_suppress=177',
    '-I' + CSRC,
]
LINK_FLAGS = ['-Xlinker', '-rpath=/usr/local/cuda/lib64']
# extra compile flags for experiments, e.g. CHI_B200_NVCC_FLAGS="-DFOO=1"
NVCC_FLAGS += os.environ.get('CHI_B200_NVCC_FLAGS', '').split()


def nvcc():
    exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(exe):
        raise RuntimeError(
            'nvcc not found: chi_b200 needs the CUDA toolchain to build its '
            'kernels and has no CPU fallback.')
    return exe


def _run(cmd):
    proc = subprocess.run(cmd, stdout=subprocess.PIPE,
                          stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        raise RuntimeError(
            'Command failed: %s\n%s' % (' '.join(cmd), proc.stdout))
    return proc.stdout


def _digest(paths, extra=''):
    h = hashlib.sha1(extra.encode())
    for p in paths:
        with open(p, 'rb') as f:
            h.update(f.read())
    return h.hexdigest()


def write_model_sources(code, directory=None):
    """Writes model_<key>.cuh/.cu (only when the content changed) into
    csrc/gen (built-in models) or the given directory."""
    directory = directory or GEN
    os.makedirs(directory, exist_ok=True)
    out = []
    for name, text in (
            ('model_%s.cuh' % code.key, code.header()),
            ('model_%s.cu' % code.key, code.translation_unit())):
        path = os.path.join(directory, name)
        old = None
        if os.path.exists(path):
            with open(path) as f:
                old = f.read()
        if old != text:
            with open(path, 'w') as f:
                f.write(text)
        out.append(path)
    return out


def generate_builtin():
    """Regenerates the sources of all built-in models; returns the index."""
    from chi_b200 import _codegen, _library
    index = {}
    for name, admin, model in _library.builtin_variants():
        code = _codegen.ModelCode(model)
        write_model_sources(code)
        index[code.key] = {
            'model': name,
            'administration': None if admin is None else
            {'compartment': admin[0], 'direct': admin[1]},
            'states': code.state_names, 'constants': code.const_names,
            'outputs': code.output_names, 'linear': code.linear,
            'variants': code.variants()}
    path = os.path.join(GEN, 'index.json')
    text = json.dumps(index, indent=1, sort_keys=True) + '\n'
    old = open(path).read() if os.path.exists(path) else None
    if old != text:
        with open(path, 'w') as f:
            f.write(text)
    return index


def _compile(src, obj, deps, verbose=False):
    stamp = obj + '.sha1'
    digest = _digest([src] + deps, ' '.join(NVCC_FLAGS))
    if os.path.exists(obj) and os.path.exists(stamp) and \
            open(stamp).read() == digest:
        return ''
    flags = list(NVCC_FLAGS)
    if verbose:
        flags += ['-Xptxas', '-v']
    log = _run([nvcc()] + flags + ['-c', src, '-o', obj])
    with open(stamp, 'w') as f:
        f.write(digest)
    return log


def build_core(jobs=None, verbose=False, regenerate=True):
    """Compiles the core library with all built-in models; returns the path
    of the shared object."""
    if regenerate:
        index = generate_builtin()
    else:
        with open(os.path.join(GEN, 'index.json')) as f:
            index = json.load(f)
    os.makedirs(OBJ, exist_ok=True)
    common = [os.path.join(CSRC, 'model_api.cuh')]
    tasks = [(os.path.join(CSRC, 'core.cu'), os.path.join(OBJ, 'core.o'),
              common + [os.path.join(HERE, '..', 'include', 'chi_b200.h')])]
    for key in sorted(index):
        tasks.append((
            os.path.join(GEN, 'model_%s.cu' % key),
            os.path.join(OBJ, 'model_%s.o' % key),
            common + [os.path.join(CSRC, 'integrator.cuh'),
                      os.path.join(CSRC, 'integrator_stream.cuh'),
                      os.path.join(GEN, 'model_%s.cuh' % key)]))
    jobs = jobs or min(len(tasks), os.cpu_count() or 4)
    logs = {}
    with concurrent.futures.ThreadPoolExecutor(jobs) as pool:
        futs = {pool.submit(_compile, s, o, d, verbose): s
                for s, o, d in tasks}
        for fut in concurrent.futures.as_completed(futs):
            logs[futs[fut]] = fut.result()
    objs = [o for _, o, _ in tasks]
    stamp = LIB + '.sha1'
    digest = _digest(objs)
    if not (os.path.exists(LIB) and os.path.exists(stamp)
            and open(stamp).read() == digest):
        _run([nvcc(), '-shared', '-cudart', 'shared', '-o', LIB] + objs
             + LINK_FLAGS)
        with open(stamp, 'w') as f:
            f.write(digest)
    return LIB, logs


def build_model_module(code):
    """Compiles one extra model into chi_b200/_modules/ and returns the
    path of its shared object (which registers itself on load)."""
    os.makedirs(MODULES, exist_ok=True)
    os.makedirs(OBJ, exist_ok=True)
    cuh, cu = write_model_sources(code, MODULES)
    deps = [cuh, os.path.join(CSRC, 'model_api.cuh'),
            os.path.join(CSRC, 'integrator.cuh'),
            os.path.join(CSRC, 'integrator_stream.cuh')]
    # the name carries a digest of the generated code and the integrator
    # headers: a module built against older headers is never picked up
    tag = _digest([cu] + deps, ' '.join(NVCC_FLAGS))[:10]
    so = os.path.join(MODULES, 'libchi_b200_model_%s_%s.so' % (code.key, tag))
    if os.path.exists(so):
        return so
    obj = os.path.join(OBJ, 'model_%s_%s.o' % (code.key, tag))
    _compile(cu, obj, deps)
    _run([nvcc(), '-shared', '-cudart', 'shared', '-o', so, obj,
          '-L' + HERE, '-lchi_b200', '-Xlinker', '-rpath=' + HERE]
         + LINK_FLAGS)
    return so


if __name__ == '__main__':
    import sys
    import time
    t0 = time.time()
    lib, logs = build_core(verbose='-v' in sys.argv)
    for src, log in sorted(logs.items()):
        if log.strip():
            print('==', os.path.basename(src))
            print(log)
    print('built', lib, 'in %.1f s' % (time.time() - t0))
