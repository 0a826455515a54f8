"""Build libronk.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo).

    python -m ron_tensorflow_b200.build [--force] [--verbose]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libronk.so')
SOURCES = ['api.cu', 'anchors.cu', 'match_encode.cu', 'match_encode_grid.cu', 'postprocess.cu', 'nms.cu', 'tpfp.cu', 'ap.cu', 'misc.cu', 'roneval.cu', 'losses.cu', 'npmethods.cu', 'voceval.cu', 'hostpath.cu']
HEADERS = ['common.cuh', 'encode_common.cuh', 'topk.cuh', os.path.join('..', '..', 'include', 'ronk.h')]

# -fmad=false: one rounding per float op (parity with one TF op per node); IEEE div/sqrt are
# nvcc defaults (-prec-div=true -prec-sqrt=true, no --use_fast_math); -lineinfo for ncu source view.
NVCC_FLAGS = ['-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
              '-fmad=false', '-Xcompiler', '-fPIC', '-shared']


def _nvcc():
    for c in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return 'nvcc'


OBJ = os.path.join(HERE, 'csrc', '_obj')


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """One nvcc -c per source, in parallel (objects cached under csrc/_obj, rebuilt when the source, a header or
    this file is newer), then one link."""
    if not force and not stale():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(OBJ, exist_ok=True)
    common = [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    flags = [f for f in NVCC_FLAGS if f != '-shared']

    def one(src):
        path = os.path.join(CSRC, src)
        obj = os.path.join(OBJ, src[:-3] + '.o')
        if not force and os.path.exists(obj) and all(os.path.getmtime(obj) >= os.path.getmtime(d) for d in [path] + common):
            return obj, ''
        cmd = [_nvcc()] + flags + (['-Xptxas', '-v'] if verbose else []) + ['-c', '-o', obj, path]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed on %s:\n%s' % (src, r.stdout))
        return obj, r.stdout

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        res = list(ex.map(one, SOURCES))
    if verbose:
        print(''.join(o for _, o in res))
    cmd = [_nvcc(), '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', LIB] + [o for o, _ in res]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError('nvcc link failed:\n' + r.stdout)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
