"""Build libronk.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo).

    python -m ron_tensorflow_b200.build [--force] [--verbose]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libronk.so')
SOURCES = ['api.cu', 'anchors.cu', 'match_encode.cu', 'postprocess.cu', 'nms.cu', 'tpfp.cu', 'misc.cu', 'roneval.cu', 'losses.cu', 'npmethods.cu', 'voceval.cu', 'hostpath.cu']
HEADERS = ['common.cuh', 'topk.cuh', os.path.join('..', '..', 'include', 'ronk.h')]

# -fmad=false: one rounding per float op (parity with one TF op per node); IEEE div/sqrt are
# nvcc defaults (-prec-div=true -prec-sqrt=true, no --use_fast_math); -lineinfo for ncu source view.
NVCC_FLAGS = ['-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
              '-fmad=false', '-Xcompiler', '-fPIC', '-shared']


def _nvcc():
    for c in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return 'nvcc'


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + \
        ['-o', LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        print(' '.join(cmd))
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + r.stdout)
    if verbose:
        print(r.stdout)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
