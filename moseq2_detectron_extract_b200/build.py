"""Build libmoseq_b200.so in-tree with nvcc for sm_100a (B200).  No torch involved: the library is a
plain C-ABI shared object (include/moseq_b200.h) loaded with ctypes."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libmoseq_b200.so')
SOURCES = ['api.cu', 'prep.cu', 'clean.cu', 'clean_stream.cu', 'features.cu', 'epilogue.cu', 'crop.cu', 'paste.cu', 'nms.cu', 'roi_align.cu', 'rcnn.cu', 'conv_tc.cu', 'stem_tc.cu', 'inpaint.cu', 'kalman.cu', 'bground.cu', 'roi.cu', 'pipeline.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '--fmad=false', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden', '--threads', '0']


def _nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found')


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, '..', 'include', 'moseq_b200.h')]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + ['-shared', '-o', LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, '-Xptxas')
        cmd.insert(2, '-v')
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError('nvcc failed building libmoseq_b200.so')
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
