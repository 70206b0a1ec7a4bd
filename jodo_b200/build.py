"""Build libjodo_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import glob
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libjodo_b200.so')
STAMP = os.path.join(HERE, '.libjodo_b200.stamp')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-std=c++17', '-lineinfo',
         '-Xcompiler', '-fPIC', '-Xptxas', '-v'] + os.environ.get('JODO_NVCC_EXTRA', '').split()


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def _digest():
    h = hashlib.sha256()
    files = _sources() + sorted(glob.glob(os.path.join(CSRC, '*.h'))) + sorted(glob.glob(os.path.join(CSRC, '*.cuh')))
    files.append(os.path.join(os.path.dirname(HERE), 'include', 'jodo_b200.h'))
    for f in files:
        h.update(f.encode())
        with open(f, 'rb') as fh:
            h.update(fh.read())
    h.update(' '.join(FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ into one shared library.  Returns the library path."""
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read() == dig:
        return LIB
    objs = []
    procs = []
    for src in _sources():
        obj = os.path.splitext(src)[0] + '.o'
        objs.append(obj)
        cmd = [NVCC] + FLAGS + ['-c', src, '-o', obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f'== {os.path.basename(src)}\n{out}')
        if p.returncode != 0:
            sys.stderr.write('\n'.join(log))
            raise RuntimeError(f'nvcc failed on {src}')
    cmd = [NVCC, '-shared', '-o', LIB] + objs + ['-lcudart']
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError('link failed')
    with open(os.path.join(HERE, 'build.log'), 'w') as f:
        f.write('\n'.join(log))
    with open(STAMP, 'w') as f:
        f.write(dig)
    if verbose:
        print('\n'.join(log))
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
