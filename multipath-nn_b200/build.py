"""Build libmpnn_sm100.so in-tree with nvcc for sm_100a (no JIT cache)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OUT = os.path.join(HERE, 'lib', 'libmpnn_sm100.so')
OBJ = os.path.join(HERE, 'build')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-Xcompiler', '-fPIC']


def _stale(src, obj, deps):
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    return any(os.path.getmtime(d) > t for d in [src] + deps)


def build(verbose=False, force=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cuh')]
    deps.append(os.path.join(HERE, '..', 'include', 'mpnn.h'))
    jobs = []
    for s in srcs:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s[:-3] + '.o')
        if force or _stale(src, obj, deps):
            cmd = [NVCC] + FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', src, '-o', obj]
            jobs.append((s, cmd))

    def run(job):
        name, cmd = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        return name, r

    with ThreadPoolExecutor(max_workers=8) as ex:
        for name, r in ex.map(run, jobs):
            if verbose or r.returncode != 0:
                sys.stderr.write('[%s]\n%s%s' % (name, r.stdout, r.stderr))
            if r.returncode != 0:
                raise RuntimeError('nvcc failed on %s' % name)
    objs = [os.path.join(OBJ, s[:-3] + '.o') for s in srcs]
    if jobs or not os.path.exists(OUT):
        cmd = [NVCC, '-shared', '-o', OUT] + objs + ['-lcudart', '-ldl']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError('link failed')
    return OUT


if __name__ == '__main__':
    print(build(verbose='-v' in sys.argv, force='-f' in sys.argv))
