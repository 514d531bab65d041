"""Build libb200nufft.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libb200nufft.so')
SOURCES = ['plan.cu', 'stages.cu', 'interp_generic.cu', 'col3d.cu', 'interp_tiled.cu', 'grid_tiled.cu', 'fft256.cu', 'batch2d.cu', 'single2d.cu', 'solver.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-O2']


def _nvcc():
    exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(exe):
        raise RuntimeError('nvcc not found; libb200nufft.so cannot be built')
    return exe


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    nvcc = _nvcc()
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    headers.append(os.path.join(HERE, '..', 'include', 'b200nufft.h'))
    objdir = os.path.join(HERE, 'build')
    os.makedirs(objdir, exist_ok=True)
    jobs = []
    objs = []
    for s in srcs:
        src = os.path.join(CSRC, s)
        obj = os.path.join(objdir, s.replace('.cu', '.o'))
        objs.append(obj)
        if force or _stale(obj, [src] + headers):
            jobs.append([nvcc] + NVCC_FLAGS + ['-Xptxas', '-v', '-c', src, '-o', obj])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed: %s\n%s\n%s' % (' '.join(cmd), r.stdout, r.stderr))
        return r.stderr

    with concurrent.futures.ThreadPoolExecutor(max_workers=8) as ex:
        logs = list(ex.map(run, jobs))
    if verbose:
        for l in logs:
            sys.stderr.write(l)
    if jobs or force or _stale(LIB, objs):
        cmd = [nvcc, '-shared', '-o', LIB] + objs + ['-lcufft', '-Xlinker', '-rpath', '-Xlinker', '/usr/local/cuda/lib64']
        run(cmd)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
