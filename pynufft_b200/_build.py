"""Build libb200nufft.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libb200nufft.so')
SOURCES = ['plan.cu', 'stages.cu', 'interp_generic.cu', 'col3d.cu', 'interp_tiled.cu', 'grid_tiled.cu', 'fft256.cu', 'sweep2d.cu', 'fftbi.cu', 'single2d.cu', 'solver.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-O2']
# tuning experiments: extra -D definitions (part of the source hash, so a change rebuilds the library)
NVCC_FLAGS += os.environ.get('B200NUFFT_NVCC_DEFS', '').split()


def _nvcc():
    exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(exe):
        raise RuntimeError('nvcc not found; libb200nufft.so cannot be built')
    return exe


HASHFILE = LIB + '.srchash'


def source_hash():
    """sha256 over the CUDA sources, the internal headers and the C header: what the shared library was built from.
    build() stores it beside the library; _lib.load() compares, so a stale binary is rebuilt instead of being loaded
    silently (the .so is git-ignored, so it can outlive a source change)."""
    import hashlib
    h = hashlib.sha256()
    files = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cu', '.cuh', '.h')))
    files.append(os.path.join(HERE, '..', 'include', 'b200nufft.h'))
    for f in files:
        h.update(os.path.basename(f).encode())
        with open(f, 'rb') as fh:
            h.update(fh.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def built_hash():
    try:
        with open(HASHFILE) as fh:
            return fh.read().strip()
    except OSError:
        return None


def is_current():
    return os.path.exists(LIB) and built_hash() == source_hash()


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
    # a library built from other sources / flags than the ones in the tree: the objects are of unknown origin too, unless
    # their own hash record says otherwise
    objhash = os.path.join(objdir, 'srchash')
    try:
        objs_current = open(objhash).read().strip() == source_hash()
    except OSError:
        objs_current = False
    if not objs_current and not is_current():
        force = True
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
    if jobs or force or _stale(LIB, objs) or not is_current():
        cmd = [nvcc, '-shared', '-o', LIB] + objs + ['-lcufft', '-Xlinker', '-rpath', '-Xlinker', '/usr/local/cuda/lib64']
        run(cmd)
        with open(HASHFILE, 'w') as fh:
            fh.write(source_hash() + '\n')
        with open(objhash, 'w') as fh:
            fh.write(source_hash() + '\n')
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
