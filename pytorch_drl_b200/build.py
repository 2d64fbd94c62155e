"""Build recipe for the C-ABI library `pytorch_drl_b200/lib/libdrl_b200.so`.

Plain `nvcc` for sm_100a only (no multi-arch, no JIT cache): objects go to `build/obj`, the shared
library stays in-tree so that it travels to the GPU box with the repository snapshot.
    python -m pytorch_drl_b200.build [--force] [--verbose]
"""
import concurrent.futures
import hashlib
import json
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, 'csrc')
OBJ = os.path.join(ROOT, 'build', 'obj')
LIB_DIR = os.path.join(PKG, 'lib')
LIB = os.path.join(LIB_DIR, 'libdrl_b200.so')

ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
FLAGS = ['-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr',
         '-Xcompiler', '-fvisibility=hidden']


def _nvcc() -> str:
    for c in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', shutil.which('nvcc')):
        if c and os.path.exists(c):
            return c
    raise RuntimeError('nvcc not found: the drl_b200 CUDA library cannot be built')


def _nccl_include() -> str:
    import importlib.util
    spec = importlib.util.find_spec('nvidia')
    for base in (list(spec.submodule_search_locations) if spec else []):
        p = os.path.join(base, 'nccl', 'include')
        if os.path.exists(os.path.join(p, 'nccl.h')):
            return p
    for p in ('/usr/include', '/usr/local/cuda/include'):
        if os.path.exists(os.path.join(p, 'nccl.h')):
            return p
    raise RuntimeError('nccl.h not found')


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))


def _digest() -> str:
    h = hashlib.sha256()
    for d in (CSRC, os.path.join(ROOT, 'include')):
        for f in sorted(os.listdir(d)):
            with open(os.path.join(d, f), 'rb') as fh:
                h.update(f.encode() + fh.read())
    h.update(' '.join(ARCH + FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIB_DIR, exist_ok=True)
    stamp = os.path.join(OBJ, 'stamp.json')
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp):
        try:
            if json.load(open(stamp)).get('digest') == dig:
                return LIB
        except Exception:
            pass
    nvcc = _nvcc()
    inc = ['-I', os.path.join(ROOT, 'include'), '-I', _nccl_include()]

    def compile_one(src):
        obj = os.path.join(OBJ, src[:-3] + '.o')
        cmd = [nvcc] + ARCH + FLAGS + inc + ['-c', os.path.join(CSRC, src), '-o', obj]
        if verbose:
            print(' '.join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f'nvcc failed on {src}:\n{r.stdout}\n{r.stderr}')
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as ex:
        objs = list(ex.map(compile_one, sources()))
    cmd = [nvcc] + ARCH + ['-shared', '-o', LIB] + objs + ['-ldl']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'link failed:\n{r.stdout}\n{r.stderr}')
    json.dump({'digest': dig}, open(stamp, 'w'))
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
