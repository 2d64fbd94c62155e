"""ctypes loader for oracle/sumtree.c (+ a numpy twin used to cross-check the C build).
TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (no reference implementation exists, see sumtree.c)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'libsumtree_oracle.so')


def build():
    subprocess.check_call(['make', '-s', '-C', _HERE])
    return _SO


def _lib():
    src = os.path.join(_HERE, 'sumtree.c')
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        build()
    lib = ctypes.CDLL(_SO)
    return lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


class SumTree:
    def __init__(self, cap: int):
        assert cap >= 1 and (cap & (cap - 1)) == 0
        self.cap = cap
        self.tree = np.zeros(2 * cap, np.float32)
        self._lib = _lib()

    def rebuild(self):
        self._lib.st_rebuild(_p(self.tree, ctypes.c_float), ctypes.c_int64(self.cap))

    def update(self, idx, prio):
        idx = np.ascontiguousarray(idx, np.int64)
        prio = np.ascontiguousarray(prio, np.float32)
        self._lib.st_update(_p(self.tree, ctypes.c_float), ctypes.c_int64(self.cap),
                            _p(idx, ctypes.c_int64), _p(prio, ctypes.c_float), ctypes.c_int64(idx.size))

    def sample(self, u):
        u = np.ascontiguousarray(u, np.float32)
        idx = np.empty(u.size, np.int64)
        pr = np.empty(u.size, np.float32)
        self._lib.st_sample(_p(self.tree, ctypes.c_float), ctypes.c_int64(self.cap), _p(u, ctypes.c_float),
                            ctypes.c_int64(u.size), _p(idx, ctypes.c_int64), _p(pr, ctypes.c_float))
        return idx, pr

    def stratified_u(self, uniform01):
        uni = np.ascontiguousarray(uniform01, np.float32)
        out = np.empty_like(uni)
        self._lib.st_stratified_u(_p(self.tree, ctypes.c_float), _p(uni, ctypes.c_float),
                                  ctypes.c_int64(uni.size), _p(out, ctypes.c_float))
        return out


    def per_sample(self, uniform01, size, beta):
        uni = np.ascontiguousarray(uniform01, np.float32)
        idx = np.empty(uni.size, np.int64)
        pr = np.empty(uni.size, np.float32)
        w = np.empty(uni.size, np.float32)
        self._lib.st_per_sample(_p(self.tree, ctypes.c_float), ctypes.c_int64(self.cap), _p(uni, ctypes.c_float),
                                ctypes.c_int64(uni.size), ctypes.c_float(size), ctypes.c_float(beta),
                                _p(idx, ctypes.c_int64), _p(pr, ctypes.c_float), _p(w, ctypes.c_float))
        return idx, pr, w


# ---- numpy twin (pure python loops: small cases only) -------------------------------------------
def np_update(tree, cap, idx, prio):
    for j in range(len(idx)):
        node = cap + int(idx[j])
        tree[node] = np.float32(prio[j])
        node >>= 1
        while node >= 1:
            tree[node] = np.float32(tree[2 * node] + tree[2 * node + 1])
            node >>= 1


def np_sample(tree, cap, u):
    idx = np.empty(len(u), np.int64)
    pr = np.empty(len(u), np.float32)
    for j in range(len(u)):
        x = np.float32(u[j])
        node = 1
        while node < cap:
            left, right = tree[2 * node], tree[2 * node + 1]
            if x < left or not (right > 0):
                node = 2 * node
            else:
                x = np.float32(x - left)
                node = 2 * node + 1
        idx[j] = node - cap
        pr[j] = tree[node]
    return idx, pr
