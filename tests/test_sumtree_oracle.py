"""The plain-C sum-tree oracle against its numpy twin and against structural invariants
(no reference implementation exists: parity unpinned, see oracle/sumtree.c)."""
import numpy as np

from oracle import sumtree_oracle as so


def test_c_oracle_matches_numpy_twin():
    rs = np.random.RandomState(0)
    cap = 64
    t = so.SumTree(cap)
    twin = np.zeros(2 * cap, np.float32)
    for _ in range(5):
        idx = rs.randint(0, cap, size=20)
        pr = rs.uniform(0.01, 3.0, size=20).astype(np.float32)
        t.update(idx, pr)
        so.np_update(twin, cap, idx, pr)
    assert np.array_equal(t.tree.view(np.uint32), twin.view(np.uint32))
    u = t.stratified_u(rs.uniform(size=16).astype(np.float32))
    i1, p1 = t.sample(u)
    i2, p2 = so.np_sample(twin, cap, u)
    assert np.array_equal(i1, i2) and np.array_equal(p1, p2)


def test_invariants_1m_leaves():
    rs = np.random.RandomState(1)
    cap = 1 << 20
    t = so.SumTree(cap)
    t.tree[cap:] = rs.uniform(0.0, 1.0, size=cap).astype(np.float32)
    t.rebuild()
    tr = t.tree
    nodes = np.arange(1, cap)
    assert np.array_equal(tr[nodes], (tr[2 * nodes] + tr[2 * nodes + 1]).astype(np.float32))
    u = t.stratified_u(rs.uniform(size=512).astype(np.float32))
    assert np.all(np.diff(u) > 0) and u[-1] < tr[1]
    idx, pr = t.sample(u)
    assert np.all(np.diff(idx) >= 0)                       # stratified masses -> sorted leaves
    assert np.array_equal(pr, tr[cap + idx])
    # duplicates: the highest j wins
    t.update(np.array([5, 5, 5]), np.array([1.0, 2.0, 3.0], np.float32))
    assert t.tree[cap + 5] == np.float32(3.0)
