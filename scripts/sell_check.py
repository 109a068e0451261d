"""Scratch: does the sliced-ELL column sort take effect?  Exact equality with the sequential
CSR-order oracle holds only when the stored order is unchanged."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, scipy.sparse as sp
from hypredrive_b200 import hdk
from oracle import oracle as O
hdk.init()
rng = np.random.default_rng(5)
n = 20000
lens = rng.integers(10, 40, size=n)
indptr = np.zeros(n + 1, dtype=np.int32); indptr[1:] = np.cumsum(lens)
cols = np.concatenate([rng.choice(n, size=l, replace=False) for l in lens]).astype(np.int32)
A = sp.csr_matrix((rng.standard_normal(indptr[-1]), cols, indptr), shape=(n, n))
x = rng.standard_normal(n)
dA = hdk.DCsr.from_scipy(A)
dx, dy = hdk.DVec(n, x), hdk.DVec(n)
dA.matvec(dx, dy)
rp, cj, va = dA.diag_arrays()
Ad = sp.csr_matrix((va, cj, rp), shape=(n, n))
ref = O.matvec(Ad, x)
print("kind", dA.spmv_kind(), "exact", np.array_equal(dy.get(), ref), "maxdiff", np.abs(dy.get() - ref).max())
