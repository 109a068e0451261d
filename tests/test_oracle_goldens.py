"""CPU tests: the oracle against every golden number the reference tree holds for this path
 and against independent scipy solves."""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import hashlib
import json
import os

from oracle import oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def test_reference_outputs_fixture():
    """tests/golden/reference_outputs.json: the values the reference tree pins (file:line inside)."""
    ref = json.load(open(os.path.join(GOLDEN, "reference_outputs.json")))
    A, _ = O.gen("lap7", 10, 10, 10)
    assert A.shape[0] == ref["ex1"]["rows"] and A.nnz == ref["ex1"]["nnz"]
    H = O.Hierarchy(A, O.default_params(False))
    for key, b in (("ex1", np.ones(1000)), ("laplacian", O.gen("lap7", 10, 10, 10)[1])):
        x, info = O.pcg(A, b, M=H, rel_tol=1e-6, max_iter=100)
        assert f"{np.linalg.norm(b):.2e}" == ref[key]["r0"]
        assert info["iters"] == ref[key]["iterations"]
        assert f"{np.linalg.norm(b - A @ x) / np.linalg.norm(b):.2e}" == ref[key]["rel_res"]
    for s_ in ref["known_answers"]["systems"]:
        D = sp.diags(s_["diag"]).tocsr()
        x, info = O.pcg(D, np.array(s_["rhs"]), M=O.Hierarchy(D, O.default_params(True)), rel_tol=1e-8)
        assert info["converged"] and np.allclose(x, s_["x"], atol=1e-6)


def test_oracle_reproduces_committed_hierarchy_fixtures():
    """tests/golden/hierarchy_fixtures.json (made by tests/golden/make_goldens.py): guards the oracle
    against drift -- C/F splitting, P and coarse-operator patterns and values, iteration counts."""
    for fx in json.load(open(os.path.join(GOLDEN, "hierarchy_fixtures.json"))):
        A, b = O.gen(fx["kind"], *fx["dims"], c=tuple(fx["c"]))
        H = O.Hierarchy(A, O.default_params(True))
        assert H.nlev == len(fx["levels"])
        assert np.array_equal(H.cf(0), np.array(fx["cf_level0"]))
        for l, e in enumerate(fx["levels"]):
            Al = H.A(l)
            assert Al.shape[0] == e["rows"] and Al.nnz == e["nnz_A"]
            assert _digest(Al.indptr.astype(np.int32), Al.indices.astype(np.int32)) == e["A_pattern"]
            assert _digest(Al.data.astype(np.float64)) == e["A_values"]
            if "nnz_P" in e:
                P = H.P(l)
                assert _digest(P.indptr.astype(np.int32), P.indices.astype(np.int32)) == e["P_pattern"]
                assert _digest(P.data.astype(np.float64)) == e["P_values"]
                assert _digest(H.cf(l).astype(np.int32)) == e["cf"]
        fn = O.pcg if fx["solver"] == "pcg" else O.gmres
        x, info = fn(A, b, M=H, rel_tol=fx["rel_tol"], max_iter=100)
        assert info["iters"] == fx["iterations"]
        assert np.allclose(x[:8], fx["x_head"], rtol=1e-12, atol=0)


def test_park_miller_stream_known_answer():
    # minimal-standard generator: seed 1 -> 10000th value 1043618065 (Park & Miller 1988)
    r = O.rand_stream(1, 10000)
    assert round(r[-1] * 2147483647) == 1043618065
    r = O.rand_stream(2747, 3)
    s = 2747
    for i in range(3):
        s = (16807 * s) % 2147483647
        assert r[i] == s / 2147483647


def test_golden_ex1_cpu_defaults():
    # examples/refOutput/ex1.txt:17,27 -- 1000 rows / 6400 nnz, r0 3.16e+01, 6 iterations, 4.98e-08
    A, _ = O.gen("lap7", 10, 10, 10)
    assert A.shape == (1000, 1000) and A.nnz == 6400
    b = np.ones(1000)
    H = O.Hierarchy(A, O.default_params(False))          # CPU defaults: HMIS, ext+i, hybrid l1-GS 13/14
    x, info = O.pcg(A, b, M=H, rel_tol=1e-6, max_iter=100)
    assert f"{np.linalg.norm(b):.2e}" == "3.16e+01"
    assert info["iters"] == 6
    assert f"{np.linalg.norm(b - A @ x) / np.linalg.norm(b):.2e}" == "4.98e-08"


def test_golden_laplacian_example_cpu_defaults():
    # examples/refOutput/laplacian.txt:34-38 -- laplacian -n 10 10 10, r0 1.00e+01, 5 iterations, 6.12e-07
    A, b = O.gen("lap7", 10, 10, 10)
    H = O.Hierarchy(A, O.default_params(False))
    x, info = O.pcg(A, b, M=H, rel_tol=1e-6, max_iter=100)
    assert f"{np.linalg.norm(b):.2e}" == "1.00e+01"
    assert info["iters"] == 5
    assert f"{np.linalg.norm(b - A @ x) / np.linalg.norm(b):.2e}" == "6.12e-07"


def test_known_answer_systems():
    # tests/test_setmatrix_from_csr.c:395-421 and interfaces/python/tests/test_solve_serial.py:127-145,262-277
    for diag, rhs, sol in (([3.0], [6.0], [2.0]), ([1.0, 2.0, 3.0, 4.0], [1.0, 4.0, 9.0, 16.0], [1, 2, 3, 4]),
                           ([2.0, 4.0], [8.0, 16.0], [4.0, 4.0]), ([4.0, 8.0], [8.0, 16.0], [2.0, 2.0])):
        A = sp.diags(diag).tocsr()
        for gpu in (True, False):
            H = O.Hierarchy(A, O.default_params(gpu))
            x, info = O.pcg(A, np.array(rhs), M=H, rel_tol=1e-8)
            assert info["converged"] and np.allclose(x, sol, atol=1e-6)


def test_north_star_config_against_scipy():
    for kind, dims, c in (("lap7", (16, 16, 16), (1, 1, 1)), ("lap27", (12, 12, 12), (1, 1, 0.01))):
        A, b = O.gen(kind, *dims, c=c)
        H = O.Hierarchy(A, O.default_params(True))
        x, info = O.pcg(A, b, M=H, rel_tol=1e-10, max_iter=200)
        xs = spla.spsolve(A.tocsc(), b)
        assert info["converged"]
        assert np.linalg.norm(x - xs) <= 1e-8 * np.linalg.norm(xs)


def test_gmres_convdif_against_scipy():
    A, b = O.gen("convdif", 24, 8, 8, c=(1e-3, 1.0, 0.1))
    assert abs(A - A.T).max() > 0                         # nonsymmetric
    H = O.Hierarchy(A, O.default_params(True))
    x, info = O.gmres(A, b, M=H, rel_tol=1e-10, max_iter=100)
    xs = spla.spsolve(A.tocsc(), b)
    assert info["converged"] and np.linalg.norm(x - xs) <= 1e-8 * np.linalg.norm(xs)


def test_looser_tolerance_fewer_iterations_and_determinism():
    # interfaces/python/tests/test_solve_serial.py:83-126 on the 1-D Laplacian n = 32
    n = 32
    A = sp.diags([-1, 2, -1], [-1, 0, 1], shape=(n, n)).tocsr()
    b = np.ones(n)
    H = O.Hierarchy(A, O.default_params(True))
    x1, i1 = O.pcg(A, b, M=H, rel_tol=1e-8)
    x2, i2 = O.pcg(A, b, M=H, rel_tol=1e-8)
    _, i3 = O.pcg(A, b, M=H, rel_tol=1e-2)
    assert np.array_equal(x1, x2) and i1["iters"] == i2["iters"]
    assert i3["iters"] < i1["iters"]


def test_hierarchy_invariants():
    A, _ = O.gen("lap7", 12, 12, 12)
    H = O.Hierarchy(A, O.default_params(True))
    for l in range(H.nlev - 1):
        P, cf = H.P(l), H.cf(l)
        Ac = H.A(l + 1)
        assert P.shape == (H.A(l).shape[0], Ac.shape[0]) and (cf == 1).sum() == Ac.shape[0]
        assert np.diff(P.indptr).max() <= 4                # max_nnz_row
        G = (P.T @ H.A(l) @ P).tocsr()
        assert abs(G - Ac).max() < 1e-12                   # Galerkin product
        assert np.array_equal(Ac.indices[Ac.indptr[:-1]], np.arange(Ac.shape[0]))  # diagonal first
        Pc = P[cf == 1]
        assert np.allclose(Pc.data, 1.0) and Pc.nnz == Ac.shape[0]


def test_fgmres_bicgstab_against_scipy():
    A, b = O.gen("convdif", 24, 8, 8, c=(1e-3, 1.0, 0.1))
    H = O.Hierarchy(A, O.default_params(True))
    xs = spla.spsolve(A.tocsc(), b)
    xg, ig = O.gmres(A, b, M=H, rel_tol=1e-10)
    for fn in (O.fgmres, O.bicgstab):
        x, info = fn(A, b, M=H, rel_tol=1e-10, max_iter=100)
        assert info["converged"] and np.linalg.norm(x - xs) <= 1e-8 * np.linalg.norm(xs)
    assert O.fgmres(A, b, M=H, rel_tol=1e-10)[1]["iters"] == ig["iters"]   # fixed preconditioner: FGMRES == GMRES
