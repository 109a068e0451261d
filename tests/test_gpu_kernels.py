"""GPU parity tests proper: every device kernel family against the CPU oracle, through the
hdk_* C-ABI of libHYPREDRV.so (ctypes).  Integer results and per-row floating-point results
are compared bit for bit; reductions (dots, Krylov scalars) within stated tolerances."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _random_csr(n, seed, max_len=40, empty_frac=0.1):
    rng = np.random.default_rng(seed)
    lens = rng.integers(0, max_len, size=n)
    lens[rng.random(n) < empty_frac] = 0
    indptr = np.zeros(n + 1, dtype=np.int64)
    indptr[1:] = np.cumsum(lens)
    cols = np.concatenate([rng.choice(n, size=l, replace=False) for l in lens] + [np.zeros(0, dtype=np.int64)])
    data = rng.standard_normal(indptr[-1])
    A = sp.csr_matrix((n, n))
    A.indptr, A.indices, A.data = indptr.astype(np.int32), cols.astype(np.int32), data
    A._shape = (n, n)
    return A


CASES = [("lap7", (12, 11, 10), (1.0, 1.0, 1.0)), ("lap7", (24, 24, 24), (1.0, 1.0, 1.0)),
         ("lap27", (10, 9, 8), (1.0, 1.0, 0.01)), ("convdif", (16, 8, 8), (1e-3, 1.0, 0.1)),
         ("lap7", (10, 10, 10), (1.0, 1.0, 1.0))]


@pytest.mark.parametrize("kind,dims,c", CASES)
def test_spmv_bit_exact(gpu, kind, dims, c):
    A, b = O.gen(kind, *dims, c=c, diag_first=False)
    dA = gpu.DCsr.from_scipy(A)
    # device assembly applies the diagonal-first swap: compare storage with the oracle's
    Ad, _ = O.gen(kind, *dims, c=c, diag_first=True)
    rp, cj, va = dA.diag_arrays()
    assert np.array_equal(rp, Ad.indptr) and np.array_equal(cj, Ad.indices) and np.array_equal(va, Ad.data)
    rng = np.random.default_rng(1)
    x = rng.standard_normal(A.shape[0])
    dx, dy = gpu.DVec(x.size, x), gpu.DVec(x.size)
    dA.matvec(dx, dy)
    ref = O.matvec(Ad, x)
    if dA.spmv_kind()["avg_row"] <= 12.0:
        assert np.array_equal(dy.get(), ref)               # one lane per row: the oracle's exact order
    else:                                                  # several lanes per row: fixed butterfly order
        assert np.allclose(dy.get(), ref, rtol=0, atol=1e-14 * np.abs(Ad).sum(axis=1).max() * np.abs(x).max())
    db, dr = gpu.DVec(x.size, b), gpu.DVec(x.size)
    dA.residual(dx, db, dr)
    assert np.allclose(dr.get(), b - Ad @ x, rtol=0, atol=1e-12 * np.abs(Ad).sum(axis=1).max() * np.abs(x).max())


def test_spmv_ragged_and_empty_rows(gpu):
    A = _random_csr(5000, 7)
    gpu.tune("sell_min_rows", 0)                           # force the sliced-ELL kernel on a small matrix
    try:
        dA = gpu.DCsr.from_scipy(A)
    finally:
        gpu.tune("sell_min_rows", 200000)
    x = np.random.default_rng(2).standard_normal(5000)
    dx, dy = gpu.DVec(5000, x), gpu.DVec(5000)
    dA.matvec(dx, dy, alpha=2.0, beta=0.0)
    ref = 2.0 * (A @ x)
    assert np.allclose(dy.get(), ref, rtol=1e-13, atol=1e-13)
    assert dA.spmv_kind()["kind"] == 2
    # user matrices keep their stored order in the slices: one lane per row = the oracle's order
    rp, cj, va = dA.diag_arrays()
    assert np.array_equal(dy.get(), 2.0 * O.matvec(sp.csr_matrix((va, cj, rp), shape=A.shape), x))
    # short ragged rows stay on the bulk-copy stream kernel; one lane per row = scipy's order
    B = _random_csr(4999, 8, max_len=12)
    dB = gpu.DCsr.from_scipy(B)
    assert dB.spmv_kind()["kind"] == 0
    xb = np.random.default_rng(3).standard_normal(4999)
    dxb, dyb, dbb = gpu.DVec(4999, xb), gpu.DVec(4999), gpu.DVec(4999, xb[::-1].copy())
    dB.matvec(dxb, dyb)
    assert np.allclose(dyb.get(), B @ xb, rtol=1e-13, atol=1e-13)
    # every fused epilogue of the sliced-ELL kernel on a row count that is not a multiple of 32
    b = np.random.default_rng(4).standard_normal(5000)
    db, dr = gpu.DVec(5000, b), gpu.DVec(5000)
    dA.residual(dx, db, dr)
    assert np.allclose(dr.get(), b - A @ x, rtol=1e-12, atol=1e-12)
    dy2 = gpu.DVec(5000, b)
    dA.matvec(dx, dy2, alpha=-0.5, beta=3.0)
    assert np.allclose(dy2.get(), -0.5 * (A @ x) + 3.0 * b, rtol=1e-12, atol=1e-12)


def test_spmv_long_rows_vector_kernel(gpu):
    n = 3000
    rng = np.random.default_rng(3)
    A = sp.random(n, n, density=0.5, format="csr", random_state=4) + sp.eye(n, format="csr")
    A = A.tocsr()
    dA = gpu.DCsr.from_scipy(A)
    assert dA.spmv_kind()["kind"] == 1 and dA.spmv_kind()["max_row"] > 1024
    x = rng.standard_normal(n)
    dx, dy = gpu.DVec(n, x), gpu.DVec(n)
    dA.matvec(dx, dy)
    assert np.allclose(dy.get(), A @ x, rtol=1e-12, atol=1e-10)


def test_device_stencil_matches_reference_generator(gpu):
    for kind, code, dims, c in (("lap7", 7, (9, 8, 7), (1.0, 2.0, 3.0)), ("lap27", 27, (7, 6, 5), (1.0, 1.0, 0.01)),
                                ("convdif", 107, (12, 6, 5), (1e-3, 1.0, 0.1))):
        Ad, b = O.gen(kind, *dims, c=c, diag_first=True)
        dA, db = gpu.DCsr.stencil(code, *dims, c=c)
        rp, cj, va = dA.diag_arrays()
        assert np.array_equal(rp, Ad.indptr) and np.array_equal(cj, Ad.indices)
        assert np.array_equal(va, Ad.data), kind
        assert np.array_equal(db.get(), b)


def test_vector_layer(gpu):
    n = 100003
    rng = np.random.default_rng(5)
    x, y = rng.standard_normal(n), rng.standard_normal(n)
    dx, dy = gpu.DVec(n, x), gpu.DVec(n, y)
    import ctypes as C
    d = C.c_double()
    gpu.check(gpu.lib().hdk_vec_dot(dx.p, dy.p, n, C.byref(d)))
    assert abs(d.value - x @ y) <= 1e-12 * np.abs(x * y).sum()
    for kind, ref in ((0, np.abs(x).sum()), (1, np.linalg.norm(x)), (2, np.abs(x).max())):
        gpu.check(gpu.lib().hdk_vec_norm(dx.p, n, kind, C.byref(d)))
        assert abs(d.value - ref) <= 1e-12 * ref
    gpu.check(gpu.lib().hdk_vec_axpy(0.5, dx.p, dy.p, n))
    assert np.array_equal(dy.get(), y + 0.5 * x)


def _compare_hierarchy(gpu, A, params_kw=None):
    params_kw = params_kw or {}
    H = O.Hierarchy(A, O.default_params(True, **params_kw))
    dA = gpu.DCsr.from_scipy(A)
    M = gpu.DAmg(dA, gpu.amg_params(**params_kw))
    assert M.nlev == H.nlev, (M.sizes(), H.sizes())
    assert M.sizes() == H.sizes()
    for l in range(H.nlev - 1):
        S = H.S(l)
        rp, cj, _ = M.matrix(l, "S")
        assert np.array_equal(rp, S.indptr) and np.array_equal(cj, S.indices), f"S level {l}"
        assert np.array_equal(M.measure(l), H.measure(l)), f"measure level {l}"
        assert np.array_equal(M.cf(l), H.cf(l)), f"C/F splitting level {l}"
        P = H.P(l)
        rp, cj, va = M.matrix(l, "P")
        assert np.array_equal(rp, P.indptr) and np.array_equal(cj, P.indices), f"P pattern level {l}"
        assert np.array_equal(va, P.data), f"P values level {l}"
        Ac = H.A(l + 1)
        rp, cj, va = M.matrix(l + 1, "A")
        assert np.array_equal(rp, Ac.indptr) and np.array_equal(cj, Ac.indices), f"A_c pattern level {l + 1}"
        assert np.array_equal(va, Ac.data), f"A_c values level {l + 1}"
    for l in range(H.nlev):
        assert np.array_equal(M.l1(l), H.l1(l)), f"l1 norms level {l}"
    return H, dA, M


@pytest.mark.parametrize("kind,dims,c", CASES)
def test_amg_setup_bit_exact(gpu, kind, dims, c):
    A, _ = O.gen(kind, *dims, c=c)
    _compare_hierarchy(gpu, A)


def test_amg_setup_other_options(gpu):
    A, _ = O.gen("lap7", 14, 13, 12)
    _compare_hierarchy(gpu, A, dict(max_nnz_row=0))            # no truncation
    _compare_hierarchy(gpu, A, dict(trunc_factor=0.2))         # hypre_BoomerAMGInterpTruncation: factor, then max 4
    _compare_hierarchy(gpu, A, dict(trunc_factor=0.35, max_nnz_row=0))
    B, _ = O.gen("lap27", 9, 8, 7, c=(1.0, 1.0, 0.01))
    _compare_hierarchy(gpu, B, dict(trunc_factor=0.1, max_nnz_row=6))
    _compare_hierarchy(gpu, A, dict(strong_th=0.5, max_coarse_size=20, max_nnz_row=6))
    _compare_hierarchy(gpu, A, dict(max_levels=2))


@pytest.mark.parametrize("kind,dims,c", CASES[:4])
def test_vcycle_and_pcg_parity(gpu, kind, dims, c):
    A, b = O.gen(kind, *dims, c=c)
    H, dA, M = _compare_hierarchy(gpu, A)
    n = A.shape[0]
    rng = np.random.default_rng(11)
    r = rng.standard_normal(n)
    dr, dz = gpu.DVec(n, r), gpu.DVec(n)
    M.apply(dr, dz)
    z_ref = H.precond(r)
    assert np.allclose(dz.get(), z_ref, rtol=1e-11, atol=1e-13 * np.abs(z_ref).max())
    # general V-cycle with a non-zero initial guess
    u0 = rng.standard_normal(n)
    du = gpu.DVec(n, u0)
    M.vcycle(dr, du)
    assert np.allclose(du.get(), H.vcycle(r, u0), rtol=1e-10, atol=1e-12 * np.abs(u0).max())
    if kind == "convdif":
        db, dx = gpu.DVec(n, b), gpu.DVec(n)
        info = gpu.gmres(dA, db, dx, M, rel_tol=1e-8, max_iter=100)
        x_ref, iref = O.gmres(A, b, M=H, rel_tol=1e-8, max_iter=100)
    else:
        db, dx = gpu.DVec(n, b), gpu.DVec(n)
        info = gpu.pcg(dA, db, dx, M, rel_tol=1e-6)
        x_ref, iref = O.pcg(A, b, M=H, rel_tol=1e-6)
    assert info["converged"] and iref["converged"]
    assert abs(info["iters"] - iref["iters"]) <= 1          # north-star: iteration count within +-1
    assert info["iters"] == iref["iters"]
    x = dx.get()
    assert np.linalg.norm(x - x_ref) <= 1e-8 * np.linalg.norm(x_ref)   # solution rel. diff <= 1e-8
    assert np.linalg.norm(b - A @ x) / np.linalg.norm(b) < (1e-8 if kind == "convdif" else 1e-6) * 1.0001


def test_cuda_path_against_committed_fixtures(gpu):
    """The CUDA path against tests/golden/hierarchy_fixtures.json without running the oracle:
    C/F splitting, P and coarse-operator patterns AND values bit for bit (SHA-256), iteration
    counts equal, leading solution entries to 1e-8."""
    import hashlib, json, os

    def digest(*arrays):
        h = hashlib.sha256()
        for a in arrays:
            h.update(np.ascontiguousarray(a).tobytes())
        return h.hexdigest()

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hierarchy_fixtures.json")
    for fx in json.load(open(path)):
        A, b = O.gen(fx["kind"], *fx["dims"], c=tuple(fx["c"]))   # input generator only
        dA = gpu.DCsr.from_scipy(A)
        M = gpu.DAmg(dA)
        assert [s[0] for s in M.sizes()] == [e["rows"] for e in fx["levels"]]
        assert np.array_equal(M.cf(0), np.array(fx["cf_level0"], dtype=np.int32))
        for l, e in enumerate(fx["levels"]):
            rp, cj, va = M.matrix(l, "A")
            assert digest(rp.astype(np.int32), cj.astype(np.int32)) == e["A_pattern"], f"A pattern level {l}"
            assert digest(va.astype(np.float64)) == e["A_values"], f"A values level {l}"
            if "nnz_P" in e:
                rp, cj, va = M.matrix(l, "P")
                assert digest(rp.astype(np.int32), cj.astype(np.int32)) == e["P_pattern"], f"P pattern level {l}"
                assert digest(va.astype(np.float64)) == e["P_values"], f"P values level {l}"
                assert digest(M.cf(l).astype(np.int32)) == e["cf"], f"C/F splitting level {l}"
        n = A.shape[0]
        db, dx = gpu.DVec(n, b), gpu.DVec(n)
        fn = gpu.pcg if fx["solver"] == "pcg" else gpu.gmres
        info = fn(dA, db, dx, M, rel_tol=fx["rel_tol"], max_iter=100)
        assert info["converged"] and info["iters"] == fx["iterations"]
        assert np.allclose(dx.get()[:8], fx["x_head"], rtol=1e-8, atol=0)


@pytest.mark.parametrize("sort", [0, 1])
def test_sliced_ell_levels_parity(gpu, sort):
    """Every operator of the hierarchy (A, P, R on all levels) on the sliced-ELL kernel: the
    hierarchy stays bit-identical to the oracle's (the slices are a copy); the V-cycle agrees
    within rounding, with the stored column order and with coarse rows sorted by column."""
    A, b = O.gen("lap7", 20, 18, 16)
    gpu.tune("sell_min_rows", 0)
    gpu.tune("sell_sort", sort)
    try:
        H, dA, M = _compare_hierarchy(gpu, A)
        assert dA.spmv_kind()["kind"] == 2
        n = A.shape[0]
        r = np.random.default_rng(12).standard_normal(n)
        dr, dz = gpu.DVec(n, r), gpu.DVec(n)
        M.apply(dr, dz)
        z_ref = H.precond(r)
        assert np.allclose(dz.get(), z_ref, rtol=1e-11, atol=1e-13 * np.abs(z_ref).max())
        db, dx = gpu.DVec(n, b), gpu.DVec(n)
        info = gpu.pcg(dA, db, dx, M, rel_tol=1e-8, max_iter=100)
        xo, io = O.pcg(A, b, M=H, rel_tol=1e-8, max_iter=100)
        assert abs(info["iters"] - io["iters"]) <= 1 and info["converged"]
        assert np.linalg.norm(dx.get() - xo) <= 1e-8 * np.linalg.norm(xo)
    finally:
        gpu.tune("sell_min_rows", 200000)
        gpu.tune("sell_sort", 1)


@pytest.mark.parametrize("down,up", [(2, 2), (1, 2), (2, 1), (0, 1), (1, 0)])
def test_sweep_counts_and_prefilled_first_sweep(gpu, down, up):
    """Buffer parity of the V-cycle for every sweep combination, including the fused paths where
    the first zero-guess sweep is produced by the restriction kernel / the PCG x-r update."""
    A, b = O.gen("lap7", 14, 12, 10)
    kw = dict(sweeps_down=down, sweeps_up=up)
    H = O.Hierarchy(A, O.default_params(True, **kw))
    dA = gpu.DCsr.from_scipy(A)
    M = gpu.DAmg(dA, gpu.amg_params(**kw))
    n = A.shape[0]
    r = np.random.default_rng(5).standard_normal(n)
    dr, dz = gpu.DVec(n, r), gpu.DVec(n)
    M.apply(dr, dz)
    z_ref = H.precond(r)
    assert np.allclose(dz.get(), z_ref, rtol=1e-11, atol=1e-13 * np.abs(z_ref).max())
    db, dx = gpu.DVec(n, b), gpu.DVec(n)
    info = gpu.pcg(dA, db, dx, M, rel_tol=1e-8, max_iter=200)
    xo, io = O.pcg(A, b, M=H, rel_tol=1e-8, max_iter=200)
    assert info["converged"] and abs(info["iters"] - io["iters"]) <= 1
    assert np.linalg.norm(dx.get() - xo) <= 1e-8 * np.linalg.norm(xo)


def test_pcg_without_preconditioner_and_zero_rhs(gpu):
    A, b = O.gen("lap7", 8, 8, 8)
    n = A.shape[0]
    dA = gpu.DCsr.from_scipy(A)
    db, dx = gpu.DVec(n, b), gpu.DVec(n)
    info = gpu.pcg(dA, db, dx, None, rel_tol=1e-8, max_iter=500)
    x_ref, iref = O.pcg(A, b, M=None, rel_tol=1e-8, max_iter=500)
    assert info["converged"] and abs(info["iters"] - iref["iters"]) <= 1
    assert np.linalg.norm(dx.get() - x_ref) <= 1e-8 * np.linalg.norm(x_ref)
    dz = gpu.DVec(n, np.zeros(n))
    dx.set(np.ones(n))
    info = gpu.pcg(dA, dz, dx, None)
    assert info["iters"] == 0 and info["converged"] and np.all(dx.get() == 0.0)


def test_two_stage_gs_smoothers(gpu):
    A, b = O.gen("lap7", 12, 12, 12)
    n = A.shape[0]
    for rt in (11, 12, 7):
        kw = dict(relax_down=rt, relax_up=rt)
        H, dA, M = _compare_hierarchy(gpu, A, kw)
        r = np.random.default_rng(rt).standard_normal(n)
        dr, dz = gpu.DVec(n, r), gpu.DVec(n)
        M.apply(dr, dz)
        z_ref = H.precond(r)
        assert np.allclose(dz.get(), z_ref, rtol=1e-10, atol=1e-12 * np.abs(z_ref).max()), rt


def test_known_answer_systems(gpu):
    # tests/test_setmatrix_from_csr.c:395-421 (3x = 6), interfaces/python/tests/test_solve_serial.py:262-277
    for diag, rhs, sol in (([3.0], [6.0], [2.0]), ([1.0, 2.0, 3.0, 4.0], [1.0, 4.0, 9.0, 16.0], [1, 2, 3, 4]),
                           ([2.0, 4.0], [8.0, 16.0], [4.0, 4.0])):
        n = len(diag)
        A = sp.diags(diag).tocsr()
        dA = gpu.DCsr.from_scipy(A)
        M = gpu.DAmg(dA)
        db, dx = gpu.DVec(n, np.array(rhs)), gpu.DVec(n)
        info = gpu.pcg(dA, db, dx, M, rel_tol=1e-8)
        assert info["converged"]
        assert np.allclose(dx.get(), sol, atol=1e-6)


@pytest.mark.parametrize("solver", ["fgmres", "bicgstab"])
def test_other_krylov_callers_match_oracle(gpu, solver):
    """FGMRES and BiCGSTAB over the same kernels (reference src/internal/solver.c:229-252)."""
    A, b = O.gen("convdif", 24, 8, 8, c=(1e-3, 1.0, 0.1))
    n = A.shape[0]
    H, dA, M = _compare_hierarchy(gpu, A)
    db, dx = gpu.DVec(n, b), gpu.DVec(n)
    info = getattr(gpu, solver)(dA, db, dx, M, rel_tol=1e-9, max_iter=100)
    xr, ir = getattr(O, solver)(A, b, M=H, rel_tol=1e-9, max_iter=100)
    assert info["converged"] and abs(info["iters"] - ir["iters"]) <= 1
    assert np.linalg.norm(dx.get() - xr) <= 1e-8 * np.linalg.norm(xr)
    assert np.linalg.norm(b - A @ dx.get()) <= 1e-9 * np.linalg.norm(b) * 1.0001


def test_random_vector_matches_hypre_stream(gpu):
    """rhs_mode random / init_guess_mode random: HYPRE_ParVectorSetRandomValues(v, 2023) on one rank
    (Park-Miller stream, x = 2 r - 1), reached by skip-ahead on the device; a slab is a slice of
    the global stream."""
    n = 100003
    dv = gpu.DVec(n)
    gpu.check(gpu.lib().hdk_vec_random(dv.p, n, 0, 2023))
    ref = O.set_random(2023, n)
    assert np.array_equal(dv.get(), ref)
    assert -1.0 <= ref.min() and ref.max() < 1.0
    ds = gpu.DVec(1000)
    gpu.check(gpu.lib().hdk_vec_random(ds.p, 1000, 5000, 2023))
    assert np.array_equal(ds.get(), ref[5000:6000])


def test_timeline_diagnostics_do_not_change_the_solve(gpu, capfd):
    """hdk_tune("timeline", 1/0) brackets a solve with CUDA events per operation and prints the
    per-(level, operation) averages to stderr; the solve itself is the same."""
    A, b = O.gen("lap7", 12, 12, 12)
    n = A.shape[0]
    dA = gpu.DCsr.from_scipy(A)
    M = gpu.DAmg(dA)
    db, dx = gpu.DVec(n, b), gpu.DVec(n)
    ref = gpu.pcg(dA, db, dx, M)
    x0 = dx.get().copy()
    gpu.tune("timeline", 1)
    dx.set(np.zeros(n))
    info = gpu.pcg(dA, db, dx, M)
    gpu.tune("timeline", 0)
    err = capfd.readouterr().err
    assert info["iters"] == ref["iters"] and np.array_equal(dx.get(), x0)
    assert "hdk timeline" in err and "pcg-spmv" in err and "residual" in err
