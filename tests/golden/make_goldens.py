"""Regenerate tests/golden/*.json.

* reference_outputs.json -- the numbers the reference tree itself pins for this path (its example
  outputs and known-answer tests), transcribed with their source file:line.  The reference
  cannot be built or imported in this image (it needs hypre and MPI), so these are the only
  reference-produced values available; tests/test_oracle_goldens.py checks the oracle against them.
* hierarchy_fixtures.json -- what the oracle (oracle/, the CPU restatement) produces for three
  small systems with the north-star options: level sizes, the C/F splitting of level 0, SHA-256
  digests of the P and coarse-operator patterns and values, PCG/GMRES iteration counts and the
  leading solution entries.  They pin the oracle against drift (CPU test) and give the CUDA
  path a committed target that does not depend on running the oracle (GPU test).

Run from the repo root:  python tests/golden/make_goldens.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = [("lap7", (12, 11, 10), (1.0, 1.0, 1.0), "pcg", 1e-6),
         ("lap27", (10, 9, 8), (1.0, 1.0, 0.01), "pcg", 1e-6),
         ("convdif", (16, 8, 8), (1e-3, 1.0, 0.1), "gmres", 1e-8)]


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def hierarchy_fixture(kind, dims, c, solver, tol):
    A, b = O.gen(kind, *dims, c=c)
    H = O.Hierarchy(A, O.default_params(True))
    levels = []
    for l in range(H.nlev):
        Al = H.A(l)
        e = {"rows": int(Al.shape[0]), "nnz_A": int(Al.nnz),
             "A_pattern": digest(Al.indptr.astype(np.int32), Al.indices.astype(np.int32)),
             "A_values": digest(Al.data.astype(np.float64))}
        if l + 1 < H.nlev:
            P = H.P(l)
            e.update({"nnz_P": int(P.nnz), "P_pattern": digest(P.indptr.astype(np.int32), P.indices.astype(np.int32)),
                      "P_values": digest(P.data.astype(np.float64)), "cf": digest(H.cf(l).astype(np.int32)),
                      "n_coarse": int((H.cf(l) > 0).sum())})
        levels.append(e)
    fn = O.pcg if solver == "pcg" else O.gmres
    x, info = fn(A, b, M=H, rel_tol=tol, max_iter=100)
    return {"kind": kind, "dims": list(dims), "c": list(c), "solver": solver, "rel_tol": tol,
            "cf_level0": H.cf(0).astype(int).tolist(), "levels": levels, "iterations": int(info["iters"]),
            "x_head": [float(v) for v in x[:8]], "x_norm": float(np.linalg.norm(x))}


def main():
    ref = {
        "ex1": {"source": "examples/refOutput/ex1.txt:17,27", "rows": 1000, "nnz": 6400, "r0": "3.16e+01",
                "iterations": 6, "rel_res": "4.98e-08", "config": "PCG tol 1e-6 + BoomerAMG CPU defaults, b = ones"},
        "laplacian": {"source": "examples/refOutput/laplacian.txt:34-38", "r0": "1.00e+01", "iterations": 5,
                      "rel_res": "6.12e-07", "config": "laplacian -n 10 10 10, PCG + AMG CPU defaults"},
        "known_answers": {"source": "tests/test_setmatrix_from_csr.c:395-421; interfaces/python/tests/test_solve_serial.py:127-145,262-277",
                          "systems": [{"diag": [3.0], "rhs": [6.0], "x": [2.0]},
                                      {"diag": [1.0, 2.0, 3.0, 4.0], "rhs": [1.0, 4.0, 9.0, 16.0], "x": [1.0, 2.0, 3.0, 4.0]},
                                      {"diag": [2.0, 4.0], "rhs": [8.0, 16.0], "x": [4.0, 4.0]},
                                      {"diag": [4.0, 8.0], "rhs": [8.0, 16.0], "x": [2.0, 2.0]}]},
    }
    with open(os.path.join(HERE, "reference_outputs.json"), "w") as f:
        json.dump(ref, f, indent=1)
    fx = [hierarchy_fixture(*c) for c in CASES]
    with open(os.path.join(HERE, "hierarchy_fixtures.json"), "w") as f:
        json.dump(fx, f, indent=None, separators=(",", ":"))
    print("wrote", len(fx), "hierarchy fixtures")


if __name__ == "__main__":
    main()
