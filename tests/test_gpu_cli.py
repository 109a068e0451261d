"""GPU test of the file-driven path: hypredrive-cli on an ex1.yml-style configuration whose IJ
files (ASCII and binary, hypre / hypredrive on-disk formats) are regenerated here -- the
reference's data/ps3d10pt7 is not in its tree (data/README.md).  Output is compared with the
layout of examples/refOutput/ex1.txt and the iteration count with the oracle."""
import os
import re
import struct
import subprocess

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "hypredrive_b200", "bin", "hypredrive-cli")


def _write_ascii(dirname, A, b):
    n = A.shape[0]
    with open(os.path.join(dirname, "IJ.out.A.00000"), "w") as f:
        f.write(f"0 {n - 1} 0 {n - 1}\n")
        for i in range(n):
            for k in range(A.indptr[i], A.indptr[i + 1]):
                f.write(f"{i} {A.indices[k]} {A.data[k]:.16e}\n")
    with open(os.path.join(dirname, "IJ.out.b.00000"), "w") as f:
        f.write(f"0 {n - 1}\n")
        for i in range(n):
            f.write(f"{i} {b[i]:.16e}\n")


def _write_binary(dirname, A, b):
    n = A.shape[0]
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(A.indptr))
    hdr = [0] * 11
    hdr[1], hdr[2], hdr[5], hdr[6], hdr[7], hdr[8] = 8, 8, n, A.nnz, 0, n - 1
    with open(os.path.join(dirname, "IJ.bin.A.00000.bin"), "wb") as f:
        f.write(struct.pack("<11Q", *hdr))
        f.write(rows.tobytes()); f.write(A.indices.astype(np.int64).tobytes()); f.write(A.data.astype(np.float64).tobytes())
    vh = [0] * 8
    vh[1], vh[5] = 8, n
    with open(os.path.join(dirname, "IJ.bin.b.00000.bin"), "wb") as f:
        f.write(struct.pack("<8Q", *vh))
        f.write(b.astype(np.float64).tobytes())


@pytest.mark.parametrize("fmt", ["ascii", "binary"])
def test_cli_ex1_style_run(gpu, tmp_path, fmt):
    assert os.path.exists(CLI), "hypredrive-cli was not built"
    A, _ = O.gen("lap7", 10, 10, 10, diag_first=False)
    b = np.ones(1000)
    d = tmp_path / "ps3d10pt7" / "np1"
    d.mkdir(parents=True)
    if fmt == "ascii":
        _write_ascii(str(d), A, b)
        names = ("IJ.out.A", "IJ.out.b")
    else:
        _write_binary(str(d), A, b)
        names = ("IJ.bin.A", "IJ.bin.b")
    yml = tmp_path / "ex1.yml"
    yml.write_text(f"general:\n  use_millisec: on # Turn this off for reporting times in [s]\n  dev_pool_size: 0.01\n\n"
                   f"linear_system:\n  rhs_filename: {d}/{names[1]}\n  matrix_filename: {d}/{names[0]}\n\n"
                   "solver: pcg\n\npreconditioner: amg\n")
    r = subprocess.run([CLI, str(yml)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    out = r.stdout
    assert "Running on 1 MPI rank" in out
    assert "Solving linear system #0 with 1000 rows and 6400 nonzeros..." in out     # refOutput/ex1.txt:17
    assert "solver: pcg" in out and "preconditioner: amg" in out                      # configuration echo
    assert "|  Entry |  times [ms] |  times [ms] |  times [ms] |  res. norm |  res. norm |  iters |" in out
    row = [l for l in out.splitlines() if l.startswith("|      0 |")][0]
    cells = [c.strip() for c in row.strip("|").split("|")]
    Ad, _ = O.gen("lap7", 10, 10, 10)
    H = O.Hierarchy(Ad, O.default_params(True))        # this library = the reference's GPU-build defaults
    x, info = O.pcg(Ad, b, M=H, rel_tol=1e-6)
    assert cells[4] == "3.16e+01"                                                     # ||b|| = sqrt(1000)
    assert int(cells[6]) == info["iters"]
    assert abs(float(cells[5]) - np.linalg.norm(b - Ad @ x) / np.linalg.norm(b)) <= 5e-3 * float(cells[5]) + 1e-12
    assert re.search(r"hypredrive-cli done!", out)


def test_cli_overrides_and_missing_file(gpu, tmp_path):
    yml = tmp_path / "bad.yml"
    yml.write_text("linear_system:\n  matrix_filename: nowhere/IJ.out.A\n  rhs_filename: nowhere/IJ.out.b\nsolver: pcg\npreconditioner: amg\n")
    r = subprocess.run([CLI, str(yml)], capture_output=True, text=True, timeout=120)
    assert r.returncode != 0 and "HYPREDRIVE Failure!!!" in r.stderr and "not found" in r.stderr


def _write_parts(dirname, A, b, cuts, binary):
    """The same system written as len(cuts)-1 parts (an np>1 dump), to be read by ONE rank."""
    for p in range(len(cuts) - 1):
        r0, r1 = cuts[p], cuts[p + 1]
        k0, k1 = A.indptr[r0], A.indptr[r1]
        rows = np.repeat(np.arange(r0, r1, dtype=np.int64), np.diff(A.indptr[r0:r1 + 1]))
        if binary:
            hdr = [0] * 11
            hdr[1], hdr[2], hdr[5], hdr[6], hdr[7], hdr[8] = 8, 8, r1 - r0, k1 - k0, r0, r1 - 1
            with open(os.path.join(dirname, f"IJ.A.{p:05d}.bin"), "wb") as f:
                f.write(struct.pack("<11Q", *hdr))
                f.write(rows.tobytes()); f.write(A.indices[k0:k1].astype(np.int64).tobytes())
                f.write(A.data[k0:k1].astype(np.float64).tobytes())
            vh = [0] * 8
            vh[1], vh[5] = 8, r1 - r0
            with open(os.path.join(dirname, f"IJ.b.{p:05d}.bin"), "wb") as f:
                f.write(struct.pack("<8Q", *vh))
                f.write(b[r0:r1].astype(np.float64).tobytes())
        else:
            n = A.shape[0]
            with open(os.path.join(dirname, f"IJ.A.{p:05d}"), "w") as f:
                f.write(f"{r0} {r1 - 1} 0 {n - 1}\n")
                for k, i in zip(range(k0, k1), rows):
                    f.write(f"{i} {A.indices[k]} {A.data[k]:.16e}\n")
            with open(os.path.join(dirname, f"IJ.b.{p:05d}"), "w") as f:
                f.write(f"{r0} {r1 - 1}\n")
                for i in range(r0, r1):
                    f.write(f"{i} {b[i]:.16e}\n")


@pytest.mark.parametrize("binary", [False, True])
def test_cli_reads_more_parts_than_ranks(gpu, tmp_path, binary):
    """A data set dumped with 3 parts, solved on 1 rank: the rank concatenates its share of
    consecutive parts (reference src/internal/matrix.c:184-235) -- nothing is silently dropped."""
    A, _ = O.gen("lap7", 10, 10, 10, diag_first=False)
    b = np.arange(1000, dtype=np.float64) % 7 + 1.0
    _write_parts(str(tmp_path), A, b, [0, 300, 640, 1000], binary)
    yml = tmp_path / "parts.yml"
    yml.write_text(f"linear_system:\n  rhs_filename: {tmp_path}/IJ.b\n  matrix_filename: {tmp_path}/IJ.A\n\n"
                   "solver: pcg\n\npreconditioner: amg\n")
    r = subprocess.run([CLI, str(yml)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Solving linear system #0 with 1000 rows and 6400 nonzeros..." in r.stdout
    row = [l for l in r.stdout.splitlines() if l.startswith("|      0 |")][0]
    cells = [c.strip() for c in row.strip("|").split("|")]
    assert cells[4] == f"{np.linalg.norm(b):.2e}"          # the whole right-hand side arrived
    assert float(cells[5]) < 1e-6


def test_cli_rejects_corrupt_part_header(gpu, tmp_path):
    """A binary header that announces more entries than the file holds is refused before any allocation."""
    hdr = [0] * 11
    hdr[1], hdr[2], hdr[5], hdr[6], hdr[7], hdr[8] = 8, 8, 10, 2 ** 40, 0, 9
    with open(tmp_path / "IJ.A.00000.bin", "wb") as f:
        f.write(struct.pack("<11Q", *hdr))
    vh = [0] * 8
    vh[1], vh[5] = 8, 10
    with open(tmp_path / "IJ.b.00000.bin", "wb") as f:
        f.write(struct.pack("<8Q", *vh)); f.write(np.ones(10).tobytes())
    yml = tmp_path / "bad.yml"
    yml.write_text(f"linear_system:\n  rhs_filename: {tmp_path}/IJ.b\n  matrix_filename: {tmp_path}/IJ.A\n\nsolver: pcg\npreconditioner: amg\n")
    r = subprocess.run([CLI, str(yml)], capture_output=True, text=True, timeout=120)
    assert r.returncode != 0 and "could not parse matrix file" in r.stderr
