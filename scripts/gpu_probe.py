"""Scratch perf probe (not the bench): device-assembled stencil, AMG setup, PCG, kernel timings."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hypredrive_b200 import hdk

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
kind = int(sys.argv[2]) if len(sys.argv) > 2 else 7
hdk.init()
t0 = time.time()
c = (1.0, 1.0, 1.0) if kind == 7 else ((1.0, 1.0, 0.01) if kind == 27 else (1e-3, 1.0, 0.1))
A, b = hdk.DCsr.stencil(kind, n, n, n, c=c)
hdk.sync(); print("assemble s", time.time() - t0, A.info(), A.spmv_kind(), flush=True)
for k, name in ((0, "spmv"), (1, "jacobi"), (2, "residual"), (3, "pcg_xr")):
    ms, by = hdk.time_kernel(A, None, k, 20)
    print(f"{name}: {ms:.4f} ms  {by/ms/1e6:.1f} GB/s", flush=True)
for rep in range(2):
    t0 = time.time()
    M = hdk.DAmg(A)
    hdk.sync(); ts = time.time() - t0
    print("setup s", ts, "levels", M.sizes(), "opc", M.operator_complexity(), flush=True)
    if rep == 0: M.free()
ms, by = hdk.time_kernel(A, M, 4, 10)
print(f"vcycle: {ms:.4f} ms  {by/ms/1e6:.1f} GB/s  bytes {by/1e9:.3f} GB", flush=True)
x = hdk.DVec(A.info()["local_rows"])
for rep in range(3):
    x.fill(0.0)
    hdk.launch_count_reset()
    if kind == 107:
        info = hdk.gmres(A, b, x, M, rel_tol=1e-8, max_iter=100)
    else:
        info = hdk.pcg(A, b, x, M, rel_tol=1e-6)
    print("solve", info, "launches", hdk.launch_count_reset(), "DOF*it/s %.3e" % (n**3 * info["iters"] / (info["solve_ms"] * 1e-3)), flush=True)
