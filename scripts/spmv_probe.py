"""Scratch: fine-level kernel timings only (SpMV / Jacobi / residual) for tuning sweeps."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hypredrive_b200 import hdk
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
kind = int(sys.argv[2]) if len(sys.argv) > 2 else 7
hdk.init()
c = (1.0, 1.0, 1.0) if kind == 7 else ((1.0, 1.0, 0.01) if kind == 27 else (1e-3, 1.0, 0.1))
A, b = hdk.DCsr.stencil(kind, n, n, n, c=c)
out = []
for k, name in ((0, "spmv"), (1, "jacobi"), (2, "residual")):
    ms, by = hdk.time_kernel(A, None, k, 30)
    out.append(f"{name} {ms:.4f} ms {by/ms/1e6:.0f} GB/s")
print(os.environ.get("HDK_SPMV_ROWS_MULT", "1"), os.environ.get("HDK_SPMV_IMPL", "tma"), kind, " | ".join(out), flush=True)
