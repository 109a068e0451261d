"""Scratch: set up 256^3 7-pt, run a few V-cycles (profiling target)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hypredrive_b200 import hdk
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
hdk.init()
A, b = hdk.DCsr.stencil(7, n, n, n)
M = hdk.DAmg(A)
z = hdk.DVec(A.info()["local_rows"])
for _ in range(3):
    M.apply(b, z)
hdk.sync()
hdk.lib().hdk_profiler_range(1)
M.apply(b, z)
hdk.sync()
hdk.lib().hdk_profiler_range(0)
print("done", M.sizes()[:3])
