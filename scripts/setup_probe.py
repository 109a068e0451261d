"""Scratch: one warm AMG setup bracketed by the profiler range (ncu --profile-from-start off)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hypredrive_b200 import hdk
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
kind = int(sys.argv[2]) if len(sys.argv) > 2 else 7
hdk.init()
A, b = hdk.DCsr.stencil(kind, n, n, n)
M = hdk.DAmg(A); M.free()
hdk.sync()
hdk.lib().hdk_profiler_range(1)
M = hdk.DAmg(A)
hdk.sync()
hdk.lib().hdk_profiler_range(0)
print("done", M.sizes()[:3])
