#!/usr/bin/env python
"""bench.py -- headline benchmark of hypredrive_b200.

Metric (BASELINE.json): AMG-PCG solve DOF*iterations/s on the synthetic 3-D 7-point Poisson
system, 256^3 unknowns per GPU, BoomerAMG(PMIS, extended+i, l1-Jacobi)-PCG, fp64, tol 1e-6.

A "step" is one Krylov solve (the reference's "solve" timer region, src/internal/solver.c:668-683
of the reference) on an already set-up hierarchy.
  value : N_global * iterations / solve_seconds, operands resident in HBM, CUDA-event timed on
          the library's compute stream, max over ranks.
  e2e   : the same metric through the HYPREDRV_* C-ABI with HOST buffers: every step uploads the
          right-hand side from pinned host memory (SetRHSFromArray), solves (LinearSolverApply,
          including the reference's untimed r0 / final-residual evaluations) and downloads the
          solution (GetSolutionValues).
  roofline     : the fine-level fused residual SpMV (the kernel with the largest share of a
                 solve), algorithmic bytes / CUDA-event time, against MEASURED_PEAKS.json.
  cpu_baseline : the CPU oracle (restated reference, OpenMP) on a bounded sample, rank 0, N=1.

`--impl reference` times the restated reference (oracle/) on the host cores; the real
hypredrive+hypre cannot be built here (no hypre sources, no MPI, no network).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NORTH_STAR_YAML = """general:
  statistics: off
solver:
  pcg:
    max_iter: 100
    relative_tol: 1.0e-6
preconditioner:
  amg:
    coarsening:
      type: pmis
      strong_th: 0.25
    interpolation:
      prolongation_type: extended+i
      max_nnz_row: 4
    relaxation:
      down_type: l1-jacobi
      up_type: l1-jacobi
      coarse_type: ge
      num_sweeps: 1
"""

# dram__bytes_read.sum + dram__bytes_write.sum per launch of the fine-level residual SpMV at the bench
# workload (256^3 7-point rows per GPU), from the `ncu --set full` capture summarised under
# profiles/ (see profiles/README.md); None when no capture exists for the selected kernel
NCU_TRAFFIC_BYTES = {"k_spmv_sell": 1.7437e9 + 0.1171e9, "k_spmv_tma": 1.7404e9 + 0.1303e9}
CPU_SAMPLE_EDGE = int(os.environ.get("HDK_BENCH_CPU_EDGE", "160"))  # cube edge of the bounded CPU sample (about 10-30 s of work on ~8 cores)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return json.load(fh).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=self.tmp, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.tmp.flush()
        self.tmp.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.tmp.read().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.tmp.name)
        except OSError:
            pass
        if sm:
            sm.sort()
            out = {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        return out


def run_cpu_oracle(edge, steps, warmup):
    """Restated reference on the host cores: returns (value, seconds per solve, iters, setup_s)."""
    from oracle import oracle as O
    A, b = O.gen("lap7", edge, edge, edge)
    t0 = time.time()
    H = O.Hierarchy(A, O.default_params(True))
    setup_s = time.time() - t0
    times, iters = [], 0
    for s in range(warmup + steps):
        t0 = time.time()
        _, info = O.pcg(A, b, M=H, rel_tol=1e-6, max_iter=100)
        dt = time.time() - t0
        iters = info["iters"]
        if s >= warmup:
            times.append(dt)
    per = sum(times) / max(len(times), 1)
    return edge ** 3 * iters / per, per, iters, setup_s


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    steps = max(1, min(args.steps, 5))
    warmup = max(0, min(args.warmup, 1))
    val, per, iters, setup_s = run_cpu_oracle(CPU_SAMPLE_EDGE, steps, warmup)
    sample = f"7pt Poisson {CPU_SAMPLE_EDGE}^3 (same solver config), {steps} solves after {warmup} warm-up"
    line = {
        "impl": "reference", "metric": "AMG-PCG solve DOF-iters/s, 7pt Poisson", "value": val, "unit": "DOF*iters/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": per * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "synthetic 3D 7-point Poisson 256^3 fp64, BoomerAMG(PMIS, ext+i, l1-Jacobi)-PCG",
                   "note": "restated reference (oracle/, C + OpenMP) on the host cores; hypredrive+hypre itself "
                           "cannot be built in this image (hypre not vendored, no MPI, no network)"},
        "cpu_baseline": {"value": val, "unit": "DOF*iters/s", "cores": cores, "kind": "port", "sample": sample,
                         "iterations": iters, "setup_s": setup_s, "solve_s": per},
        "e2e": {"value": val, "unit": "DOF*iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


def ours(args):
    import numpy as np
    from hypredrive_b200 import hdk, driver

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    if hdk.device_count() <= 0:
        raise SystemExit("bench.py: no CUDA device visible and there is no CPU fallback "
                         "(use --impl reference for the host-core baseline)")
    hdk.init(local)
    if world > 1:
        import torch
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            buf = (C.c_ubyte * 128)()
            hdk.check(hdk.lib().hdk_comm_unique_id(buf))
            uid = torch.tensor(list(buf), dtype=torch.uint8, device="cuda")
        dist.broadcast(uid, 0)
        raw = bytes(uid.cpu().tolist())
        hdk.check(hdk.lib().hdk_comm_init(rank, world, raw))

    edge = args.n
    nz_total = edge * world                      # weak scaling: one edge^3 slab per GPU (z-slabs)
    n_loc = edge ** 3
    n_glob = n_loc * world
    row_start, row_end = rank * n_loc, (rank + 1) * n_loc - 1

    drv = driver.HypreDrive(options=NORTH_STAR_YAML)
    L = driver.api()
    t0 = time.time()
    drv.set_stencil(7, edge, edge, nz_total, (1.0, 1.0, 1.0), row_start, row_end)   # assembled in HBM
    hdk.sync()
    build_s = time.time() - t0
    driver._check(L.HYPREDRV_LinearSystemSetInitialGuess(drv._h, None), "SetInitialGuess")
    driver._check(L.HYPREDRV_LinearSolverCreate(drv._h), "LinearSolverCreate")
    driver._check(L.HYPREDRV_LinearSolverSetup(drv._h), "LinearSolverSetup")      # warm-up setup (pool growth)
    driver._check(L.HYPREDRV_LinearSolverSetup(drv._h), "LinearSolverSetup")
    d = C.c_double()
    driver._check(L.HYPREDRV_LinearSolverGetSetupTime(drv._h, C.byref(d)), "GetSetupTime")
    setup_s = d.value

    def barrier():
        hdk.sync()
        if dist is not None:
            dist.barrier()
        hdk.sync()

    def one_solve():
        driver._check(L.HYPREDRV_LinearSystemResetInitialGuess(drv._h), "ResetInitialGuess")
        driver._check(L.HYPREDRV_LinearSolverApply(drv._h), "LinearSolverApply")
        it = C.c_int()
        driver._check(L.HYPREDRV_LinearSolverGetNumIter(drv._h, C.byref(it)), "GetNumIter")
        driver._check(L.HYPREDRV_LinearSolverGetSolveTime(drv._h, C.byref(d)), "GetSolveTime")
        return it.value, d.value

    # ---- device-resident leg -----------------------------------------------------------
    for _ in range(args.warmup):
        one_solve()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    prof = os.environ.get("HDK_PROFILE_RANGE") == "1"     # ncu --profile-from-start off
    if prof:
        hdk.lib().hdk_profiler_range(1)
    hdk.launch_count_reset()
    wall0 = time.time()
    solve_s, iters = 0.0, 0
    for _ in range(args.steps):
        it, s = one_solve()
        iters = it
        solve_s += s
    barrier()
    wall = time.time() - wall0
    if prof:
        hdk.lib().hdk_profiler_range(0)
    launches = hdk.launch_count_reset()
    if dist is not None:
        import torch
        t = torch.tensor([solve_s, wall], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        solve_s, wall = float(t[0]), float(t[1])
    per_step = solve_s / args.steps
    value = n_glob * iters / per_step

    # ---- end-to-end leg: host buffers through the C-ABI ---------------------------------
    import torch
    b_host = torch.empty(n_loc, dtype=torch.float64).pin_memory().numpy()
    pv = C.POINTER(C.c_double)()
    # fetch the generated right-hand side once so each step re-uploads the same data
    rhs_p = C.POINTER(C.c_double)()
    hdk.lib().HYPREDRV_LinearSystemGetRHSValues.argtypes = [C.c_void_p, C.POINTER(C.POINTER(C.c_double))]
    hdk.lib().HYPREDRV_LinearSystemGetRHSValues.restype = C.c_uint32
    driver._check(hdk.lib().HYPREDRV_LinearSystemGetRHSValues(drv._h, C.byref(rhs_p)), "GetRHSValues")
    b_host[:] = np.ctypeslib.as_array(rhs_p, shape=(n_loc,))

    def e2e_step():
        driver._check(L.HYPREDRV_LinearSystemSetRHSFromArray(drv._h, row_start, row_end, b_host.ctypes.data), "SetRHS")
        driver._check(L.HYPREDRV_LinearSystemResetInitialGuess(drv._h), "ResetInitialGuess")
        driver._check(L.HYPREDRV_LinearSolverApply(drv._h), "LinearSolverApply")
        driver._check(L.HYPREDRV_LinearSystemGetSolutionValues(drv._h, C.byref(pv)), "GetSolutionValues")
        it = C.c_int()
        driver._check(L.HYPREDRV_LinearSolverGetNumIter(drv._h, C.byref(it)), "GetNumIter")
        return it.value

    e2e_warm = min(args.warmup, 2)
    for _ in range(e2e_warm):
        e2e_step()
    barrier()
    t0 = time.time()
    for _ in range(args.steps):
        e2e_it = e2e_step()
    barrier()
    e2e_wall = time.time() - t0
    if dist is not None:
        t = torch.tensor([e2e_wall], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_wall = float(t[0])
    e2e_value = n_glob * e2e_it / (e2e_wall / args.steps)
    clocks = sampler.stop() if sampler else None

    # ---- roofline of the dominant kernel (fine-level SpMV), live -----------------------------
    hA, hM = drv.device_handles()
    peak, peak_src = measured_peaks()
    roof = None
    extra_kernels = {}
    ms, by = C.c_double(), C.c_double()
    names = {0: "spmv", 1: "l1_jacobi_fused", 2: "residual", 3: "pcg_xr_update", 4: "vcycle"}
    for kid, name in names.items():          # collective at N > 1 (halo exchange inside): every rank runs it
        hdk.check(hdk.lib().hdk_time_kernel(hA, hM, kid, 20, C.byref(ms), C.byref(by)))
        extra_kernels[name] = {"ms": ms.value, "GBps": by.value / ms.value / 1e6, "bytes": by.value}
    if rank == 0:
        k0 = extra_kernels["residual"]   # largest share of a solve (profiles/r01_launch_shares.csv)
        kk, ka, km = C.c_int(), C.c_double(), C.c_int()
        hdk.check(hdk.lib().hdk_csr_spmv_kind(hA, C.byref(kk), C.byref(ka), C.byref(km)))
        kname = {0: "k_spmv_tma", 1: "k_spmv_vector", 2: "k_spmv_sell"}.get(kk.value, "k_spmv")
        roof = {"bound": "hbm", "achieved": k0["GBps"], "peak": peak, "unit": "GB/s", "frac": k0["GBps"] / peak,
                "traffic": NCU_TRAFFIC_BYTES.get(kname), "kernel": kname + "<RESIDUAL> (fine level, r = b - A x, per GPU)",
                "peak_source": peak_src,
                "note": "peak is the measured copy bandwidth (half reads, half writes); this kernel is 94% reads",
                "algorithmic_bytes": k0["bytes"], "ms": k0["ms"]}
    if dist is not None:
        dist.barrier()

    # ---- CPU baseline on a bounded sample (rank 0, N = 1 only) -----------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        os.environ.setdefault("OMP_NUM_THREADS", str(cores))
        cval, cper, cit, cset = run_cpu_oracle(CPU_SAMPLE_EDGE, 2, 1)
        cpu = {"value": cval, "unit": "DOF*iters/s", "cores": cores, "kind": "port",
               "sample": f"7pt Poisson {CPU_SAMPLE_EDGE}^3, same solver config, oracle/ (C + OpenMP), 2 solves after 1 warm-up",
               "iterations": cit, "setup_s": cset, "solve_s": cper}

    if rank == 0:
        line = {
            "metric": "AMG-PCG solve DOF-iters/s, 7pt Poisson 256^3 per GPU", "value": value, "unit": "DOF*iters/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"synthetic 3D 7-point Poisson {edge}x{edge}x{nz_total} fp64 ({edge}^3 rows per GPU, z-slabs), "
                                   "BoomerAMG(PMIS, ext+i, max_nnz_row 4, l1-Jacobi, GE coarse)-PCG tol 1e-6, x0 = 0",
                       "inputs_vs_L2": "operator (1.4 GB/GPU) and vectors (134 MB each) exceed the 126 MB L2; no flush needed",
                       "assembly": "device (HYPREDRV_LinearSystemSetStencil)"},
            "iterations": iters, "solve_s": per_step, "setup_s": setup_s, "build_s": build_s, "wall_s_timed_region": wall,
            "e2e": {"value": e2e_value, "unit": "DOF*iters/s", "h2d_bytes_per_step": 8 * n_loc * world,
                    "d2h_bytes_per_step": 8 * n_loc * world, "ms_per_step": e2e_wall / args.steps * 1e3,
                    "iterations": e2e_it},
            "gpu_launches": int(launches),
            "halo_exchange": {0: "none (1 rank)", 1: "nccl send/recv",
                              2: "peer-memory stores (CUDA IPC)"}.get(int(hdk.lib().hdk_comm_halo_mode()), "?"),
            "clocks": clocks, "roofline": roof, "kernels": extra_kernels,
            "cpu_baseline": cpu,
            "published_reference": {"what": "hypre CUDA driven by hypredrive, 8xB200, lap-7 256^3 (docs figure, +-10%)",
                                    "setup_s": 0.10, "solve_s": 0.076},
        }
        print(json.dumps(line), flush=True)
    drv.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=256, help="cube edge per GPU (default 256: the BASELINE config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return reference_arm(args)
    return ours(args)


if __name__ == "__main__":
    sys.exit(main())
