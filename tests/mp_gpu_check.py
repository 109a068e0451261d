"""Multi-GPU parity check, one process per GPU (launched by torchrun; see test_gpu_multi.py).
Every rank assembles its z-slab of the 7-point Poisson system on its GPU, the ranks solve it
together (NCCL halo exchange + allreduce) through the HYPREDRV C API, and rank 0 compares
iterations and solution with the CPU oracle run on the global problem."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from hypredrive_b200 import hdk, driver

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    hdk.init(local)
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        buf = (C.c_ubyte * 128)()
        hdk.check(hdk.lib().hdk_comm_unique_id(buf))
        uid = torch.tensor(list(buf), dtype=torch.uint8, device="cuda")
    dist.broadcast(uid, 0)
    hdk.check(hdk.lib().hdk_comm_init(rank, world, bytes(uid.cpu().tolist())))

    kind = sys.argv[1] if len(sys.argv) > 1 else "lap7"
    nx, ny, nzl = (int(v) for v in (sys.argv[2:5] if len(sys.argv) > 4 else (12, 11, 6)))
    nz = nzl * world
    code, c, solver, tol = {"lap7": (7, (1.0, 1.0, 1.0), "pcg", 1e-6), "lap27": (27, (1.0, 1.0, 0.01), "pcg", 1e-6),
                            "convdif": (107, (1e-3, 1.0, 0.1), "gmres", 1e-8)}[kind]
    n_tot = nx * ny * nz
    if os.environ.get("MPCHECK_RAGGED") == "1":
        # uneven slabs that cut through grid planes: rank r owns rows [cuts[r], cuts[r+1])
        w = np.array([1.0 + 0.7 * ((r * 37) % 5) for r in range(world)])
        cuts = np.concatenate([[0], np.floor(np.cumsum(w) / w.sum() * n_tot).astype(np.int64)])
        cuts[-1] = n_tot
    else:
        cuts = np.arange(world + 1, dtype=np.int64) * (nx * ny * nzl)
    rs, re = int(cuts[rank]), int(cuts[rank + 1]) - 1
    n_loc = re - rs + 1
    opts = {"general": {"statistics": False}, "solver": {solver: {"relative_tol": tol, "max_iter": 100}},
            "preconditioner": "amg"}
    with driver.HypreDrive(options=opts) as drv:
        drv.set_stencil(code, nx, ny, nz, c, rs, re)
        drv.solve()
        x_loc = drv.get_solution()
        iters, conv = drv.last_iterations, drv.last_converged
    xs = [torch.zeros(int(cuts[r + 1] - cuts[r]), dtype=torch.float64, device="cuda") for r in range(world)]
    for r in range(world):                       # uneven pieces: one broadcast per owner
        if r == rank:
            xs[r].copy_(torch.from_numpy(x_loc).cuda())
        dist.broadcast(xs[r], r)
    ok = True
    if rank == 0:
        from oracle import oracle as O
        x = torch.cat(xs).cpu().numpy()
        A, b = O.gen(kind, nx, ny, nz, c=c)
        H = O.Hierarchy(A, O.default_params(True))
        xr, ir = (O.pcg if solver == "pcg" else O.gmres)(A, b, M=H, rel_tol=tol, max_iter=100)
        rel = np.linalg.norm(x - xr) / np.linalg.norm(xr)
        res = np.linalg.norm(b - A @ x) / np.linalg.norm(b)
        ok = bool(conv) and abs(iters - ir["iters"]) <= 1 and rel <= 1e-8 and res < tol * 1.0001
        print(f"MPCHECK kind={kind} world={world} grid={nx}x{ny}x{nz} iters={iters} oracle_iters={ir['iters']} "
              f"rel_diff={rel:.2e} true_res={res:.2e} ok={ok}", flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
