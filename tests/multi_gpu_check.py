"""Multi-GPU parity check (run under torchrun on N >= 2 GPUs of one box; not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py

Every rank advances its x-slab of the laser-plasma case (halo exchange + rho/J all-gather over NCCL) and, on its own
GPU, the whole domain as one slab; the two must agree BIT FOR BIT in f on the rank's columns and in every replicated 1-D
array (the oracle cannot run config 5, N-GPU == 1-GPU is the multi-GPU parity test)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import veritas_b200 as vb
    from veritas_b200 import solver as S
    from veritas_b200.parallel import slab_bounds, broadcast_unique_id

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nx, np_, steps = int(os.environ.get("VRT_CHECK_NX", 512)), int(os.environ.get("VRT_CHECK_NP", 192)), 4
    ok = True
    for graph in (False, True):
        runs = {}
        for mode in ("slab", "full"):
            run = vb.LaserPlasmaRun(nx, np_, density=0.3, device=local, slab=(rank, world) if mode == "slab" else None, graph=graph)
            if mode == "slab":
                uid = broadcast_unique_id(dist, run.L, rank, device="cuda")
                run.ctx.call("vrt_comm_init", uid, rank, world)
            run.init_device()
            run.run_fields_phase()
            for _ in range(steps):
                run.advance(run.calculate_dt())
            run.ctx.sync()
            runs[mode] = run
        x0, x1 = slab_bounds(nx, rank, world)
        for s in range(2):
            a = runs["slab"].ctx.download_f(s, 0, 1)[x0 + 2:x1 + 2]
            b = runs["full"].ctx.download_f(s, 0, 1)[x0 + 2:x1 + 2]
            same = np.array_equal(a, b)
            ok &= same
            print(f"rank {rank}: species {s} f[{x0}:{x1}] bitwise equal to the 1-GPU run: {same} (max |diff| {np.abs(a - b).max():.3e}, max |f| {np.abs(b).max():.3e})", flush=True)
        for name, get in (("charge", lambda r: r.ctx.get_1d(S.CHARGE)), ("J", lambda r: r.ctx.get_1d(S.J)), ("PHI", lambda r: r.ctx.get_1d(S.PHI)),
                          ("a_squared", lambda r: r.ctx.get_1d(S.A_SQUARED)), ("Ey", lambda r: r.ctx.download_field(S.EY, 0)),
                          ("Bz", lambda r: r.ctx.download_field(S.BZ, 0))):
            a, b = get(runs["slab"]), get(runs["full"])
            same = np.array_equal(a, b)
            ok &= same
            print(f"rank {rank}: {name} bitwise equal: {same} (max |diff| {np.abs(a - b).max():.3e}; |.|max {np.abs(b).max():.3e})", flush=True)
        n_tot = torch.tensor([sum(float(runs["slab"].ctx.download_f(s, 0, 1)[x0 + 2:x1 + 2, 2:-2].sum()) for s in range(2))], dtype=torch.float64, device="cuda")
        dist.all_reduce(n_tot)
        n_full = sum(float(runs["full"].ctx.download_f(s, 0, 1)[2:-2, 2:-2].sum()) for s in range(2))
        print(f"rank {rank}: particle number, sum over slabs {n_tot.item():.15e} vs 1-GPU {n_full:.15e}", flush=True)
        for r in runs.values():
            r.ctx.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTI_GPU_CHECK", "PASS" if flag.item() == 1 else "FAIL", f"world={world} mesh={nx}x{np_}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
