"""Multi-GPU parity check (slab partition + NCCL halo exchange + all-reduce), one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multigpu_check.py

Every rank holds an arbitrary shard of the points, exchanges them with one all-to-all (partition.exchange_points),
solves its slab, and rank 0 compares the gathered theta with the single-process oracle:
identical Counter, max|dtheta| <= 1e-9."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    import multivartv_b200 as mv
    from multivartv_b200 import partition
    from oracle import c_oracle as co
    from tests.helpers import synth

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    J, C1 = mv.PRECOND_JACOBI, mv.PRECOND_CHEB1
    cases = [([24, 22], 3000, "rcpp", 1.0, 0, J), ([24, 22], 3000, "cpp", 3.0, 0, J), ([12, 12, 13], 4000, "rcpp", 0.7, 0, J),
             ([8, 8, 8, 9], 5000, "rcpp", 1.0, 25, J), ([64, 64], 20000, "py", 0.8, 0, J),
             ([40, 40, 40], 64000, "rcpp", 1.0, 15, J), ([24, 22], 3000, "rcpp", 1.0, 0, C1),
             ([12, 12, 13], 4000, "rcpp", 0.7, 0, C1), ([8, 8, 8, 9], 5000, "rcpp", 1.0, 25, C1),
             ([40, 40, 40], 64000, "rcpp", 1.0, 15, C1),
             # 3-D strip kernels on the peer path: fused update + Horner passes of degree 2..4, and what AUTO picks
             ([40, 40, 40], 64000, "rcpp", 1.0, 15, mv.PRECOND_CHEB2), ([40, 40, 40], 64000, "rcpp", 1.0, 15, mv.PRECOND_CHEB3),
             ([40, 40, 40], 64000, "rcpp", 1.0, 15, mv.PRECOND_CHEB4), ([40, 40, 40], 64000, "rcpp", 1.0, 15, mv.PRECOND_AUTO),
             ([66, 12, 24], 30000, "rcpp", 0.7, 10, mv.PRECOND_AUTO), ([24, 24, 24], 20000, "cpp", 2.0, 0, mv.PRECOND_CHEB3)]
    if os.environ.get("MVTV_MG_ONLY2D"):   # short run: the 2-D cases only (k_cg_step2d's ghost-row / peer-memory protocol)
        cases = [c for c in cases if len(c[0]) == 2 and c[2] != "py"] + [([64, 64], 20000, "rcpp", 0.8, 40, C1)]
    for dims, n, mode, lam, max_passes, precond in cases:
        p = len(dims)
        imode = {"cpp": 0, "rcpp": 1, "py": 2}[mode]
        x, y = synth(41 + p, n, p, 0.0, 1.0, 0.5)
        axes = mv.mesh_axes(x, dims, mode)
        # arbitrary initial shard: round-robin, then one all-to-all to the owners
        xs, ys = x[rank::world], y[rank::world]
        xo, yo = partition.exchange_points(xs, ys, axes[-1])
        box = [mv.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        variant = mv.VARIANT_REFERENCE if (p < 3 or dims[0] == dims[1]) else mv.VARIANT_INTENDED   # reference operator: cpp-code/utils.cpp:187
        with mv.Plan(dims, variant=variant, device=local, rank=rank, world=world, nccl_unique_id=box[0]) as pl:
            pl.set_points(xo, yo, axes)
            out = pl.solve(lam, mode=mode, max_passes=max_passes, want_fitted=False, cg_rtol=1e-13, precond=precond)
            gathered = [None] * world
            dist.all_gather_object(gathered, (pl.z0, out["theta"], out["counter"], out["passes"], out["inner_iters"]))
        if rank == 0:
            gathered.sort(key=lambda t: t[0])
            theta = np.concatenate([g[1] for g in gathered])
            big = int(np.prod(dims)) > 20000
            ref = co.mbs_one(x, y, dims, axes, lam, mode=imode, max_passes=max_passes, variant=variant,
                             solver=co.SOLVER_PCG if big else co.SOLVER_BANDCHOL, cg_rtol=1e-13)
            err = float(np.abs(theta - ref["theta"]).max())
            same = all(g[2] == ref["counter"] for g in gathered)
            good = same and err <= 1e-9
            ok = ok and good
            print("%s precond=%d dims=%s mode=%s lam=%g: Counter=%d (oracle %d) passes=%d inner=%d max|dtheta|=%.2e %s"
                  % ("OK  " if good else "FAIL", precond, dims, mode, lam, gathered[0][2], ref["counter"], gathered[0][3],
                     gathered[0][4], err, ""), flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.barrier()
    dist.destroy_process_group()
    if int(flag.item()) != 1:
        sys.exit(1)
    if rank == 0:
        print("multigpu_check: all cases OK on %d GPUs" % world)


if __name__ == "__main__":
    main()
