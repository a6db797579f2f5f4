"""Host-side logic of the multi-GPU path, on CPU with world_size-2/3 gloo process groups:
slab bounds, point ownership, the setup all-to-all, and the ghost-plane protocol (halo exchange + clamped
stencil on a slab == the global operator).  The device kernels are not involved; the checker is the oracle."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_slab_bounds_and_ownership():
    from multivartv_b200 import partition
    for m, w in [(10, 3), (512, 8), (7, 7), (9, 2)]:
        b = partition.slab_bounds(m, w)
        assert b[0][0] == 0 and sum(n for _, n in b) == m
        assert all(b[i][0] + b[i][1] == b[i + 1][0] for i in range(w - 1))
        assert max(n for _, n in b) - min(n for _, n in b) <= 1
    ax = np.linspace(0, 1, 10)
    x = np.array([0.0, 0.05, 0.0555555, 0.0555556, 0.5, 0.99, 1.0])
    z = partition.nearest_knot(x, ax)
    assert list(z) == [0, 0, 0, 1, 4, 9, 9] or list(z) == [0, 0, 1, 1, 4, 9, 9]
    own = partition.owner_of(x, ax, 3)          # slabs: [0,4) [4,7) [7,10)
    assert list(own[[0, 4, 6]]) == [0, 1, 2]
    for r, (z0, nz) in enumerate(partition.slab_bounds(10, 3)):
        lo, hi = partition.slab_interval(ax, z0, nz)
        mid = 0.5 * (lo + hi)
        assert partition.owner_of(np.array([mid]), ax, 3)[0] == r


def _worker(rank, world, port, q):
    try:
        sys.path.insert(0, ROOT)
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        import torch
        import torch.distributed as dist

        from multivartv_b200 import partition
        from oracle import c_oracle as co
        dist.init_process_group("gloo", rank=rank, world_size=world)
        dims = [6, 6, 7]   # m0 == m1: the reference operator exists (cpp-code/utils.cpp:187,216)
        N = int(np.prod(dims))
        plane = dims[0] * dims[1]
        rng = np.random.RandomState(9)
        n = 500
        x = rng.uniform(0, 1, (n, 3))
        y = rng.normal(size=n)
        axes = [np.linspace(0, 1, d) for d in dims]
        # 1. all-to-all of an arbitrary shard -> every rank ends with exactly its slab's points
        xo, yo = partition.exchange_points(x[rank::world], y[rank::world], axes[-1])
        z0, nz = partition.slab_bounds(dims[-1], world)[rank]
        idx = co.nearest(dims, axes, xo) if len(yo) else np.zeros(0, dtype=np.int64)
        zz = idx // plane
        assert np.all((zz >= z0) & (zz < z0 + nz))
        counts = [None] * world
        dist.all_gather_object(counts, len(yo))
        assert sum(counts) == n
        # 2. per-slab Oty / counts assemble to the global ones
        Oty_g, cnt_g = co.scatter(co.nearest(dims, axes, x), y, N)
        Oty_l = np.zeros(nz * plane)
        cnt_l = np.zeros(nz * plane)
        np.add.at(Oty_l, idx - z0 * plane, yo)
        np.add.at(cnt_l, idx - z0 * plane, 1.0)
        assert np.allclose(Oty_l, Oty_g[z0 * plane:(z0 + nz) * plane], rtol=0, atol=1e-12)
        assert np.array_equal(cnt_l, cnt_g[z0 * plane:(z0 + nz) * plane])
        # 3. ghost-plane protocol: ghosted slab, one plane to each neighbour, clamped 27-point stencil of
        #    D^T D on the slab == the global D^T(D x)
        op = co.Operator(dims)
        xg = rng.normal(size=N)                          # same on every rank (same seed)
        ref = op.Dt(op.D(xg))
        slab = np.zeros((nz + 2, plane))
        slab[1:nz + 1] = xg[z0 * plane:(z0 + nz) * plane].reshape(nz, plane)
        reqs = []
        if rank + 1 < world:
            reqs.append(dist.isend(torch.from_numpy(slab[nz].copy()), rank + 1))
        if rank > 0:
            reqs.append(dist.isend(torch.from_numpy(slab[1].copy()), rank - 1))
        if rank > 0:
            t = torch.empty(plane, dtype=torch.float64)
            dist.recv(t, rank - 1)
            slab[0] = t.numpy()
        if rank + 1 < world:
            t = torch.empty(plane, dtype=torch.float64)
            dist.recv(t, rank + 1)
            slab[nz + 1] = t.numpy()
        for r_ in reqs:
            r_.wait()
        # stencil coefficients exactly as solver.cu build_tables derives them
        masks, scales = op.masks, op.scales
        t3 = np.array([-1.0, 2.0, -1.0])
        coef = np.zeros((3, 3, 3))                      # [dz, dy, dx]
        for S, sc in zip(masks, scales):
            w = [t3 if (S >> a) & 1 else np.array([0.0, 1.0, 0.0]) for a in range(3)]
            coef += sc * sc * np.einsum("k,j,i->kji", w[2], w[1], w[0])
        vol = slab.reshape(nz + 2, dims[1], dims[0])
        out = np.zeros((nz, dims[1], dims[0]))
        for zl in range(nz):
            gz = z0 + zl
            for dz in (-1, 0, 1):
                zs = gz + dz
                zs = min(max(zs, 0), dims[2] - 1)       # clamp at the GLOBAL boundary only
                src = vol[zs - z0 + 1]
                for dy in (-1, 0, 1):
                    ys = np.clip(np.arange(dims[1]) + dy, 0, dims[1] - 1)
                    for dx in (-1, 0, 1):
                        xs = np.clip(np.arange(dims[0]) + dx, 0, dims[0] - 1)
                        out[zl] += coef[dz + 1, dy + 1, dx + 1] * src[np.ix_(ys, xs)]
        err = np.abs(out.reshape(-1) - ref[z0 * plane:(z0 + nz) * plane]).max()
        assert err < 1e-12, err
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL: %s\n%s" % (e, traceback.format_exc())))


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_exchange_and_ghost_protocol(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
