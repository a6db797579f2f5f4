"""Host-side slab partition of the LAST mesh axis across ranks (one process per GPU).

The mesh is split into `world` contiguous slabs of the last axis (contiguous in the reference's column-major
flattening, cpp-code/utils.cpp:40-52).  A point belongs to the rank that owns its nearest vertex; the
solver needs each rank to hold exactly its own points (libmvtv_b200 refuses foreign points when world > 1),
so points are bucketed once at setup and exchanged with one all-to-all.
"""
from __future__ import annotations

import numpy as np


def slab_bounds(m_last: int, world: int):
    """[(z0, nz)] per rank: contiguous, sizes differ by at most one plane (must match solver.cu build_tables)."""
    base, extra = divmod(int(m_last), int(world))
    out = []
    for r in range(world):
        z0 = r * base + min(r, extra)
        out.append((z0, base + (1 if r < extra else 0)))
    return out


def nearest_knot(x, axis):
    """argmin_j (x - axis[j])^2, ties -> lower index (cpp-code/utils.cpp:311-321 restricted to one axis)."""
    axis = np.asarray(axis, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    if axis.size == 1:
        return np.zeros(x.shape, dtype=np.int64)
    j = np.clip(np.searchsorted(axis, x), 1, axis.size - 1)
    dl = (x - axis[j - 1]) ** 2
    dr = (x - axis[j]) ** 2
    return np.where(dl <= dr, j - 1, j).astype(np.int64)


def owner_of(x_last, axis_last, world: int):
    """Rank that owns each point, from its last coordinate."""
    z = nearest_knot(x_last, axis_last)
    bounds = slab_bounds(len(axis_last), world)
    starts = np.array([b[0] for b in bounds], dtype=np.int64)
    return (np.searchsorted(starts, z, side="right") - 1).astype(np.int64)


def slab_interval(axis_last, z0: int, nz: int):
    """Interval of the last coordinate whose nearest knot lies in planes [z0, z0+nz)."""
    ax = np.asarray(axis_last, dtype=np.float64)
    lo = ax[0] if z0 == 0 else 0.5 * (ax[z0 - 1] + ax[z0])
    hi = ax[-1] if z0 + nz >= ax.size else 0.5 * (ax[z0 + nz - 1] + ax[z0 + nz])
    return float(lo), float(hi)


def bucket_points(data, y, axis_last, world: int):
    """Split (data, y) into per-owner lists, preserving point order inside each bucket."""
    data = np.asarray(data, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).ravel()
    own = owner_of(data[:, -1], axis_last, world)
    return [(data[own == r], y[own == r]) for r in range(world)]


def exchange_points(data, y, axis_last, group=None):
    """One all-to-all of (point, response) pairs: every rank contributes an arbitrary shard of the data
    and receives the points of its own slab, ordered by (source rank, original order).  Works on any
    torch.distributed backend (gloo on CPU in the tests, nccl on the GPU box)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    buckets = bucket_points(data, y, axis_last, world)
    p = np.asarray(data).shape[1]
    send = [np.concatenate([b[0], b[1][:, None]], axis=1) for b in buckets]
    me = dist.get_rank(group)
    out = []
    for src in range(world):
        box = [send[dst] if src == me else None for dst in range(world)]
        got = [None]
        dist.scatter_object_list(got, box if src == me else None, src=src, group=group)
        out.append(got[0])
    allpts = np.concatenate(out, axis=0) if out else np.zeros((0, p + 1))
    return np.ascontiguousarray(allpts[:, :p]), np.ascontiguousarray(allpts[:, p])
