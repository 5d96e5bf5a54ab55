"""Time-axis sharding across the GPUs of one box: one process per GPU, torch.distributed for the plumbing.

Propagator multiplication is associative but not commutative, so the N effective steps are split into `world`
contiguous slices; rank g reduces slice g to one partial propagator on its own GPU (no communication), the
partials (dim x dim each, 2 KiB .. 1 MiB) are all-gathered (NCCL over NVLink on GPUs, gloo in the CPU tests) and
multiplied in slice order, later slice on the left.  The reference has no multi-GPU path.
"""
from __future__ import annotations

import numpy as np


def slice_bounds(nsteps: int, world: int):
    """Contiguous, balanced step ranges [b[g], b[g+1]) for g < world."""
    return [nsteps * g // world for g in range(world + 1)]


def gather_partials(partial, group=None):
    """All-gather one dim x dim partial propagator per rank; returns them in rank (= time) order.

    `partial` is a numpy array (moved through a tensor on the backend's device) or a torch tensor that already
    lives where the backend needs it."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return [partial]
    was_numpy = isinstance(partial, np.ndarray)
    t = torch.from_numpy(np.ascontiguousarray(partial)) if was_numpy else partial
    if dist.get_backend(group) == "nccl" and not t.is_cuda:
        t = t.cuda()
    out = [torch.empty_like(t) for _ in range(dist.get_world_size(group))]
    dist.all_gather(out, t, group=group)
    return [o.cpu().numpy() for o in out] if was_numpy else out


def time_sliced_equiprop(ctx, dt, carr, group=None, root=0):
    """Whole-pulse propagator computed by all ranks of `group` together.

    ctx: a parament_b200.Parament (or any object with .steps_of(pts), .equiprop_slice, .combine) bound to this
    rank's GPU with the Hamiltonian already set; carr: (amps, pts), identical on every rank.
    Returns the propagator on rank `root`, None elsewhere."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    carr = np.atleast_2d(np.asarray(carr))
    nsteps = ctx.steps_of(carr.shape[1])
    b = slice_bounds(nsteps, world)
    part = ctx.equiprop_slice(dt, carr, b[rank], b[rank + 1])
    parts = gather_partials(part, group)
    if rank != root:
        return None
    return ctx.combine(np.stack(parts))
