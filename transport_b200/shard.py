"""Multi-GPU sharding of the (parameter x energy) grid.

Every point is an independent solve, so there is no exchange step inside the computation: rank r
takes a contiguous slice of the parameter axis (its Hamiltonians stay local and each is reused over
the whole energy axis); when there are fewer parameters than ranks the energy axis is split instead.
The only collective is the gather of the compact per-point totals (`gather_totals`), NCCL on GPUs,
gloo in the CPU tests.
"""
import numpy as np


def split_range(n_items, world, rank):
    """Contiguous near-equal partition: returns (start, stop) of rank's slice of range(n_items)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad world/rank")
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def plan(n_params, n_energies, world, rank):
    """Which (parameter slice, energy slice) this rank evaluates.

    Parameter axis first (keeps every Hamiltonian on one GPU); if n_params < world the energy axis is
    split instead and the Hamiltonians are replicated (they are at most tens of MB)."""
    if n_params >= world:
        return split_range(n_params, world, rank), (0, n_energies)
    return (0, n_params), split_range(n_energies, world, rank)


def gather_totals(local, world, rank, n_params, n_energies, group=None):
    """All-gather the per-point totals [local_params, local_energies] into the full [n_params, n_energies]
    array on every rank.  `local` is a torch tensor on the device the process group works with."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return local
    sizes = []
    for r in range(world):
        (p0, p1), (e0, e1) = plan(n_params, n_energies, world, r)
        sizes.append(((p0, p1), (e0, e1)))
    max_elems = max((p1 - p0) * (e1 - e0) for (p0, p1), (e0, e1) in sizes)
    pad = torch.zeros(max_elems, dtype=local.dtype, device=local.device)
    pad[:local.numel()] = local.reshape(-1)
    out = torch.empty(world * max_elems, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    full = torch.empty((n_params, n_energies), dtype=local.dtype, device=local.device)
    for r, ((p0, p1), (e0, e1)) in enumerate(sizes):
        cnt = (p1 - p0) * (e1 - e0)
        full[p0:p1, e0:e1] = out[r * max_elems:r * max_elems + cnt].reshape(p1 - p0, e1 - e0)
    return full
