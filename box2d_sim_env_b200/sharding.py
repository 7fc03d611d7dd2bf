"""World sharding across the GPUs of one box.

Worlds are independent (each Box2DWorld owns its own b2World, include/sim_env/Box2DWorld.h:854), so rank r of R owns
the contiguous slice [r * N / R, (r + 1) * N / R); the model is replicated and inputs are generated from
(seed, world index), so nothing is scattered.  The only collective on the path is the final gather of rollout
states (or costs): one all_gather over NCCL / NVLink (gloo in the CPU tests).
"""
import numpy as np


def world_slice(n_total, rank, world_size):
    """Contiguous slice of worlds owned by `rank`; the first n_total % world_size ranks get one extra world."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank %d of %d" % (rank, world_size))
    base, extra = divmod(int(n_total), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def slice_sizes(n_total, world_size):
    return [world_slice(n_total, r, world_size)[1] - world_slice(n_total, r, world_size)[0] for r in range(world_size)]


def seeded_rows(n_total, lo, hi, seed, draw):
    """Rows lo:hi of a [n_total, ...] array that `draw(rng, n_total)` would produce — every rank draws from the same
    counter-based stream and keeps its own rows, so the union over ranks is independent of the GPU count."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    return draw(rng, n_total)[lo:hi]


def gather_states(local, n_total, group=None, out=None):
    """All-gather the per-rank [n_local, D] state rows into the full [n_total, D] tensor on every rank.

    Uses all_gather_into_tensor when every rank holds the same number of rows (the benchmark configs), and the
    padded list form otherwise.  `local` is a torch tensor on the rank's device (CUDA for nccl, CPU for gloo)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local if out is None else out.copy_(local)
    ws = dist.get_world_size(group)
    sizes = slice_sizes(n_total, ws)
    tail = tuple(local.shape[1:])
    if out is None:
        out = torch.empty((n_total,) + tail, dtype=local.dtype, device=local.device)
    if len(set(sizes)) == 1:
        dist.all_gather_into_tensor(out.view(-1), local.contiguous().view(-1), group=group)
        return out
    width = max(sizes)
    padded = torch.zeros((width,) + tail, dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    parts = [torch.empty_like(padded) for _ in range(ws)]
    dist.all_gather(parts, padded, group=group)
    pos = 0
    for r in range(ws):
        out[pos:pos + sizes[r]] = parts[r][:sizes[r]]
        pos += sizes[r]
    return out
