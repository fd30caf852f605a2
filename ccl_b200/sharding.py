"""Multi-GPU sharding of the Limber path.

Every (cosmology, pair, ell) is independent: there is no halo and no reduction.  A batch is
partitioned by cosmology into contiguous blocks, one process per GPU computes its block, and
the C_ell blocks are assembled on the caller's rank with ONE collective (a gather).  For a
single cosmology with many pairs (BASELINE configs[4]) the pair x ell-tile list is partitioned
instead, balanced on the number of integrand nodes because nk varies 17..600 per C_ell.

The reference has no counterpart (it is single-process OpenMP over ell, src/ccl_cls.c:282-356).
"""
import numpy as np


def shard_range(n, world, rank):
    """Contiguous block of `n` items owned by `rank`: (offset, count); remainder to the first ranks."""
    base, rem = divmod(int(n), int(world))
    return rank * base + min(rank, rem), base + (1 if rank < rem else 0)


def shard_sizes(n, world):
    return [shard_range(n, world, r)[1] for r in range(world)]


def balance_tiles(weights, world):
    """Longest-processing-time partition of tiles with the given weights (integrand nodes).

    Returns a list of index arrays, one per rank; deterministic (ties broken by tile index)."""
    w = np.asarray(weights, dtype=np.float64)
    order = np.lexsort((np.arange(w.size), -w))
    load = np.zeros(world)
    owner = np.empty(w.size, dtype=np.int64)
    for t in order:
        r = int(np.argmin(load))
        owner[t] = r
        load[r] += w[t]
    return [np.flatnonzero(owner == r) for r in range(world)]


def make_tiles(npair, nl, pair_tile=32, ell_tile=128):
    """Rectangular (pair, ell) tiles covering an [npair, nl] C_ell block: list of (p0, p1, l0, l1)."""
    return [(p0, min(p0 + pair_tile, npair), l0, min(l0 + ell_tile, nl))
            for p0 in range(0, int(npair), int(pair_tile)) for l0 in range(0, int(nl), int(ell_tile))]


def tile_weights(batch, collections, pairs, ells, tiles):
    """Integrand nodes (sum of nk over the tile's (pair, ell) entries, first cosmology of the batch) of every
    tile: the work measure balance_tiles partitions on (nk varies 17..600 per C_ell)."""
    ells = np.asarray(ells, dtype=np.float64)
    return np.array([batch.spline_nodes(collections, list(pairs[p0:p1]), ells[l0:l1]) for p0, p1, l0, l1 in tiles],
                    dtype=np.float64)


def gather_blocks(local, sizes, dst=0, group=None):
    """Assemble per-rank blocks [sizes[r], ...] on rank `dst` with one collective.

    Equal-sized shards use one gather (NCCL over NVLink on GPU tensors, gloo on CPU tensors);
    ragged shards use grouped point-to-point transfers into the preallocated output.
    Returns the concatenated tensor on `dst`, None elsewhere."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    assert len(sizes) == world and local.shape[0] == sizes[rank]
    tail = tuple(local.shape[1:])
    out = None
    if rank == dst:
        out = torch.empty((int(sum(sizes)),) + tail, dtype=local.dtype, device=local.device)
    if len(set(sizes)) == 1:
        views = None
        if rank == dst:
            views = list(out.split(sizes[0], dim=0))
        dist.gather(local.contiguous(), views, dst=dst, group=group)
        return out
    if rank == dst:
        offs = np.concatenate([[0], np.cumsum(sizes)])
        out[offs[dst]:offs[dst + 1]].copy_(local)
        reqs = [dist.irecv(out[offs[r]:offs[r + 1]], src=r, group=group)
                for r in range(world) if r != dst and sizes[r] > 0]
        for q in reqs:
            q.wait()
        return out
    if sizes[rank] > 0:
        dist.send(local.contiguous(), dst=dst, group=group)
    return None


def interleaved_indices(n, world, rank):
    """Indices rank::world of an axis of length n: the ell-tile shard of a single-cosmology problem
    (BASELINE configs[4]); interleaving balances the slowly varying work per ell (nk grows with ell)."""
    return np.arange(int(rank), int(n), int(world))


def gather_interleaved(local, n, dst=0, group=None):
    """Reassemble an axis-(-1) interleaved shard: local[..., j] holds global index rank + j * world.

    Shards may differ by one element; they are padded to a common length for ONE gather and the
    padding is dropped on `dst`.  Returns the full [..., n] tensor on `dst`, None elsewhere."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    nmax = (n + world - 1) // world
    pad = torch.zeros(tuple(local.shape[:-1]) + (nmax,), dtype=local.dtype, device=local.device)
    pad[..., :local.shape[-1]] = local
    parts = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, parts, dst=dst, group=group)
    if rank != dst:
        return None
    out = torch.empty(tuple(local.shape[:-1]) + (n,), dtype=local.dtype, device=local.device)
    for r in range(world):
        idx = interleaved_indices(n, world, r)
        out[..., idx] = parts[r][..., :idx.size]
    return out
