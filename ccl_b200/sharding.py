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


# ---------------------------------------------------------------------------------------------
# Product entries: the two partitionings of SURVEY.md 8e behind one call each
# ---------------------------------------------------------------------------------------------
def pack_stacked(stk):
    """Flatten a stacked-table dict (LibBatch.set_stacked) into (meta, flat float64 array) for ONE broadcast."""
    arrays, meta = [], {"count": int(stk["count"]), "tracers": []}

    def put(x):
        if x is None:
            return None
        x = np.ascontiguousarray(x, dtype=np.float64)
        arrays.append(x.ravel())
        return x.shape

    n_per, chi, a = stk["achi"]
    meta["achi"] = ([int(v) for v in np.asarray(n_per)], put(chi), put(a))
    meta["growth"] = None if stk.get("growth") is None else tuple(put(x) for x in stk["growth"])
    meta["chi_max_nokernel"] = None if stk.get("chi_max_nokernel") is None else put(stk["chi_max_nokernel"])
    lk, apk, pk, olo, ohi, islog = stk["pk"]
    meta["pk"] = (put(lk), put(apk), put(pk), int(olo), int(ohi), bool(islog))
    for t in stk["tracers"]:
        m = {k: v for k, v in t.items() if k not in ("chi", "w", "a", "lk", "fka", "fk", "fa")}
        for k in ("chi", "w", "a", "lk", "fka", "fk", "fa"):
            m[k] = put(t.get(k))
        meta["tracers"].append(m)
    return meta, (np.concatenate(arrays) if arrays else np.zeros(0))


def unpack_stacked(meta, flat):
    """Inverse of pack_stacked."""
    pos = [0]

    def take(shape):
        if shape is None:
            return None
        n = int(np.prod(shape)) if len(shape) else 1
        out = np.ascontiguousarray(flat[pos[0]:pos[0] + n]).reshape(shape)
        pos[0] += n
        return out

    n_per, s_chi, s_a = meta["achi"]
    stk = {"count": meta["count"], "achi": (np.asarray(n_per, dtype=np.int32), take(s_chi), take(s_a))}
    stk["growth"] = None if meta["growth"] is None else tuple(take(s_) for s_ in meta["growth"])
    stk["chi_max_nokernel"] = None if meta["chi_max_nokernel"] is None else take(meta["chi_max_nokernel"])
    s_lk, s_apk, s_pk, olo, ohi, islog = meta["pk"]
    stk["pk"] = (take(s_lk), take(s_apk), take(s_pk), olo, ohi, islog)
    stk["tracers"] = []
    for m in meta["tracers"]:
        t = {k: v for k, v in m.items() if k not in ("chi", "w", "a", "lk", "fka", "fk", "fa")}
        for k in ("chi", "w", "a", "lk", "fka", "fk", "fa"):
            v = take(m[k])
            if v is not None:
                t[k] = v
        stk["tracers"].append(t)
    return stk


def broadcast_stacked(stk, src=0, group=None, device=None):
    """The table pack of a small batch from rank `src` to every rank: one object broadcast for the shapes and ONE
    tensor broadcast (NCCL over NVLink when `device` is a CUDA device) for the 2.3 MB per cosmology of tables."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return stk
    rank = dist.get_rank(group)
    meta, flat = pack_stacked(stk) if rank == src else (None, None)
    box = [meta, None if flat is None else int(flat.size)]
    dist.broadcast_object_list(box, src=src, group=group)
    meta, n = box
    t = torch.empty(n, dtype=torch.float64, device=device)
    if rank == src:
        t.copy_(torch.from_numpy(flat))
    dist.broadcast(t, src=src, group=group)
    return stk if rank == src else unpack_stacked(meta, t.cpu().numpy())


def ell_tile_plan(nodes_per_ell, world):
    """Partition of the ell axis (tiles of ALL pairs x one ell: one launch per rank) balanced on the integrand
    nodes per ell (LPT): list of sorted index arrays, one per rank."""
    return [np.sort(p) for p in balance_tiles(np.asarray(nodes_per_ell, dtype=np.float64), world)]


def nodes_per_ell(batch, collections, pairs, ells):
    """Sum over pairs (and the batch's cosmologies) of nk for every ell: the weights of ell_tile_plan."""
    ells = np.asarray(ells, dtype=np.float64)
    return np.array([batch.spline_nodes(collections, pairs, ells[i:i + 1]) for i in range(ells.size)], dtype=np.float64)


def gather_ell_tiles(local, plan, nl, dst=0, group=None):
    """Reassemble cl[..., ell] from per-rank shards local[..., len(plan[rank])] with ONE gather (shards are padded
    to the longest one).  Returns the full [..., nl] tensor on `dst`, None elsewhere."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        out = torch.empty(tuple(local.shape[:-1]) + (nl,), dtype=local.dtype, device=local.device)
        out[..., torch.as_tensor(plan[0], device=local.device)] = local
        return out
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    nmax = max(len(p) for p in plan)
    pad = torch.zeros(tuple(local.shape[:-1]) + (nmax,), dtype=local.dtype, device=local.device)
    pad[..., :local.shape[-1]] = local
    parts = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, parts, dst=dst, group=group)
    if rank != dst:
        return None
    out = torch.empty(tuple(local.shape[:-1]) + (nl,), dtype=local.dtype, device=local.device)
    for r in range(world):
        if len(plan[r]):
            out[..., torch.as_tensor(plan[r], device=local.device)] = parts[r][..., :len(plan[r])]
    return out


class TileShardedLimber:
    """BASELINE configs[4] / SURVEY 8e: ONE cosmology (or a small batch), many pairs x ells, over the ranks of a
    process group.  Rank `src` holds the stacked tables; they are broadcast once, every rank commits them into its
    own arena, computes its ell tiles (all pairs x the ells ell_tile_plan assigns it, one launch) and one gather
    assembles cl[ncosmo, npair, nl] on `dst`.  With a single process it is a plain LibBatch call."""

    def __init__(self, stk, collections, pairs, ells, src=0, group=None, params=None):
        import torch
        import torch.distributed as dist
        from .lowlevel import LibBatch
        self.group = group
        self.dist = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.rank = dist.get_rank(group) if self.dist else 0
        self.world = dist.get_world_size(group) if self.dist else 1
        dev = torch.device("cuda", torch.cuda.current_device())
        stk = broadcast_stacked(stk, src=src, group=group, device=dev)
        self.collections, self.pairs = list(collections), list(pairs)
        self.ells = np.ascontiguousarray(ells, dtype=np.float64)
        self.batch = LibBatch(int(stk["count"]), params)
        self.batch.set_stacked(0, stk)
        self.batch.commit()
        # every rank computes the same plan from the same tables (no exchange)
        self.plan = ell_tile_plan(nodes_per_ell(self.batch, self.collections, self.pairs, self.ells), self.world)
        self.mine = self.plan[self.rank]
        self.ncosmo = int(stk["count"])
        self.d_cl = torch.empty((self.ncosmo, len(self.pairs), max(len(self.mine), 1)), dtype=torch.float64, device=dev)
        self.d_st = torch.zeros((2, self.ncosmo, len(self.pairs)), dtype=torch.int32, device=dev)

    def angular_cl(self, method=501, dst=0):
        """One pass: this rank's tiles, then the gather.  Returns cl[ncosmo, npair, nl] (device tensor) on `dst`."""
        import torch
        stream = torch.cuda.current_stream().cuda_stream
        if len(self.mine):
            self.batch.angular_cl_dev(self.collections, self.pairs, self.ells[self.mine], self.d_cl.data_ptr(),
                                      self.d_st.data_ptr(), method, stream)
        return gather_ell_tiles(self.d_cl[..., :len(self.mine)], self.plan, self.ells.size, dst=dst, group=self.group)


def angular_cl_by_cosmology(batch, collections, pairs, ells, sizes, method=501, out=None, d_status=None, dst=0,
                            group=None):
    """BASELINE configs[3] / SURVEY 8e: this rank's block of cosmologies (a committed LibBatch), then ONE gather of
    the C_ell blocks to `dst`.  `sizes` = cosmologies per rank (shard_sizes).  `out` / `d_status`: this rank's
    preallocated device tensors [nloc, npair, nl] / [2, nloc, npair].  Returns the assembled tensor on `dst`."""
    import torch
    stream = torch.cuda.current_stream().cuda_stream
    nloc, npair, nl = batch.ncosmo, len(pairs), len(np.atleast_1d(ells))
    if out is None:
        out = torch.empty((nloc, npair, nl), dtype=torch.float64, device="cuda")
    if d_status is None:
        d_status = torch.zeros((2, nloc, npair), dtype=torch.int32, device="cuda")
    batch.angular_cl_dev(collections, pairs, ells, out.data_ptr(), d_status.data_ptr(), method, stream)
    return gather_blocks(out, sizes, dst=dst, group=group)
