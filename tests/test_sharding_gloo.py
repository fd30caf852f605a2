"""N > 1 host logic on CPU: world_size-2 gloo processes shard a batch by cosmology and assemble
the C_ell blocks on rank 0 with one gather (ccl_b200/sharding.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ccl_b200 import sharding


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ncosmo, npair, nl, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        off, cnt = sharding.shard_range(ncosmo, world, rank)
        # stand-in for the device result of this rank's block: value encodes (cosmology, pair, ell)
        ic = torch.arange(off, off + cnt, dtype=torch.float64)[:, None, None]
        ip = torch.arange(npair, dtype=torch.float64)[None, :, None]
        il = torch.arange(nl, dtype=torch.float64)[None, None, :]
        local = ic * 1e6 + ip * 1e3 + il
        out = sharding.gather_blocks(local, sharding.shard_sizes(ncosmo, world), dst=0)
        if rank == 0:
            q.put(out.numpy())
        else:
            assert out is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("ncosmo", [8, 7])   # equal shards -> one gather; ragged -> send/recv
def test_gather_by_cosmology_world2(ncosmo):
    world, npair, nl = 2, 6, 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ncosmo, npair, nl, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    ic, ip, il = np.meshgrid(np.arange(ncosmo), np.arange(npair), np.arange(nl), indexing="ij")
    assert got.shape == (ncosmo, npair, nl)
    assert np.array_equal(got, ic * 1e6 + ip * 1e3 + il)


def test_shard_ranges_cover_exactly():
    for n in (0, 1, 7, 4096, 4099):
        for world in (1, 2, 4, 8):
            spans = [sharding.shard_range(n, world, r) for r in range(world)]
            assert sum(c for _, c in spans) == n
            assert all(spans[r][0] + spans[r][1] == spans[r + 1][0] for r in range(world - 1))
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def test_balance_tiles_on_nodes():
    rng = np.random.default_rng(0)
    w = rng.integers(17, 600, size=2100)          # nk per (pair, ell-tile), SURVEY 8e
    parts = sharding.balance_tiles(w, 8)
    assert sorted(np.concatenate(parts).tolist()) == list(range(w.size))
    loads = np.array([w[p].sum() for p in parts])
    assert loads.max() / loads.mean() < 1.01


def _worker_ell(rank, world, port, npair, nl, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        idx = sharding.interleaved_indices(nl, world, rank)
        ip = torch.arange(npair, dtype=torch.float64)[:, None]
        local = ip * 1e4 + torch.as_tensor(idx, dtype=torch.float64)[None, :]   # stand-in for cl[pair, ell shard]
        out = sharding.gather_interleaved(local, nl, dst=0)
        if rank == 0:
            q.put(out.numpy())
        else:
            assert out is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nl", [10, 11])
def test_gather_by_ell_tile_world2(nl):
    """Single-cosmology problems (BASELINE configs[4]) shard the ell axis; one gather reassembles it."""
    world, npair = 2, 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_ell, args=(r, world, port, npair, nl, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    ip, il = np.meshgrid(np.arange(npair), np.arange(nl), indexing="ij")
    assert np.array_equal(got, ip * 1e4 + il)


def test_pipeline_chunk_sizes_and_slices():
    """Host logic of the pipelined host-buffer entry: chunk sizes cover the batch, ramp up, respect the
    regular size; slice_stacked cuts every per-cosmology array and keeps shared grids."""
    from ccl_b200.pipeline import chunk_sizes, slice_stacked
    assert chunk_sizes(0) == []
    for n in (1, 37, 100, 256, 512, 1000, 4096, 5000):
        for k in (None, 1, 4, 8, 16):
            sz = chunk_sizes(n, k)
            assert sum(sz) == n and min(sz) >= 1
            assert sz[0] <= max(sz) and sz[-1] <= max(sz) and sz[0] == sz[-1] or len(sz) < 3   # small at both ends
            if k is None:
                assert len(sz) <= 16 + 12 and (n < 1024 or max(sz) >= 64)
    assert chunk_sizes(512) == [32, 48, 64, 64, 64, 64, 64, 32, 48, 32]
    assert sum(chunk_sizes(4096)) == 4096 and max(chunk_sizes(4096)) == 256 and chunk_sizes(4096)[0] == 32
    n, nz = 6, 4
    stk = dict(count=n, achi=(np.arange(n), np.arange(n * 5.).reshape(n, 5), np.ones((n, 5))),
               growth=(np.linspace(0.1, 1, 3), np.arange(n * 3.).reshape(n, 3)), chi_max_nokernel=np.arange(n, dtype=float),
               pk=(np.zeros(7), np.zeros(2), np.arange(n * 14.).reshape(n, 2, 7), 1, 2, True),
               tracers=[dict(der_bessel=0, chi=np.arange(n * nz, dtype=float).reshape(n, nz), w=np.ones((n, nz)),
                             a=np.linspace(0.5, 1, 3), fa=np.arange(n * 3.).reshape(n, 3))])
    cut = slice_stacked(stk, 2, 5)
    assert cut["count"] == 3 and cut["achi"][1].shape == (3, 5) and cut["achi"][1][0, 0] == 10.0
    assert cut["growth"][0].shape == (3,) and cut["growth"][1][0, 0] == 6.0           # shared a grid kept
    assert cut["pk"][2].shape == (3, 2, 7) and cut["pk"][0] is stk["pk"][0]
    t = cut["tracers"][0]
    assert t["chi"].shape == (3, nz) and t["chi"][0, 0] == 2 * nz and t["a"].shape == (3,) and t["fa"][0, 0] == 6.0
    assert np.shares_memory(cut["pk"][2], stk["pk"][2])                                # views, no copies


def _toy_stacked(n=2):
    return dict(count=n, achi=(np.array([4, 5][:n]), np.arange(n * 5.).reshape(n, 5), np.ones((n, 5))),
                growth=(np.linspace(0.1, 1, 3), np.arange(n * 3.).reshape(n, 3)), chi_max_nokernel=np.arange(n, dtype=float),
                pk=(np.zeros(7), np.zeros(2), np.arange(n * 14.).reshape(n, 2, 7), 1, 2, True),
                tracers=[dict(der_bessel=-1, der_angles=2, chi=np.arange(n * 4, dtype=float).reshape(n, 4), w=np.ones((n, 4)),
                              a=np.linspace(0.5, 1, 3), fa=np.arange(n * 3.).reshape(n, 3)),
                         dict(der_bessel=0, chi=np.arange(4.), w=np.ones((n, 4)))])


def test_pack_unpack_stacked_roundtrip():
    stk = _toy_stacked()
    meta, flat = sharding.pack_stacked(stk)
    back = sharding.unpack_stacked(meta, flat)
    assert back["count"] == 2 and np.array_equal(back["achi"][0], [4, 5])
    assert np.array_equal(back["pk"][2], stk["pk"][2]) and back["pk"][3:] == (1, 2, True)
    assert back["tracers"][0]["der_angles"] == 2 and np.array_equal(back["tracers"][0]["fa"], stk["tracers"][0]["fa"])
    assert np.array_equal(back["tracers"][1]["chi"], np.arange(4.)) and "fa" not in back["tracers"][1]


def _worker_tiles(rank, world, port, npair, nl, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # the table pack lives on rank 0 only and reaches the others with one broadcast
        stk = sharding.broadcast_stacked(_toy_stacked() if rank == 0 else None, src=0)
        assert np.array_equal(stk["pk"][2], _toy_stacked()["pk"][2]) and stk["tracers"][0]["der_bessel"] == -1
        # every rank derives the same ell plan from the same weights (nk falls with ell)
        plan = sharding.ell_tile_plan(np.linspace(600., 300., nl), world)
        mine = plan[rank]
        ip = torch.arange(npair, dtype=torch.float64)[None, :, None]
        local = ip * 1e4 + torch.as_tensor(mine, dtype=torch.float64)[None, None, :]      # stand-in for cl[1, pair, my ells]
        out = sharding.gather_ell_tiles(local, plan, nl, dst=0)
        if rank == 0:
            q.put((out.numpy(), [p.tolist() for p in plan]))
        else:
            assert out is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nl", [12, 13])
def test_tile_sharded_plan_broadcast_gather_world2(nl):
    """configs[4] host logic: broadcast of the table pack, node-balanced ell tiles, one gather."""
    world, npair = 2, 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_tiles, args=(r, world, port, npair, nl, q)) for r in range(world)]
    for p in procs:
        p.start()
    got, plan = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert sorted(plan[0] + plan[1]) == list(range(nl))
    w = np.linspace(600., 300., nl)
    assert abs(w[plan[0]].sum() - w[plan[1]].sum()) <= w.max()
    ip, il = np.meshgrid(np.arange(npair), np.arange(nl), indexing="ij")
    assert got.shape == (1, npair, nl) and np.array_equal(got[0], ip * 1e4 + il)
