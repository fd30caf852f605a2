"""Host-buffer entry for large likelihood batches: stage -> upload/prepare -> integrate, pipelined.

`PipelinedBatch.angular_cl(stacked_tables, ...)` splits the batch into chunks of cosmologies, each
with its own ccl_b200_batch (arena + non-blocking stream), and runs three stages on three host
threads so that they overlap across chunks:
  1. planning, and for pageable inputs staging of the caller's stacked host tables into the chunk's
     pinned arena (host threads); stacked arrays that already live in pinned memory (pin_stacked)
     are not staged at all,
  2. H2D copies (one DMA per pinned stacked array, one for the staged arena) + device table
     preparation (copy engine + small kernels),
  3. the Limber kernels + D2H of the chunk's C_ell block (SMs + copy engine).
The C ABI calls release the GIL, so the stages really run concurrently.  Results are identical to
a single LibBatch (same kernels on the same tables).  No counterpart in the reference, which has
no batch axis at all (pairs and cosmologies are Python loops, benchmarks/test_timing.py:31-34).
"""
import queue
import threading
import time

import numpy as np

from . import _lib
from .lowlevel import LibBatch

__all__ = ("PipelinedBatch", "slice_stacked", "pin_stacked", "chunk_sizes")


def slice_stacked(stk, lo, hi):
    """Rows [lo, hi) of a stacked-table dict (see LibBatch.set_stacked); shared grids are kept."""
    def rows(x):
        x = np.asarray(x)
        return x[lo:hi]

    n_per, chi, a = stk["achi"]
    out = dict(count=hi - lo, achi=(rows(n_per), rows(chi), rows(a)), tracers=[])
    if stk.get("growth") is not None:
        ag, D = stk["growth"]
        out["growth"] = (ag if np.ndim(ag) == 1 else rows(ag), rows(D))
    if stk.get("chi_max_nokernel") is not None:
        out["chi_max_nokernel"] = rows(stk["chi_max_nokernel"])
    lk, apk, pk, olo, ohi, islog = stk["pk"]
    out["pk"] = (lk, apk, rows(pk), olo, ohi, islog)
    for t in stk["tracers"]:
        u = dict(t)
        for key in ("chi", "a"):
            if t.get(key) is not None and np.ndim(t[key]) == 2:
                u[key] = rows(t[key])
        for key in ("w", "fka", "fk", "fa"):
            if t.get(key) is not None:
                u[key] = rows(t[key])
        out["tracers"].append(u)
    return out


def pin_stacked(stk):
    """Copy of a stacked-table dict whose float64 arrays live in pinned (page-locked) host memory.

    The stacked setters recognise pinned sources (cudaPointerGetAttributes) and skip the staging
    copy: commit() moves each stacked array with one DMA straight from the caller's buffer, which
    therefore has to stay alive and unchanged until commit() returns (LibBatch keeps a reference)."""
    import torch

    def pin(x):
        if x is None or np.ndim(x) == 0:
            return x
        x = np.asarray(x)
        if x.dtype != np.float64:
            return x
        t = torch.empty(x.shape, dtype=torch.float64, pin_memory=True)
        v = t.numpy()
        v[...] = x
        return v

    out = dict(stk)
    n_per, chi, a = stk["achi"]
    out["achi"] = (n_per, pin(chi), pin(a))
    if stk.get("growth") is not None:
        out["growth"] = tuple(pin(x) for x in stk["growth"])
    lk, apk, pk, olo, ohi, islog = stk["pk"]
    out["pk"] = (lk, apk, pin(pk), olo, ohi, islog)
    out["tracers"] = [{k: (pin(v) if k in ("chi", "w", "a", "fka", "fk", "fa") else v) for k, v in t.items()}
                      for t in stk["tracers"]]
    return out


def chunk_sizes(ncosmo, nchunks=None):
    """Cosmologies per chunk of the pipelined entry.

    Regular chunks hold >= 64 cosmologies, at most 16 of them unless `nchunks` says otherwise (a chunk only
    reaches the SMs after its own planning + upload + preparation, 6 ms for 200 cosmologies on a busy GPU, so
    chunks must stay small against the whole batch); the sizes
    ramp up x1.5 from an eighth of the regular size (not below 32), and down again at the end (chunks are queued
    asynchronously, so small chunks leave no gaps behind them)."""
    ncosmo = int(ncosmo)
    if ncosmo <= 0:
        return []
    if nchunks is None:
        nchunks = max(1, min(16, ncosmo // 64))
    nchunks = max(1, min(int(nchunks), ncosmo))
    base = -(-ncosmo // nchunks)
    # ramp up x1.5 from `first` to the regular size, plateau, then the same ramp down: a chunk's planning + upload +
    # preparation latency is only exposed for the first chunk (nothing to overlap it with yet) and for the last
    # ones (the SMs drain while they are still being committed), so both ends are small
    first = max(1, base // 8, min(32, base))
    up, size = [], first
    while size < base and 2 * (sum(up) + size) <= ncosmo:
        up.append(size)
        size = -(-size * 3 // 2)
    rem = ncosmo - 2 * sum(up)
    mid = []
    while rem > 0:
        take = min(base, rem)
        mid.append(take)
        rem -= take
    if len(mid) > 1 and mid[-1] < first:      # fold a tiny remainder into its neighbour
        mid[-2] += mid.pop()
    return up + mid + up[::-1]


class PipelinedBatch:
    NSTAGE = 2   # host threads planning/staging chunks concurrently
    NCOMMIT = 2  # host threads uploading + preparing chunks concurrently: a chunk's ~10 small preparation kernels
                 # each wait for SM slots while an earlier chunk's Limber grid is running, so one serial commit
                 # thread is latency-bound (3.9 ms per 64 cosmologies against 1.3 ms on an idle GPU)

    def __init__(self, ncosmo, nchunks=None, params=None):
        import os
        self.NSTAGE = int(os.environ.get("CCL_B200_NSTAGE", self.NSTAGE))
        self.NCOMMIT = int(os.environ.get("CCL_B200_NCOMMIT", self.NCOMMIT))
        self.ncosmo = int(ncosmo)
        sizes = chunk_sizes(self.ncosmo, nchunks)
        offs = np.concatenate([[0], np.cumsum(sizes)])
        self.spans = [(int(offs[c]), int(sizes[c])) for c in range(len(sizes))]
        self.nchunks = len(sizes)
        self.batches = [LibBatch(cnt, params) for _, cnt in self.spans]
        self.trace = None     # set to [] to collect (stage, chunk, t_start, t_end) of the next call

    @property
    def h2d_bytes(self):
        return sum(int(b.h2d_bytes) for b in self.batches)

    def angular_cl(self, stk, collections, pairs, ells, method=_lib.INTEGRATION_SPLINE, out=None):
        """cl[ncosmo, npair, nl], status[ncosmo, npair], status -- like LibBatch.angular_cl."""
        nl = len(np.atleast_1d(ells))
        shape = (self.ncosmo, len(pairs), nl)
        if out is None:
            out = np.empty(shape)
        assert out.shape == shape and out.dtype == np.float64 and out.flags.c_contiguous
        stat = np.zeros((self.ncosmo, len(pairs)), dtype=np.int32)
        ready = queue.Queue()
        errors = []
        trace = self.trace

        def timed(name, c, fn, *a, **k):
            if trace is None:
                return fn(*a, **k)
            t0 = time.perf_counter()
            r = fn(*a, **k)
            trace.append((name, c, t0, time.perf_counter()))
            return r

        nchunks = len(self.spans)
        staged = [threading.Event() for _ in range(nchunks)]
        failed = threading.Event()

        def stage(first):
            # planning is serial host work per chunk: NSTAGE threads take alternate chunks
            try:
                for c in range(first, nchunks, self.NSTAGE):
                    if failed.is_set():
                        break
                    off, cnt = self.spans[c]
                    b = self.batches[c]
                    b.reset()
                    timed("stage", c, b.set_stacked, 0, slice_stacked(stk, off, off + cnt))
                    staged[c].set()
            except BaseException as e:  # noqa: BLE001 -- re-raised on the caller's thread
                errors.append(e)
                failed.set()
                for ev in staged:
                    ev.set()

        def commit(first):
            try:
                for c in range(first, nchunks, self.NCOMMIT):
                    staged[c].wait()
                    if failed.is_set():
                        break
                    timed("commit", c, self.batches[c].commit)
                    ready.put(c)
            except BaseException as e:  # noqa: BLE001
                errors.append(e)
                failed.set()
            finally:
                ready.put(None)

        threads = [threading.Thread(target=stage, args=(k,), daemon=True) for k in range(self.NSTAGE)]
        threads += [threading.Thread(target=commit, args=(k,), daemon=True) for k in range(self.NCOMMIT)]
        for t in threads:
            t.start()
        # the Limber kernels and the D2H copy of every chunk are QUEUED as soon as its tables are ready (each chunk
        # on its own stream) and collected afterwards: the SMs never wait for the host between chunks, and the tail
        # of one chunk's grid overlaps the head of the next
        started, done = [], 0
        while done < self.NCOMMIT:
            c = ready.get()
            if c is None:
                done += 1
                continue
            off, cnt = self.spans[c]
            try:
                timed("start", c, self.batches[c].angular_cl_start, collections, pairs, ells, method, out=out[off:off + cnt])
                started.append(c)
            except BaseException as e:  # noqa: BLE001
                errors.append(e)
                failed.set()
        worst = 0
        for c in started:
            off, cnt = self.spans[c]
            _, s_c, st = timed("wait", c, self.batches[c].angular_cl_wait)
            stat[off:off + cnt] = s_c
            worst = worst or st
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        return out, stat, worst
