"""Multi-process (gloo, world_size 2, CPU) check of the ray-sharding scheme: time-sharded
forward needs no exchange, and the allreduce of per-shard backprojections equals the
single-process adjoint.  The compute here is the oracle (CPU); the GPU kernels plug into the
same hooks (ionotomo_b200.sharding.allreduce_sum_ as ``reduce_fn``)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import ionotomo_oracle as O
    from tests.problems import small_problem
    from ionotomo_b200 import sharding
    P = small_problem(42, 4, 5, 6, 16, 10, 9, 12)
    rays = O.cast_ray(P["origins"], P["directions"], P["tmax"], 16)
    i0 = 1
    g = O.forward_equation(rays, P["K_ne"], P["xvec"], P["yvec"], P["zvec"], P["m"], i0)
    rng = np.random.RandomState(0)
    dobs = g + 0.01 * rng.normal(size=g.shape)
    CdCt = np.full(g.shape, 1e-4)
    t0, t1 = sharding.time_shard(rays.shape[1], rank, world)
    # forward on the shard alone reproduces the full forward on those times
    g_loc = O.forward_equation(rays[:, t0:t1], P["K_ne"], P["xvec"], P["yvec"], P["zvec"], P["m"], i0)
    assert np.array_equal(g_loc, g[:, t0:t1])
    dd = O.weighted_residual(g_loc, dobs[:, t0:t1], CdCt[:, t0:t1])
    acc = O.backproject(rays[:, t0:t1], P["xvec"], P["yvec"], P["zvec"], O.adjoint_ray_coefficients(dd, i0))
    acc_t = torch.from_numpy(acc)
    sharding.allreduce_sum_(acc_t)
    S = sharding.sharded_misfit(torch.tensor(O.misfit(g_loc, dobs[:, t0:t1], CdCt[:, t0:t1])))
    grad = O.ne_from_m(P["m"], P["K_ne"]) * acc_t.numpy()
    ref = O.gradient_exact(rays, g, dobs, i0, P["K_ne"], P["xvec"], P["yvec"], P["zvec"], P["m"], CdCt)
    ok = np.abs(grad - ref).max() <= 1e-12 * np.abs(ref).max() and abs(S - O.misfit(g, dobs, CdCt)) <= 1e-12 * S
    assert sharding.world() == (rank, world)
    if rank == 0:
        out.put(bool(ok))
    dist.destroy_process_group()


def test_time_shard_partition():
    from ionotomo_b200.sharding import time_shard
    for Nt in (1, 7, 100, 101):
        for world in (1, 2, 3, 8):
            blocks = [time_shard(Nt, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == Nt
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_sharded_gradient_gloo_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert out.get(timeout=5) is True


def test_released_slices_are_rank_independent():
    """The overlapped apply hands the reducer the same equal slices in the same order on every rank,
    whatever the rank's own chunk boundaries are (a collective must match in size and order)."""
    from ionotomo_b200.inversion.gradient import released_slices
    V, K = 1000, 4
    bounds = [V * j // K for j in range(K + 1)]
    for progress in ([100, 300, 620, 1000], [250, 500, 750, 1000], [999, 999, 1000], [1000]):
        nxt, got = 0, []
        for done in progress:
            rel = released_slices(bounds, nxt, done)
            got += rel
            nxt += len(rel)
        assert got == list(zip(bounds[:-1], bounds[1:]))


def test_apply_overlapped_host_logic_with_fake_library(monkeypatch):
    """BackProjector.apply_overlapped without a GPU: the C entry points are replaced by fakes that
    record the chunk calls; checks chunk order, the released slices and the waits."""
    from ionotomo_b200 import _lib
    from ionotomo_b200.inversion import gradient as G
    V = 6 * 5 * 4
    chunk_vox = [0, 3, 9, 20, 20, 31, 40, 47, 55, 60, 71, 80, 88, 97, 105, 111, V]   # this rank's progress table
    calls = []

    class FakeLib(object):
        def iono_backprojector_chunk_voxels(self, handle, c):
            return chunk_vox[c]

    monkeypatch.setattr(_lib, "load", lambda: FakeLib())
    monkeypatch.setattr(_lib, "to_device", lambda a, device=None: a)
    monkeypatch.setattr(_lib, "ptr", lambda t: None)
    monkeypatch.setattr(_lib, "stream_ptr", lambda: None)
    monkeypatch.setattr(_lib, "call", lambda name, *args: calls.append((name, args[4], args[5])))
    bp = object.__new__(G.BackProjector)
    bp.handle, bp.shape, bp.ray_shape = None, (6, 5, 4), (2, 2, 2)

    class Handle(object):
        waited = 0

        def wait(self):
            Handle.waited += 1

    for n_chunks in (1, 2, 4, 8, 16):
        del calls[:]
        Handle.waited = 0
        seen = []
        out = torch.zeros(6, 5, 4, dtype=torch.float64)
        coef = torch.zeros(2, 2, 2, dtype=torch.float64)

        def reducer(sl):
            seen.append((sl.data_ptr() - out.data_ptr(), sl.numel()))
            return Handle()
        got = bp.apply_overlapped(coef, out=out, n_chunks=n_chunks, reduce_slice=reducer)
        assert got is out
        step = 16 // n_chunks
        assert calls == [("iono_backprojector_apply_chunks_f64", c, c + step) for c in range(0, 16, step)]
        bounds = [V * j // n_chunks for j in range(n_chunks + 1)]
        assert seen == [(8 * a, b - a) for a, b in zip(bounds[:-1], bounds[1:])]
        assert Handle.waited == n_chunks


def _overlap_worker(rank, world, port, out):
    """Two ranks with DIFFERENT chunk-progress tables run apply_overlapped with a real asynchronous
    all_reduce (gloo): the collectives must match in order and size, or this dead-locks."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ionotomo_b200 import _lib, sharding
    from ionotomo_b200.inversion import gradient as G
    shape = (8, 7, 9)
    V = shape[0] * shape[1] * shape[2]
    rng = np.random.RandomState(100 + rank)
    cuts = np.sort(rng.randint(0, V + 1, size=15))
    chunk_vox = [0] + [int(c) for c in cuts] + [V]          # this rank's own voxel boundaries per sixteenth

    class FakeLib(object):
        def iono_backprojector_chunk_voxels(self, handle, c):
            return chunk_vox[c]

    def fake_call(name, handle, coef, scale, acc, c0, c1, stream):
        assert name == "iono_backprojector_apply_chunks_f64"
        flat = acc.reshape(-1)
        if c0 == 0:
            flat.zero_()
        flat[chunk_vox[c0]:chunk_vox[c1]] = float(rank + 1)      # "final" values of this rank's finished voxels

    _lib.load = lambda: FakeLib()
    _lib.to_device = lambda a, device=None: a
    _lib.ptr = lambda t: t
    _lib.stream_ptr = lambda: None
    _lib.call = fake_call
    bp = object.__new__(G.BackProjector)
    bp.handle, bp.shape, bp.ray_shape = None, shape, (1,)
    ok = True
    for n_chunks in (1, 2, 4, 8, 16):
        acc = torch.full(shape, -1.0, dtype=torch.float64)
        got = bp.apply_overlapped(torch.zeros(1, dtype=torch.float64), out=acc, n_chunks=n_chunks,
                                  reduce_slice=sharding.allreduce_sum_async)
        ok = ok and bool(torch.all(got == float(sum(range(1, world + 1)))))
    if rank == 0:
        out.put(ok)
    dist.destroy_process_group()


def test_overlapped_apply_collectives_match_across_ranks_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29700 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_overlap_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        if p.is_alive():
            p.terminate()
            raise AssertionError("overlapped apply dead-locked")
        assert p.exitcode == 0
    assert out.get(timeout=5) is True
