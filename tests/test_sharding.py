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


def _compact_worker(rank, world, port, out):
    """The sharded session's host logic without a GPU: every rank's non-empty operator rows are numbered in the
    union of all ranks' rows, the compact accumulators (+ the misfit as last element) are summed, and the
    expansion ``grad[union_voxels] = ne * sum`` equals the single-process gradient."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import ionotomo_oracle as O
    from tests.problems import small_problem
    from ionotomo_b200 import sharding
    P = small_problem(43, 4, 3, 7, 16, 12, 11, 12)
    rays = O.cast_ray(P["origins"], P["directions"], P["tmax"], 16)
    i0 = 2
    g = O.forward_equation(rays, P["K_ne"], P["xvec"], P["yvec"], P["zvec"], P["m"], i0)
    rng = np.random.RandomState(0)
    dobs = g + 0.01 * rng.normal(size=g.shape)
    CdCt = np.full(g.shape, 1e-4)
    d0, d1 = sharding.direction_shard(rays.shape[2], rank, world)
    sl = (slice(None), slice(None), slice(d0, d1))
    g_loc = O.forward_equation(rays[sl], P["K_ne"], P["xvec"], P["yvec"], P["zvec"], P["m"], i0)
    assert np.array_equal(g_loc, g[sl])                       # direction blocks keep the reference antenna local
    dd = O.weighted_residual(g_loc, dobs[sl], CdCt[sl])
    acc = O.backproject(rays[sl], P["xvec"], P["yvec"], P["zvec"], O.adjoint_ray_coefficients(dd, i0)).reshape(-1)
    # rows of this rank's operator = voxels with a structurally non-zero entry; emulate with the support of |A| 1
    support = O.backproject(rays[sl], P["xvec"], P["yvec"], P["zvec"], np.ones(g_loc.shape)).reshape(-1) != 0
    row_voxels = torch.from_numpy(np.nonzero(support)[0].astype(np.int32))
    V = acc.size
    row_dst, union_voxels, n_union = sharding.union_index(row_voxels, V)
    assert n_union >= row_voxels.numel() and bool((union_voxels[row_dst.long()] == row_voxels).all())
    assert bool((union_voxels[1:] > union_voxels[:-1]).all())
    acc_c = torch.zeros(n_union + 1, dtype=torch.float64)
    acc_c[row_dst.long()] = torch.from_numpy(acc)[row_voxels.long()]
    acc_c[n_union] = float(O.misfit(g_loc, dobs[sl], CdCt[sl]))
    sharding.allreduce_sum_(acc_c)
    grad = np.zeros(V)
    uv = union_voxels.numpy()
    grad[uv] = O.ne_from_m(P["m"], P["K_ne"]).reshape(-1)[uv] * acc_c[:n_union].numpy()
    ref = O.gradient_exact(rays, g, dobs, i0, P["K_ne"], P["xvec"], P["yvec"], P["zvec"], P["m"], CdCt).reshape(-1)
    S_ref = O.misfit(g, dobs, CdCt)
    ok = np.abs(grad - ref).max() <= 1e-12 * np.abs(ref).max() and abs(float(acc_c[n_union]) - S_ref) <= 1e-12 * S_ref
    if rank == 0:
        out.put(bool(ok))
    dist.destroy_process_group()


def test_compact_union_accumulator_gloo_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_compact_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert out.get(timeout=5) is True


def test_direction_shard_and_union_index_single_process():
    from ionotomo_b200 import sharding
    assert [sharding.direction_shard(200, r, 8) for r in range(8)] == [(25 * r, 25 * r + 25) for r in range(8)]
    row_dst, uv, n = sharding.union_index(torch.tensor([3, 4, 10], dtype=torch.int32), 16)
    assert n == 3 and uv.tolist() == [3, 4, 10] and row_dst.tolist() == [0, 1, 2]
    row_dst, uv, n = sharding.union_index(torch.zeros(0, dtype=torch.int32), 16)
    assert n == 0 and uv.numel() == 0 and row_dst.numel() == 0


def test_simpson_grid_weights_reproduce_the_reference_inner_product():
    """``solver.simpson_grid_weights`` (the optional metric of the L-BFGS driver) == triple ``simps`` over the grid
    (geometry/tri_cubic.py:61-67, bfgs_dask.py:165-167), odd and even axis lengths, non-uniform axes."""
    from oracle import ionotomo_oracle as O
    from ionotomo_b200.inversion.solver import simpson_grid_weights
    rng = np.random.RandomState(8)
    xv = np.cumsum(rng.uniform(0.5, 1.5, 9))
    yv = np.cumsum(rng.uniform(0.5, 1.5, 6))
    zv = np.linspace(0., 10., 8)
    a, b = rng.normal(size=(9, 6, 8)), rng.normal(size=(9, 6, 8))
    w = simpson_grid_weights(xv, yv, zv)
    np.testing.assert_allclose((w * a * b).sum(), O.tci_inner(xv, yv, zv, a, b), rtol=1e-12)
