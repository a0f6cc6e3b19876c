"""Two-rank NCCL check of the sharded gradient on real GPUs (skipped with fewer than 2).
Run on a multi-GPU box: ``pytest -m gpu tests/test_gpu_multi.py``."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import ionotomo_b200 as ib
    from ionotomo_b200 import sharding
    from oracle import ionotomo_oracle as O
    from tests.problems import small_problem
    P = small_problem(42, 4, 6, 6, 32, 20, 18, 24)
    rays = O.cast_ray(P["origins"], P["directions"], P["tmax"], 32)
    i0 = 1
    g = O.forward_equation(rays, P["K_ne"], P["xvec"], P["yvec"], P["zvec"], P["m"], i0)
    rng = np.random.RandomState(0)
    dobs = g + 0.01 * rng.normal(size=g.shape)
    CdCt = np.full(g.shape, 1e-4)
    t0, t1 = sharding.time_shard(rays.shape[1], rank, world)
    tci = ib.TriCubic(P["xvec"], P["yvec"], P["zvec"], P["m"])
    rs = np.ascontiguousarray(rays[:, t0:t1])
    g_loc = ib.forward_equation(rs, P["K_ne"], tci, i0)
    ok = np.abs(g_loc - g[:, t0:t1]).max() < 1e-11 * np.abs(g).max() + 1e-9
    for bp in (None, ib.BackProjector(rs, tci)):
        grad = ib.compute_gradient(torch.as_tensor(rs).cuda(), torch.as_tensor(g_loc).cuda(),
                                   torch.as_tensor(np.ascontiguousarray(dobs[:, t0:t1])).cuda(), i0, P["K_ne"], tci,
                                   None, torch.as_tensor(np.ascontiguousarray(CdCt[:, t0:t1])).cuda(), 1., 4, 5.,
                                   reduce_fn=sharding.allreduce_sum_, backprojector=bp)
        ref = O.gradient_exact(rays, g, dobs, i0, P["K_ne"], P["xvec"], P["yvec"], P["zvec"], P["m"], CdCt)
        ok = ok and np.abs(grad.cpu().numpy() - ref).max() < 1e-10 * np.abs(ref).max()
    if os.environ.get("IONO_TEST_OVERLAP") != "1":     # experimental path, opt-in (see apply_overlapped)
        if rank == 0:
            out.put(bool(ok))
        dist.destroy_process_group()
        return
    # overlapped variant: chunked apply + asynchronous all_reduce of the finished slices
    bp = ib.BackProjector(rs, tci)
    from ionotomo_b200.inversion.gradient import adjoint_coefficients
    from ionotomo_b200.inversion.forward_equation import _ne_from_m
    coef = adjoint_coefficients(torch.as_tensor(g_loc).cuda(), torch.as_tensor(np.ascontiguousarray(dobs[:, t0:t1])).cuda(),
                                torch.as_tensor(np.ascontiguousarray(CdCt[:, t0:t1])).cuda(), i0)
    ne = _ne_from_m(tci.device_M(), P["K_ne"])
    grad2 = bp.apply_overlapped(coef, scale=ne, n_chunks=4, reduce_slice=sharding.allreduce_sum_async)
    ok = ok and np.abs(grad2.cpu().numpy() - ref).max() < 1e-10 * np.abs(ref).max()
    if rank == 0:
        out.put(bool(ok))
    dist.destroy_process_group()


def test_sharded_gradient_nccl_world2():
    import torch
    import torch.multiprocessing as mp
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29600 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert out.get(timeout=5) is True
