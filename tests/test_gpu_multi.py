"""Two-rank checks of the sharded path on real GPUs (skipped with fewer than 2): the gradient with the NCCL
allreduce hook, and the device session with both reducers -- the fused peer-memory kernel over NVLink
(``iono_peer_reduce_expand_f64``) and ``torch.distributed.all_reduce`` -- for direction and time sharding.
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
    from ionotomo_b200.inversion.session import DeviceSession
    from oracle import ionotomo_oracle as O
    from tests.problems import small_problem
    P = small_problem(42, 4, 6, 6, 32, 20, 18, 24)
    rays = O.cast_ray(P["origins"], P["directions"], P["tmax"], 32)
    i0 = 1
    g = O.forward_equation(rays, P["K_ne"], P["xvec"], P["yvec"], P["zvec"], P["m"], i0)
    rng = np.random.RandomState(0)
    dobs = g + 0.01 * rng.normal(size=g.shape)
    CdCt = np.full(g.shape, 1e-4)
    ref = O.gradient_exact(rays, g, dobs, i0, P["K_ne"], P["xvec"], P["yvec"], P["zvec"], P["m"], CdCt)
    S_ref = O.misfit(g, dobs, CdCt)
    t0, t1 = sharding.time_shard(rays.shape[1], rank, world)
    tci = ib.TriCubic(P["xvec"], P["yvec"], P["zvec"], P["m"])
    rs = np.ascontiguousarray(rays[:, t0:t1])
    g_loc = ib.forward_equation(rs, P["K_ne"], tci, i0)
    ok = np.abs(g_loc - g[:, t0:t1]).max() < 1e-11 * np.abs(g).max() + 1e-9
    why = []
    # 1. the reference-signature gradient with the allreduce hook
    for bp in (None, ib.BackProjector(rs, tci)):
        grad = ib.compute_gradient(torch.as_tensor(rs).cuda(), torch.as_tensor(g_loc).cuda(),
                                   torch.as_tensor(np.ascontiguousarray(dobs[:, t0:t1])).cuda(), i0, P["K_ne"], tci,
                                   None, torch.as_tensor(np.ascontiguousarray(CdCt[:, t0:t1])).cuda(), 1., 4, 5.,
                                   reduce_fn=sharding.allreduce_sum_, backprojector=bp)
        e = np.abs(grad.cpu().numpy() - ref).max() / np.abs(ref).max()
        if not e < 1e-10:
            ok = False
            why.append(("compute_gradient", bp is not None, e))
    # 2. the device session, both reducers, both sharding axes, graph replay with a changing model
    d0, d1 = sharding.direction_shard(rays.shape[2], rank, world)
    shards = {"time": (slice(None), slice(t0, t1)), "direction": (slice(None), slice(None), slice(d0, d1))}
    grads = {}
    for axis, sl in shards.items():
        for reducer, adjoint in (("peer", "binned"), ("nccl", "binned"), ("peer", "prepared"), ("nccl", "prepared")):
            ses = DeviceSession(np.ascontiguousarray(rays[sl]), P["K_ne"], tci, i0, np.ascontiguousarray(dobs[sl]),
                                np.ascontiguousarray(CdCt[sl]), reducer=reducer, adjoint=adjoint)
            reducer = reducer if adjoint == "binned" else reducer + "+" + adjoint
            for k in range(4):
                m = P["m"] + 0.03 * k * np.cos(np.arange(P["m"].size)).reshape(P["m"].shape)
                S, grad = ses.misfit_and_gradient(torch.as_tensor(m).cuda())
                gk = O.forward_equation(rays, P["K_ne"], P["xvec"], P["yvec"], P["zvec"], m, i0)
                rk = O.gradient_exact(rays, gk, dobs, i0, P["K_ne"], P["xvec"], P["yvec"], P["zvec"], m, CdCt)
                Sk = O.misfit(gk, dobs, CdCt)
                e = np.abs(grad.cpu().numpy() - rk).max() / np.abs(rk).max()
                es = abs(float(S) - Sk) / Sk
                if not (e < 1e-7 and es < 1e-7):      # the coefficients amplify the forward's rounding by 1/CdCt
                    ok = False
                    why.append((axis, reducer, k, e, es))
                if k == 0:
                    grads[(axis, reducer)] = grad.clone()
            dtec, S_f = ses.forward(torch.as_tensor(P["m"]).cuda())
            if not abs(float(S_f) - S_ref) <= 1e-7 * S_ref:
                ok = False
                why.append((axis, reducer, "forward S", float(S_f), S_ref))
            ses.close()
    # every rank holds the same bits (fixed summation order in the peer kernel)
    gp = grads[("direction", "peer")]
    others = [torch.empty_like(gp) for _ in range(world)]
    dist.all_gather(others, gp)
    if not all(torch.equal(o, gp) for o in others):
        ok = False
        why.append("peer reducer: ranks differ")
    if rank == 0:
        out.put((bool(ok), why))
    dist.destroy_process_group()


def test_sharded_session_and_gradient_world2():
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
        p.join(420)
        assert p.exitcode == 0
    ok, why = out.get(timeout=5)
    assert ok, why
