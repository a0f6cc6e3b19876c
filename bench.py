#!/usr/bin/env python
"""Benchmark of the ray-integral forward model + exact adjoint (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this framework (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm

One *step* = one forward + adjoint pass over the LOFAR-like synthetic case
(BASELINE.json configs[1]: 62 stations x 200 directions x 100 times, 256x256x128 grid,
Ns = 128 samples per ray, fp64) for a new model m: ne = K exp(m)/TECU, TEC integrals, dTEC,
misfit, adjoint coefficients, back-projection, sum over ranks, chain-rule factor ne -- what an
iteration of the reference's drivers evaluates on a fixed ray set (tests/test_inversion.py:30-39).

N > 1 is STRONG scaling of that fixed case (BASELINE.json configs[2]): the 200 directions are split
into N blocks (rank r holds 62 x 100 x 200/N rays; the reference antenna is local, so the forward
needs no exchange), the grid is replicated, and the per-rank back-projections are summed over NVLink
peer memory by the fused reduce/scale/expand kernel (``--reducer nccl``: torch.distributed
all_reduce instead).  ``--scaling weak`` keeps round 1's mode (every rank a full 200-direction block).

Prints ONE JSON line (rank 0).  ``value`` is rays/s for the whole job with everything resident in
HBM; ``e2e`` is the same step through the host-level session API (``HostSession.misfit_and_gradient``:
the model comes from pinned host memory and dTEC, misfit and gradient return to it inside the timed
region); ``e2e_cold`` is one pass with the 5 GB ray array itself streamed from the host.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NA, NT, ND = 62, 100, 200
NX, NY, NZ = 256, 256, 128


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nt", type=int, default=NT, help="time steps per rank (default: the named config)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--order", default=os.environ.get("IONO_BENCH_ORDER", "time"))
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="N > 1: strong = the fixed 62x100x200 case split by direction blocks (default); "
                         "weak = every rank its own full block of 200 directions")
    ap.add_argument("--reducer", default="peer", choices=["peer", "nccl"],
                    help="N > 1: peer = fused reduce/scale/expand kernel over NVLink peer memory (default); "
                         "nccl = torch.distributed.all_reduce of the compact accumulator")
    ap.add_argument("--forward", default="prepared", choices=["sweep", "prepared"],
                    help="prepared: per-geometry forward projector (28-36 B/sample records assembled once per ray "
                         "geometry, outside the timed steps; also applied transposed as the adjoint); sweep: stateless "
                         "ray sweep")
    ap.add_argument("--adjoint", default=None, choices=["binned", "prepared", "scatter"],
                    help="prepared: the forward projector's records applied transposed, run-aggregated fp64 reductions "
                         "(default with --forward prepared); binned: pre-assembled voxel-binned gather, bitwise "
                         "reproducible; scatter: stateless kernel on the rays (default with --forward sweep)")
    ap.add_argument("--e2e-adjoint", default=None, choices=["binned", "prepared", "scatter"],
                    help="adjoint of the host-level e2e legs (default: HostSession's own choice -- binned for full-grid "
                         "gradients on one GPU, whose download it pipelines behind the kernel; prepared otherwise)")
    ap.add_argument("--emulate-shard", type=int, default=0,
                    help="ONE GPU: run rank 0's share of an N-way direction split through the sharded (compact "
                         "accumulator) step without the link -- what one rank of N executes; for profiling only")
    ap.add_argument("--no-graph", action="store_true", help="launch the step kernel by kernel instead of replaying a CUDA graph")
    ap.add_argument("--no-verify", action="store_true", help="skip the single-GPU stateless recompute of S and the gradient")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------
# CPU arm: the oracle (NumPy restatement of the reference's algorithm) on all host cores
# ----------------------------------------------------------------------------------
_CPU = {}


def _cpu_init(xvec, yvec, zvec, ne, K_ne):
    _CPU.update(xvec=xvec, yvec=yvec, zvec=zvec, ne=ne, K_ne=K_ne)


def _cpu_forward(args):
    from oracle import ionotomo_oracle as O
    origins, directions, tmax, Ns = args
    rays = O.cast_ray(origins, directions, tmax, Ns)
    return O.tec(rays, _CPU["xvec"], _CPU["yvec"], _CPU["zvec"], _CPU["ne"])


def _cpu_adjoint(args):
    from oracle import ionotomo_oracle as O
    origins, directions, tmax, Ns, coef = args
    rays = O.cast_ray(origins, directions, tmax, Ns)
    return O.backproject(rays, _CPU["xvec"], _CPU["yvec"], _CPU["zvec"], coef)


def cpu_workload(n_times):
    """Host-side copy of the benchmark case restricted to the first ``n_times`` time steps
    (same generator and seed as the GPU arm)."""
    from ionotomo_b200.ionosphere.synthetic import make_workload
    from oracle import ionotomo_oracle as O
    w = make_workload(Na=NA, Nt=NT, Nd=ND, nx=NX, ny=NY, nz=NZ, device="cpu", t_slice=(0, n_times))
    m = w["m_true"].numpy()
    return dict(xvec=w["xvec"], yvec=w["yvec"], zvec=w["zvec"], ne=O.ne_from_m(m, w["K_ne"]), K_ne=w["K_ne"],
                origins=w["origins"].numpy(), directions=w["directions"].numpy(), tmax=w["tmax"], Ns=w["Ns"])


def cpu_pass(pool, cw, cores):
    """One forward + adjoint pass of the oracle over the sample, time steps spread over the
    worker processes (mirrors the reference's dask.multiprocessing fan-out,
    inversion/forward_equation.py:53-67, inversion/gradient.py:52-54)."""
    o, d = cw["origins"], cw["directions"]
    nt = o.shape[1]
    # one contiguous block of time steps per worker: each returns one TEC block / one voxel cube
    edges = np.linspace(0, nt, min(cores, nt) + 1).astype(int)
    blocks = [(a, b) for a, b in zip(edges[:-1], edges[1:]) if b > a]
    parts = [(o[:, a:b], d[:, a:b], cw["tmax"], cw["Ns"]) for a, b in blocks]
    tec = np.concatenate(pool.map(_cpu_forward, parts), axis=1)
    g = tec - tec[0]
    rng = np.random.RandomState(0)
    dd = (g - (g + 0.01 * rng.normal(size=g.shape))) / (1e-4 + 1e-15)
    coef = dd.copy()
    coef[0] -= dd.sum(0)
    acc = 0.0
    for part in pool.imap_unordered(_cpu_adjoint, [p + (coef[:, a:b],) for (a, b), p in zip(blocks, parts)]):
        acc = acc + part
    return cw["ne"] * acc


def run_cpu(steps, warmup, target_seconds=12.0):
    """Time the oracle on a bounded sample; returns (rays_per_s, cores, sample description)."""
    import multiprocessing as mp
    cores = min(os.cpu_count() or 1, 64)    # each worker returns a 67 MB cube per pass
    n_times = max(1, min(NT, cores))        # one time step (12 400 rays) per worker and pass
    cw = cpu_workload(n_times)
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_cpu_init, initargs=(cw["xvec"], cw["yvec"], cw["zvec"], cw["ne"], cw["K_ne"])) as pool:
        t = time.time()
        cpu_pass(pool, cw, cores)           # calibration / warm-up
        one = time.time() - t
        for _ in range(max(0, warmup - 1)):
            cpu_pass(pool, cw, cores)
        steps = max(1, min(steps, int(target_seconds / max(one, 1e-3)) or 1))
        t = time.time()
        for _ in range(steps):
            cpu_pass(pool, cw, cores)
        dt = (time.time() - t) / steps
    rays = NA * n_times * ND
    sample = "%d of %d time steps (%d rays) per step, %d timed steps, NumPy oracle, %d processes" % (
        n_times, NT, rays, steps, cores)
    return rays / dt, cores, sample, dt


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rps, cores, sample, dt = run_cpu(args.steps, args.warmup, target_seconds=60.0)
    line = {
        "impl": "reference", "metric": "forward+adjoint ray passes per second", "value": rps, "unit": "rays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": args.scaling if args.gpus > 1 else "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus, NT, nd_rank0(args.gpus, args.scaling), args.scaling),
        "cpu_baseline": {"value": rps, "unit": "rays/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rps, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the reference is Python and cannot be imported on this image (astropy/h5py/dask absent, "
                "scipy.integrate.simps removed); this arm times oracle/ (its NumPy restatement, pinned to "
                "golden vectors from the reference's own modules) on all host cores",
    }
    print(json.dumps(line), flush=True)


def nd_rank0(n_gpus, scaling):
    """Directions held by rank 0 (both arms name the same workload)."""
    if n_gpus > 1 and scaling == "strong":
        base, rem = divmod(ND, n_gpus)
        return base + (1 if rem else 0)
    return ND


def workload_config(n_gpus, nt, nd_rank, scaling):
    if n_gpus > 1 and scaling == "strong":
        shard = "the %d directions split into %d blocks of %d (strong scaling of the fixed case)" % (ND, n_gpus, nd_rank)
    elif n_gpus > 1:
        shard = "every rank its own block of %d directions of a %d-direction field (weak scaling)" % (ND, ND * n_gpus)
    else:
        shard = "one GPU"
    return {"workload": "LOFAR-like forward+adjoint: %d stations x %d directions x %d times, %dx%dx%d grid, Ns=%d, "
                        "fp64 (BASELINE.json configs[1..2])" % (NA, ND, nt, NX, NY, NZ, NZ),
            "rays_total": NA * nt * (ND if (n_gpus == 1 or scaling == "strong") else ND * n_gpus),
            "rays_per_gpu": NA * nt * nd_rank, "grid": [NX, NY, NZ], "samples_per_ray": NZ, "box": "tight",
            "sharding": shard + "; reference antenna local, grid replicated",
            "l2_policy": "inputs larger than L2 (>= %.2f GB of operator stream per GPU and pass); no flush"
                         % (NA * nt * nd_rank * 2 * 4 * NZ * 8 / 1e9),
            "seed": 1234, "i0": 0, "tmax_km": 1000.0}


# ----------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------
class ClockSampler(object):
    """Samples SM clock and throttle reasons every ~5 ms through NVML (pynvml) in a thread;
    falls back to polling nvidia-smi."""
    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index):
        self.rows, self.stop_flag, self.proc, self.max_mhz = [], False, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = None
            try:   # the NVML index can differ from the CUDA index: match by UUID
                import torch
                uuid = str(torch.cuda.get_device_properties(index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
        except Exception:
            self.nv = None
            self.thread = threading.Thread(target=self._poll_smi, args=(index,), daemon=True)
        self.thread.start()

    def _poll_nvml(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.rows.append((time.time(), mhz, mask))
            except Exception:
                pass
            time.sleep(0.005)

    def _poll_smi(self, index):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.active"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                self.max_mhz = float(f[1])
                self.rows.append((time.time(), float(f[0]), int(f[2], 16)))
            except Exception:
                time.sleep(0.05)

    def window(self, t0, t1):
        sel = [(m, k) for ts, m, k in self.rows if t0 <= ts <= t1]
        reasons = sorted(n for n, bit in self.REASONS.items() if any(k & bit for _, k in sel))
        return {"sm_mhz": float(np.median([m for m, _ in sel])) if sel else None, "sm_max_mhz": self.max_mhz,
                "reasons": reasons, "samples": len(sel)}

    def stop(self):
        self.stop_flag = True


# ----------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------
def ev():
    import torch
    return torch.cuda.Event(enable_timing=True)


def main():
    args = parse()
    if args.impl == "reference":
        return reference_arm(args)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # before CUDA is initialised in this process (the workers are forked)
        rps, cores, sample, _ = run_cpu(3, 1)
        cpu = {"value": rps, "unit": "rays/s", "cores": cores, "kind": "port", "sample": sample}

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a CUDA device; there is no CPU fallback"
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import ionotomo_b200 as ib
    from ionotomo_b200 import _lib, sharding
    from ionotomo_b200.ionosphere.synthetic import make_workload
    from ionotomo_b200.inversion.forward_equation import _ne_from_m, ne_quads_from_m, tec_from_quads
    from ionotomo_b200.inversion.gradient import backproject, residual
    from ionotomo_b200.inversion.session import DeviceSession
    from ionotomo_b200.inversion.host_stream import HostSession, misfit_and_gradient

    nt = args.nt
    strong = world == 1 or args.scaling == "strong"
    if args.emulate_shard:
        assert world == 1
        nd_total = ND
        d0, d1 = sharding.direction_shard(ND, 0, args.emulate_shard)
        args.no_verify = args.no_e2e = True
    elif strong:
        nd_total = ND
        d0, d1 = sharding.direction_shard(ND, rank, world)
    else:
        nd_total = ND * world
        d0, d1 = rank * ND, (rank + 1) * ND
    w = make_workload(Na=NA, Nt=nt, Nd=nd_total, nx=NX, ny=NY, nz=NZ, device="cuda", d_slice=(d0, d1))
    m_true = ib.TriCubic(w["xvec"], w["yvec"], w["zvec"], w["m_true"])
    m_tci = ib.TriCubic(w["xvec"], w["yvec"], w["zvec"], w["m_prior"])      # model being fitted
    grid = m_tci.grid()
    rays = ib.cast_ray((w["origins"], w["directions"]), ib.Fermat(m_tci), w["tmax"], w["Ns"])
    # ray generation, timed for the record (write-only kernel; once per geometry, not part of the steps)
    torch.cuda.synchronize()
    c0, c1 = ev(), ev()
    c0.record()
    _lib.call("iono_cast_rays_straight_f64", _lib.ptr(w["origins"]), _lib.ptr(w["directions"]),
              rays.shape[0] * rays.shape[1] * rays.shape[2], float(w["tmax"]), int(w["Ns"]), _lib.ptr(rays),
              _lib.stream_ptr())
    c1.record()
    torch.cuda.synchronize()
    cast_ms = c0.elapsed_time(c1)
    origins_h, directions_h = w["origins"].cpu(), w["directions"].cpu()
    del w["origins"], w["directions"]
    Na, Nt, Nd, _, Ns = rays.shape
    R, V = Na * Nt * Nd, NX * NY * NZ
    R_total = Na * Nt * nd_total
    i0 = 0
    K_ne = w["K_ne"]
    # observations: forward of the true model + N(0, 0.01 TECU); the noise field is drawn for the WHOLE case with
    # one seed on every rank and sliced, so that a single-GPU recompute sees the same data
    gen = torch.Generator(device="cuda").manual_seed(1234)
    noise = 0.01 * torch.randn((Na, Nt, nd_total), dtype=torch.float64, device="cuda", generator=gen)
    dobs = ib.forward_equation(rays, K_ne, m_true, i0, order=args.order) + noise[:, :, d0:d1]
    CdCt = torch.full_like(dobs, 0.01 ** 2)
    m_dev = m_tci.device_M()

    # The ray geometry is fixed for the whole inversion (reference: rays are computed once per solve,
    # inversion_pipeline.py:195-197): the session assembles both prepared operators once, outside the timed steps.
    torch.cuda.synchronize()
    t_b = time.time()
    note = None
    try:
        ses = DeviceSession(rays, K_ne, m_tci, i0, dobs, CdCt, forward=args.forward, adjoint=args.adjoint,
                            order=args.order, use_graph=not args.no_graph, keep_rays=True, reducer=args.reducer,
                            compact=bool(args.emulate_shard))
    except _lib.IonoError as exc:           # e.g. not enough free HBM for the assembly: stay on the GPU, stateless kernels
        if world > 1:
            raise
        note = "prepared operators unavailable (%s); stateless kernels used" % str(exc)[:200]
        args.forward, args.adjoint = "sweep", "scatter"
        ses = DeviceSession(rays, K_ne, m_tci, i0, dobs, CdCt, forward="sweep", adjoint="scatter", order=args.order,
                            use_graph=not args.no_graph)
    torch.cuda.synchronize()
    build_s = time.time() - t_b
    adjoint_used = ses.adjoint_kind

    def fence():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None
    t_load0 = time.time()
    # the model lives in the session's buffer and is updated there in place by the optimiser (lbfgs_solve's trial
    # steps scatter into ses.m): a step does not copy the 67 MB grid
    ses.m.copy_(m_dev)
    for _ in range(max(args.warmup, 1)):     # the first call runs eagerly and captures the graph
        ses.misfit_and_gradient()
    fence()
    l0 = _lib.launch_count
    t_wall0 = time.time()
    start, stop = ev(), ev()
    start.record()
    for _ in range(args.steps):
        S, grad = ses.misfit_and_gradient()
    stop.record()
    fence()
    t_wall1 = time.time()
    launches = _lib.launch_count - l0
    ms = start.elapsed_time(stop) / args.steps
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    ms_ranks = None
    if world > 1:
        tl = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(tl, t)
        ms_ranks = [float(x[0]) for x in tl]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    S_val = float(S)
    clocks = None
    if sampler:
        clocks = sampler.window(t_wall0, t_wall1)
        clocks["window"] = "timed region"
        if clocks["samples"] < 3:   # timed region shorter than the sampling period: include the warm-up
            clocks = sampler.window(t_load0, t_wall1)
            clocks["window"] = "warm-up + timed region"

    # ---- per-component durations: each call of the step captured REP times into its own CUDA graph and one
    # replay timed with CUDA events on the launching stream (device time without host launch gaps) ----------
    REP = 8
    kms = {}
    phases = None

    def timed(name, fn):
        fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(REP):
                fn()
        ts = []
        for _ in range(3):
            if world > 1:
                dist.barrier()
                torch.cuda.synchronize()
            a, b = ev(), ev()
            a.record()
            g.replay()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) / REP)
        kms[name] = float(np.median(ts))

    ses.m.copy_(m_dev)
    if ses.fp is not None:
        timed("quads_from_m", lambda: _lib.call("iono_forwardprojector_quads_from_m_f64", ses.fp.handle,
                                               _lib.ptr(ses.m), K_ne / 1e13, _lib.ptr(ses.quads), _lib.stream_ptr()))
        timed("prepared_forward", lambda: ses.fp.tec_quads(ses.quads, out=ses.tec))
    else:
        timed("quads_from_m", lambda: ne_quads_from_m(ses.m, K_ne, ne_out=ses.ne, quads_out=ses.quads,
                                                      want_ne=ses.ne is not None))
        timed("ray_sweep_forward", lambda: tec_from_quads(ses.rays, grid, ses.quads, order=args.order,
                                                          check_bounds=False, out=ses.tec, oob=ses.oob))
    timed("residual", ses._enqueue_residual)
    timed({"binned": "binned_adjoint", "prepared": "prepared_adjoint", "scatter": "ray_sweep_adjoint_scatter"}[ses.adjoint_kind],
          ses._enqueue_adjoint)
    if ses.sharded:
        if world > 1 and args.reducer == "nccl":
            ts = []
            for _ in range(5):
                dist.barrier()
                torch.cuda.synchronize()
                a, b = ev(), ev()
                a.record()
                dist.all_reduce(ses.acc_c)
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            kms["allreduce_nccl"] = float(np.median(ts))
        timed("peer_reduce_expand" if (world > 1 and args.reducer == "peer") else "expand", ses._enqueue_reduce)
        if world > 1 and args.reducer == "peer":
            phases = ses.reducer.phase_times_us()
    ses.misfit_and_gradient()          # leave the session's buffers in the state of a whole step
    torch.cuda.synchronize()

    # ---- verification: S and the gradient against a single-GPU recompute with the STATELESS kernels ----------
    verify = None
    if not args.no_verify and strong:
        grad_dev = grad.clone()
        if world > 1:
            dist.barrier()
        if rank == 0:
            if world > 1:
                wf = make_workload(Na=NA, Nt=nt, Nd=ND, nx=NX, ny=NY, nz=NZ, device="cuda")
                rays_f = ib.cast_ray((wf["origins"], wf["directions"]), ib.Fermat(m_tci), wf["tmax"], wf["Ns"])
                dobs_f = ib.forward_equation(rays_f, K_ne, m_true, i0, order=args.order) + noise
                del wf
            else:
                rays_f, dobs_f = rays, dobs
            ne_f, quads_f = ne_quads_from_m(m_dev, K_ne)
            tec_f = tec_from_quads(rays_f, grid, quads_f, order=args.order, check_bounds=True)
            g_f, S_f, coef_f, _ = residual(tec_f, dobs_f, torch.full_like(dobs_f, 0.01 ** 2), i0, want_coef=True)
            acc_f = backproject(rays_f, grid, coef_f, (NX, NY, NZ), order=args.order, check_bounds=False)
            grad_f = acc_f * ne_f
            gerr = float((grad_dev - grad_f).abs().max() / grad_f.abs().max())
            serr = abs(S_val - float(S_f)) / abs(float(S_f))
            verify = {"against": "single-GPU stateless sweep + scatter adjoint over all %d rays" % (Na * Nt * ND),
                      "grad_max_rel_err": gerr, "misfit_rel_err": serr, "tolerance": 1e-9,
                      "ok": bool(gerr < 1e-9 and serr < 1e-9)}
            assert verify["ok"], "sharded/prepared step disagrees with the stateless single-GPU recompute: %r" % verify
            del rays_f, dobs_f, acc_f, grad_f, ne_f, quads_f, tec_f
        if world > 1:
            dist.barrier()
        del grad_dev

    # ---- end to end through the host-level session API ------------------------------------------
    e2e, e2e_active, e2e_cold = None, None, None
    if not args.no_e2e:
        ses.close()
        del ses
        torch.cuda.empty_cache()
        dobs_h, C_h = dobs.cpu().numpy(), CdCt.cpu().numpy()

        def time_host_session(active):
            hs = HostSession(None, K_ne, m_tci, i0, dobs_h, C_h, origins=origins_h, directions=directions_h,
                             tmax=w["tmax"], Ns=w["Ns"], forward=args.forward, adjoint=args.e2e_adjoint, order=args.order,
                             use_graph=not args.no_graph, reducer=args.reducer, active_only=active)
            if not active:
                hs.m_host.copy_(m_dev)
            n_steps = max(3, min(args.steps, 20))
            for _ in range(3):
                hs.misfit_and_gradient()
            fence()
            t0 = time.time()
            for _ in range(n_steps):
                g_h, S_h, grad_h = hs.misfit_and_gradient()
            fence()
            dt = (time.time() - t0) / n_steps
            tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt[0])
            res = {"value": R_total / dt, "unit": "rays/s", "h2d_bytes_per_step": int(hs.h2d_bytes_per_call),
                   "d2h_bytes_per_step": int(hs.d2h_bytes_per_call), "ms_per_step": dt * 1e3, "steps": n_steps,
                   "misfit": S_h, "adjoint": hs.session.adjoint_kind}
            hs.close()
            del hs
            torch.cuda.empty_cache()
            return res

        e2e = time_host_session(False)
        e2e["api"] = ("ionotomo_b200.inversion.host_stream.HostSession(...).misfit_and_gradient(m_host): the model comes "
                      "from pinned host memory, dTEC (local rays), misfit and gradient return to pinned host memory; "
                      "geometry, data and operators stay resident (the reference computes its rays once per solve)")
        e2e["multi_gpu"] = ("the host program is rank 0's: its model is broadcast to the other GPUs over NVLink, the "
                            "gradient returns on rank 0 only; byte counts are rank 0's")
        if args.forward == "prepared" or args.e2e_adjoint == "binned":
            e2e_active = time_host_session(True)
            e2e_active["api"] = ("HostSession(..., active_only=True): model and gradient as vectors over the voxels some "
                                 "ray touches (the gradient is zero elsewhere)")
        # (stateless kernels only: no operator knows the active voxels)
        if world == 1:
            # cold call: the materialised 5 GB ray array itself comes from the host, time block by time block
            rays_h = torch.empty(rays.shape, dtype=torch.float64, pin_memory=True)
            rays_h.copy_(rays)
            m_host_tci = ib.TriCubic(w["xvec"], w["yvec"], w["zvec"], m_dev.cpu().numpy())
            misfit_and_gradient(rays_h, K_ne, m_host_tci, i0, dobs_h, C_h, order=args.order, copy_results=False)
            fence()
            t0 = time.time()
            for _ in range(2):
                g_c, S_c, grad_c = misfit_and_gradient(rays_h, K_ne, m_host_tci, i0, dobs_h, C_h, order=args.order,
                                                       copy_results=False)
            fence()
            dtc = (time.time() - t0) / 2
            e2e_cold = {"value": R_total / dtc, "unit": "rays/s", "ms_per_step": dtc * 1e3,
                        "h2d_bytes_per_step": int(rays_h.numel() * 8 + V * 8 + 2 * dobs_h.nbytes),
                        "d2h_bytes_per_step": int(g_c.nbytes + grad_c.nbytes + 8),
                        "api": "ionotomo_b200.inversion.host_stream.misfit_and_gradient(rays_host, ...): stateless "
                               "kernels, rays streamed over PCIe"}
            del rays_h

    kms_ranks = None
    if world > 1:
        names = sorted(kms)
        tk = torch.tensor([kms[n] for n in names], dtype=torch.float64, device="cuda")
        tl = [torch.zeros_like(tk) for _ in range(world)]
        dist.all_gather(tl, tk)
        kms_ranks = {n: [round(float(x[i]), 4) for x in tl] for i, n in enumerate(names)}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    hbm, peak_src = peaks()
    Rr = R      # rays this rank's kernels processed
    bytes_fwd = Rr * (4 * Ns * 8 + 8) + V * 8
    bytes_adj = Rr * (4 * Ns * 8 + 8) + 2 * V * 8
    kernels = {}
    for name, msk in kms.items():
        k = {"ms": msk}
        if name in ("prepared_forward", "ray_sweep_forward"):
            k.update(algorithmic_bytes=bytes_fwd, achieved_gbs=bytes_fwd / msk / 1e6, frac=bytes_fwd / msk / 1e6 / hbm,
                     rays_per_s=Rr / msk * 1e3)
        elif name in ("binned_adjoint", "prepared_adjoint", "ray_sweep_adjoint_scatter"):
            k.update(algorithmic_bytes=bytes_adj, achieved_gbs=bytes_adj / msk / 1e6, frac=bytes_adj / msk / 1e6 / hbm,
                     rays_per_s=Rr / msk * 1e3)
        if name == "peer_reduce_expand" and phases:
            k["phases_us_rank0_last_call"] = phases
        kernels[name] = k
    kernels["cast_rays"] = {"ms": cast_ms, "algorithmic_bytes": Rr * 4 * Ns * 8, "achieved_gbs": Rr * 4 * Ns * 8 / cast_ms / 1e6,
                            "frac": Rr * 4 * Ns * 8 / cast_ms / 1e6 / hbm, "in_step": False}
    fwd_name = "prepared_forward" if "prepared_forward" in kernels else "ray_sweep_forward"
    adj_name = [n for n in ("binned_adjoint", "prepared_adjoint", "ray_sweep_adjoint_scatter") if n in kernels][0]
    dom = adj_name if kernels[adj_name]["ms"] >= kernels[fwd_name]["ms"] else fwd_name
    traffic, traffic_note = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and world == 1:
        try:
            tj = json.load(open(tpath))
            ent = tj.get(dom)
            # only a capture of the SAME kernel at the SAME size counts
            if isinstance(ent, dict) and ent.get("rays") == Rr and ent.get("kernel"):
                traffic, traffic_note = ent.get("dram_bytes"), "ncu --set full, %s (%s)" % (ent["kernel"], ent.get("source"))
        except Exception:
            traffic = None
    line = {
        "metric": "forward+adjoint ray passes per second", "value": R_total / ms * 1e3, "unit": "rays/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "ms_per_step_ranks": ms_ranks, "kernel_ms_ranks": kms_ranks,
        "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": dict(workload_config(world, nt, Nd, args.scaling), dx_km=w["dx_km"], dy_km=w["dy_km"], dz_km=w["dz_km"],
                       ray_order=args.order, forward=args.forward, adjoint=adjoint_used,
                       cuda_graph=not args.no_graph, reducer=(args.reducer if world > 1 else None),
                       model="resident in the session's buffer, updated there in place by the optimiser "
                             "(no per-step copy of the grid); the e2e legs upload it from the host every step"),
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["achieved_gbs"], "peak": hbm,
                     "unit": "GB/s", "frac": kernels[dom]["frac"], "traffic": traffic, "traffic_source": traffic_note,
                     "peak_source": peak_src},
        "kernels": kernels,
        "kernel_timing": "separate pass: each call of the step captured 8x into its own CUDA graph, one replay timed "
                         "with CUDA events (the timed region replays the whole step as one CUDA graph)",
        "pass_frac_of_hbm_roofline": (bytes_fwd + bytes_adj) / ms / 1e6 / hbm,     # per GPU: this rank's rays
        "setup_once_per_geometry": {"ray_generation_ms": cast_ms, "operators_build_s": build_s,
                                    "amortised_ms_per_step_over_50_iterations": ms + (build_s * 1e3 + cast_ms) / 50.0},
        "cpu_baseline": cpu, "e2e": e2e, "e2e_active_voxels": e2e_active, "e2e_cold": e2e_cold, "gpu_launches": launches, "clocks": clocks,
        "misfit": S_val, "verify": verify,
    }
    if note:
        line["note"] = note
    print(json.dumps(line), flush=True)
    if sampler:
        sampler.stop()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
