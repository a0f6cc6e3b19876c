#!/usr/bin/env python
"""Benchmark of the ray-integral forward model + exact adjoint (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this framework (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm

One *step* = one forward + adjoint pass over the LOFAR-like synthetic case
(BASELINE.json configs[1]: 62 stations x 200 directions x 100 times, 256x256x128 grid,
Ns = 128 samples per ray, fp64): ne = K exp(m)/TECU, TEC integrals, dTEC, misfit,
adjoint coefficients, back-projection, (allreduce across ranks), ne * acc.
Weak scaling: every rank holds a full 62x100x200 ray block (its own 200 directions of a
200*N-direction field, same 100 time steps, same 4-degree field of view and therefore the
same grid), the grid is replicated, the only collective is the allreduce of the voxel
accumulator.

Prints ONE JSON line (rank 0).  ``value`` is rays/s for the whole job with inputs resident
in HBM; ``e2e`` is the same pass through the public host-array API
(``misfit_and_gradient``) with every input copied host->device and every result copied
back inside the timed region.
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
    ap.add_argument("--overlap", type=int, default=0, choices=[0, 1, 2, 4, 8, 16],
                    help="EXPERIMENTAL: chunks of the binned apply whose allreduce overlaps the next chunk "
                         "(0 = one NCCL allreduce after the apply, the validated path)")
    ap.add_argument("--forward", default="sweep", choices=["sweep", "prepared"],
                    help="sweep: stateless ray sweep; prepared: per-geometry forward projector (36 B/sample "
                         "records assembled once, outside the timed steps, like the binned adjoint)")
    ap.add_argument("--adjoint", default="binned", choices=["binned", "scatter"],
                    help="binned: pre-assembled voxel-binned gather (default); scatter: stateless fp64 atomics")
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
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus, NT),
        "cpu_baseline": {"value": rps, "unit": "rays/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rps, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the reference is Python and cannot be imported on this image (astropy/h5py/dask absent, "
                "scipy.integrate.simps removed); this arm times oracle/ (its NumPy restatement, pinned to "
                "golden vectors from the reference's own modules) on all host cores",
    }
    print(json.dumps(line), flush=True)


def workload_config(n_gpus, nt):
    return {"workload": "LOFAR-like forward+adjoint: %d stations x %d directions x %d times per GPU, "
                        "%dx%dx%d grid, Ns=%d, fp64 (BASELINE.json configs[1..2])" % (NA, ND, nt, NX, NY, NZ, NZ),
            "rays_per_gpu": NA * nt * ND, "grid": [NX, NY, NZ], "samples_per_ray": NZ, "box": "tight",
            "sharding": "direction blocks per rank (reference antenna local), grid replicated, allreduce(acc) fp64",
            "l2_policy": "inputs larger than L2 (%.2f GB of rays per pass); no flush" % (NA * nt * ND * 4 * NZ * 8 / 1e9),
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
    from ionotomo_b200.inversion.forward_equation import ForwardProjector, _ne_from_m, tec_from_ne
    from ionotomo_b200.inversion.gradient import BackProjector, adjoint_coefficients, backproject, misfit
    from ionotomo_b200.inversion.host_stream import misfit_and_gradient

    nt = args.nt
    w = make_workload(Na=NA, Nt=nt, Nd=ND * world, nx=NX, ny=NY, nz=NZ, device="cuda",
                      d_slice=(rank * ND, (rank + 1) * ND))
    m_true = ib.TriCubic(w["xvec"], w["yvec"], w["zvec"], w["m_true"])
    m_tci = ib.TriCubic(w["xvec"], w["yvec"], w["zvec"], w["m_prior"])      # model being fitted
    grid = m_tci.grid()
    fermat = ib.Fermat(m_tci)
    rays = ib.cast_ray((w["origins"], w["directions"]), fermat, w["tmax"], w["Ns"])
    # ray generation, timed for the record (write-only kernel; not part of the steps)
    torch.cuda.synchronize()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    _lib.call("iono_cast_rays_straight_f64", _lib.ptr(w["origins"]), _lib.ptr(w["directions"]),
              rays.shape[0] * rays.shape[1] * rays.shape[2], float(w["tmax"]), int(w["Ns"]), _lib.ptr(rays),
              _lib.stream_ptr())
    c1.record()
    torch.cuda.synchronize()
    cast_ms = c0.elapsed_time(c1)
    del w["origins"], w["directions"]
    Na, Nt, Nd, _, Ns = rays.shape
    R, V = Na * Nt * Nd, NX * NY * NZ
    i0 = 0
    gen = torch.Generator(device="cuda").manual_seed(1234 + rank)
    dobs = ib.forward_equation(rays, w["K_ne"], m_true, i0, order=args.order)
    dobs = dobs + 0.01 * torch.randn(dobs.shape, dtype=torch.float64, device="cuda", generator=gen)
    CdCt = torch.full_like(dobs, 0.01 ** 2)
    K_ne = w["K_ne"]
    m_dev = m_tci.device_M()
    acc = torch.empty((NX, NY, NZ), dtype=torch.float64, device="cuda")
    ev = lambda: torch.cuda.Event(enable_timing=True)
    kt = {"fwd": [], "adj": []}
    # The ray geometry is fixed for the whole inversion (reference: rays are computed once per
    # solve, inversion_pipeline.py:195-197): assemble the voxel-binned back-projector once,
    # outside the timed steps, and report its build time and size.
    torch.cuda.synchronize()
    t_b = time.time()
    bp, bp_note = None, None
    if args.adjoint != "scatter":
        try:
            bp = BackProjector(rays, m_tci)
        except _lib.IonoError as exc:       # e.g. not enough free HBM for the assembly: stay on the GPU, use atomics
            bp_note = "binned adjoint unavailable (%s); scatter adjoint used" % str(exc)[:200]
            args.adjoint = "scatter"
    torch.cuda.synchronize()
    bp_build_s = time.time() - t_b
    fp, fp_build_s = None, None
    if args.forward == "prepared":
        t_b = time.time()
        fp = ForwardProjector(rays, m_tci)
        torch.cuda.synchronize()
        fp_build_s = time.time() - t_b

    def step(timed):
        ne = _ne_from_m(m_dev, K_ne)
        e0, e1 = ev(), ev()
        e0.record()
        if fp is not None:
            tec = fp.tec(ne)
        else:
            tec = tec_from_ne(rays, grid, ne, order=args.order, check_bounds=False)
        e1.record()
        g = torch.empty_like(tec)
        _lib.call("iono_dtec_f64", _lib.ptr(tec), Na, Nt, Nd, i0, _lib.ptr(g), _lib.stream_ptr())
        S = misfit(g, dobs, CdCt)
        coef = adjoint_coefficients(g, dobs, CdCt, i0)
        e2, e3 = ev(), ev()
        e2.record()
        if bp is not None and world > 1 and args.overlap:
            # chunked apply; the allreduce of each finished voxel slice overlaps the next chunk
            bp.apply_overlapped(coef, scale=ne, out=acc, n_chunks=args.overlap,
                                reduce_slice=sharding.allreduce_sum_async)
            e3.record()
        elif bp is not None:
            bp.apply(coef, scale=ne, out=acc)            # ne[v] * sum_ray A[v,ray] coef[ray]
            e3.record()
            sharding.allreduce_sum_(acc)
        else:
            backproject(rays, grid, coef, (NX, NY, NZ), order=args.order, check_bounds=False, out=acc)
            e3.record()
            sharding.allreduce_sum_(acc)
            _lib.call("iono_mul_f64", _lib.ptr(ne), _lib.ptr(acc), V, _lib.ptr(acc), _lib.stream_ptr())
        if timed:
            kt["fwd"].append((e0, e1))
            kt["adj"].append((e2, e3))
        return S

    def fence():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None
    t_load0 = time.time()
    for _ in range(args.warmup):
        step(False)
    fence()
    l0 = _lib.launch_count
    t_wall0 = time.time()
    start, stop = ev(), ev()
    start.record()
    for _ in range(args.steps):
        S = step(True)
    stop.record()
    fence()
    t_wall1 = time.time()
    launches = _lib.launch_count - l0
    ms = start.elapsed_time(stop) / args.steps
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    fwd_ms = float(np.mean([a.elapsed_time(b) for a, b in kt["fwd"]]))
    adj_ms = float(np.mean([a.elapsed_time(b) for a, b in kt["adj"]]))
    # the stateless scatter adjoint, timed beside it for the record (not part of the steps)
    scat_ms = None
    if bp is not None:
        coef_s = torch.randn((Na, Nt, Nd), dtype=torch.float64, device="cuda")
        tmp = torch.empty_like(acc)
        ts = []
        for i in range(3):
            a, b = ev(), ev()
            a.record()
            backproject(rays, grid, coef_s, (NX, NY, NZ), order=args.order, check_bounds=False, out=tmp)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        scat_ms = float(np.mean(ts[1:]))
        del tmp
    # likewise the stateless forward sweep when the prepared projector is the one in the steps
    sweep_ms = None
    if fp is not None:
        ne_s = _ne_from_m(m_dev, K_ne)
        ts = []
        for i in range(3):
            a, b = ev(), ev()
            a.record()
            tec_s = tec_from_ne(rays, grid, ne_s, order=args.order, check_bounds=False)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        sweep_ms = float(np.mean(ts[1:]))
        assert torch.equal(tec_s, fp.tec(ne_s)), "prepared forward differs from the stateless sweep"
    clocks = None
    if sampler:
        clocks = sampler.window(t_wall0, t_wall1)
        clocks["window"] = "timed region"
        if clocks["samples"] < 3:   # timed region shorter than the sampling period: include the warm-up
            clocks = sampler.window(t_load0, t_wall1)
            clocks["window"] = "warm-up + timed region"

    # ---- end to end through the host-array API -------------------------------------
    e2e = None
    if not args.no_e2e:
        rays_h = torch.empty(rays.shape, dtype=torch.float64, pin_memory=True)
        rays_h.copy_(rays)
        m_h = torch.empty(m_dev.shape, dtype=torch.float64, pin_memory=True).copy_(m_dev).numpy()
        dobs_h = torch.empty(dobs.shape, dtype=torch.float64, pin_memory=True).copy_(dobs).numpy()
        C_h = torch.empty(dobs.shape, dtype=torch.float64, pin_memory=True).copy_(CdCt).numpy()
        m_host_tci = ib.TriCubic(w["xvec"], w["yvec"], w["zvec"], m_h)
        red = sharding.allreduce_sum_ if world > 1 else None
        e2e_steps = max(2, min(args.steps, 4))
        for _ in range(1):
            misfit_and_gradient(rays_h, K_ne, m_host_tci, i0, dobs_h, C_h, order=args.order, reduce_fn=red,
                                copy_results=False)
        fence()
        t0 = time.time()
        for _ in range(e2e_steps):
            g_h, S_h, grad_h = misfit_and_gradient(rays_h, K_ne, m_host_tci, i0, dobs_h, C_h, order=args.order,
                                                   reduce_fn=red, copy_results=False)
        fence()
        dt = (time.time() - t0) / e2e_steps
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt[0])
        e2e = {"value": R * world / dt, "unit": "rays/s",
               "h2d_bytes_per_step": int(rays_h.numel() * 8 + m_h.nbytes + dobs_h.nbytes + C_h.nbytes),
               "d2h_bytes_per_step": int(g_h.nbytes + grad_h.nbytes + 8), "ms_per_step": dt * 1e3,
               "steps": e2e_steps,
               "api": "ionotomo_b200.inversion.host_stream.misfit_and_gradient(rays_host, K_ne, m_tci, i0, dobs, CdCt)"}
        del rays_h

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    hbm, peak_src = peaks()
    bytes_fwd = R * (4 * Ns * 8 + 8) + V * 8
    bytes_adj = R * (4 * Ns * 8 + 8) + 2 * V * 8
    fwd_name = "prepared_forward" if fp is not None else "ray_sweep_forward"
    kernels = {
        fwd_name: {"ms": fwd_ms, "algorithmic_bytes": bytes_fwd,
                   "achieved_gbs": bytes_fwd / fwd_ms / 1e6, "frac": bytes_fwd / fwd_ms / 1e6 / hbm,
                   "rays_per_s": R / fwd_ms * 1e3},
    }
    if fp is not None:
        kernels[fwd_name].update({"build_s_once_per_geometry": fp_build_s, "operator_bytes": fp.nbytes,
                                  "streamed_gbs": (fp.nbytes + V * 8 + R * 8) / fwd_ms / 1e6})
        kernels["ray_sweep_forward"] = {"ms": sweep_ms, "algorithmic_bytes": bytes_fwd,
                                        "achieved_gbs": bytes_fwd / sweep_ms / 1e6,
                                        "frac": bytes_fwd / sweep_ms / 1e6 / hbm, "rays_per_s": R / sweep_ms * 1e3,
                                        "in_step": False}
    bytes_cast = R * 4 * Ns * 8
    kernels["cast_rays"] = {"ms": cast_ms, "algorithmic_bytes": bytes_cast, "achieved_gbs": bytes_cast / cast_ms / 1e6,
                            "frac": bytes_cast / cast_ms / 1e6 / hbm, "rays_per_s": R / cast_ms * 1e3,
                            "in_step": False}
    adj_name = "binned_adjoint" if bp is not None else "ray_sweep_adjoint_scatter"
    kernels[adj_name] = {"ms": adj_ms, "algorithmic_bytes": bytes_adj, "achieved_gbs": bytes_adj / adj_ms / 1e6,
                         "frac": bytes_adj / adj_ms / 1e6 / hbm, "rays_per_s": R / adj_ms * 1e3}
    if bp is not None:
        kernels[adj_name].update({"build_s_once_per_geometry": bp_build_s, "nnz": bp.nnz,
                                  "operator_bytes": bp.nbytes,
                                  "streamed_gbs": (bp.nbytes + 2 * V * 8) / adj_ms / 1e6,
                                  "launches": "permute_coef + backproject_segments + backproject_combine"})
        kernels["ray_sweep_adjoint_scatter"] = {"ms": scat_ms, "algorithmic_bytes": bytes_adj,
                                                "achieved_gbs": bytes_adj / scat_ms / 1e6,
                                                "frac": bytes_adj / scat_ms / 1e6 / hbm,
                                                "rays_per_s": R / scat_ms * 1e3, "in_step": False}
    dom = adj_name if adj_ms >= fwd_ms else fwd_name
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(dom)
        except Exception:
            traffic = None
    line = {
        "metric": "forward+adjoint ray passes per second", "value": R * world / ms * 1e3, "unit": "rays/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(workload_config(world, nt), dx_km=w["dx_km"], dy_km=w["dy_km"], dz_km=w["dz_km"],
                       ray_order=args.order, forward=args.forward, adjoint=args.adjoint,
                       allreduce_overlap_chunks=(args.overlap if world > 1 else 0)),
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["achieved_gbs"], "peak": hbm,
                     "unit": "GB/s", "frac": kernels[dom]["frac"], "traffic": traffic, "peak_source": peak_src},
        "kernels": kernels,
        "pass_frac_of_hbm_roofline": (bytes_fwd + bytes_adj) / ms / 1e6 / hbm,
        "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        "misfit": float(S),
    }
    if bp_note:
        line["note"] = bp_note
    print(json.dumps(line), flush=True)
    if sampler:
        sampler.stop()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
