#!/usr/bin/env python
"""Per-kernel timings of the hot path at the benchmark size (BASELINE configs[1], one GPU), CUDA events,
each kernel alone.  Prints one JSON object.  Environment knobs (IONO_*) select kernel variants; the
tuning experiments of profiles/ were produced with this.
    python tools/kernel_bench.py [--nt 100] [--reps 20] [--skip sweep,scatter]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import ionotomo_b200 as ib
from ionotomo_b200 import _lib
from ionotomo_b200.ionosphere.synthetic import make_workload
from ionotomo_b200.inversion.forward_equation import ForwardProjector, _ne_from_m, ne_quads_from_m, tec_from_ne, \
    tec_from_quads
from ionotomo_b200.inversion.gradient import BackProjector, backproject, residual
from ionotomo_b200.inversion.session import DeviceSession


def timeit(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return {"ms": float(np.median(ts)), "min": float(np.min(ts))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nt", type=int, default=100)
    ap.add_argument("--nd", type=int, default=200)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--skip", default="")
    ap.add_argument("--grid", default="256,256,128")
    args = ap.parse_args()
    skip = set(args.skip.split(","))
    nx, ny, nz = [int(v) for v in args.grid.split(",")]
    w = make_workload(Na=62, Nt=args.nt, Nd=args.nd, nx=nx, ny=ny, nz=nz, device="cuda")
    m_tci = ib.TriCubic(w["xvec"], w["yvec"], w["zvec"], w["m_prior"])
    grid = m_tci.grid()
    rays = ib.cast_ray((w["origins"], w["directions"]), ib.Fermat(m_tci), w["tmax"], w["Ns"])
    Na, Nt, Nd, _, Ns = rays.shape
    R, V = Na * Nt * Nd, nx * ny * nz
    m = m_tci.device_M()
    K = w["K_ne"]
    out = {"R": R, "V": V, "Ns": Ns, "env": {k: v for k, v in os.environ.items() if k.startswith("IONO_")}}
    ne = _ne_from_m(m, K)
    ne2, quads = ne_quads_from_m(m, K)
    out["ne_from_m"] = timeit(lambda: _ne_from_m(m, K), args.reps)
    out["ne_quads_from_m"] = timeit(lambda: ne_quads_from_m(m, K, ne_out=ne2, quads_out=quads), args.reps)
    tec = torch.empty((Na, Nt, Nd), dtype=torch.float64, device="cuda")
    if "sweep" not in skip:
        os.environ["IONO_FWD_LAYOUT"] = "plain"
        out["sweep_plain"] = timeit(lambda: tec_from_ne(rays, grid, ne, check_bounds=False), args.reps)
        del os.environ["IONO_FWD_LAYOUT"]
        out["sweep_quads"] = timeit(lambda: tec_from_quads(rays, grid, quads, check_bounds=False, out=tec), args.reps)
    t0 = time.time()
    fp = ForwardProjector(rays, m_tci)
    torch.cuda.synchronize()
    out["fp_build_s"] = time.time() - t0
    out["fp_bytes"] = fp.nbytes
    if "prepared" not in skip:
        os.environ["IONO_FWD_LAYOUT"] = "plain"
        out["prepared_plain"] = timeit(lambda: fp.tec(ne, out=tec), args.reps)
        del os.environ["IONO_FWD_LAYOUT"]
        out["prepared_quads"] = timeit(lambda: fp.tec_quads(quads, out=tec), args.reps)
        q2 = torch.zeros_like(quads)
        out["quads_from_m_touched"] = timeit(lambda: _lib.call("iono_forwardprojector_quads_from_m_f64", fp.handle,
                                                               _lib.ptr(m), K / 1e13, _lib.ptr(q2), _lib.stream_ptr()),
                                             args.reps)
        out["fp_records"] = int(_lib.load().iono_forwardprojector_n_records(fp.handle))
        t_ref = fp.tec_quads(quads).clone()
        out["touched_quads_same_tec"] = bool(torch.equal(fp.tec_quads(q2), t_ref))
        del q2
        # launch-shape variants of the prepared forward (environment is read at every apply)
        out["fp_factored"] = bool(fp.factored)
        for name, env in (("w32_s2", {"IONO_PREP_WARPS": "32", "IONO_PREP_STAGES": "2"}),
                          ("w32_s3", {"IONO_PREP_WARPS": "32", "IONO_PREP_STAGES": "3"}),
                          ("w32_s4", {"IONO_PREP_WARPS": "32", "IONO_PREP_STAGES": "4"}),
                          ("w24_s4", {"IONO_PREP_WARPS": "24", "IONO_PREP_STAGES": "4"})):
            os.environ.update(env)
            out["prepared_quads_" + name] = timeit(lambda: fp.tec_quads(quads, out=tec), args.reps)
            assert torch.equal(tec, t_ref)
            for k in env:
                del os.environ[k]
        # per-sample weights (36 B per sample) for comparison
        os.environ["IONO_PREP_FACTOR"] = "0"
        fp0 = ForwardProjector(rays, m_tci)
        del os.environ["IONO_PREP_FACTOR"]
        out["prepared_quads_unfactored"] = timeit(lambda: fp0.tec_quads(quads, out=tec), args.reps)
        out["factored_vs_unfactored_relerr"] = float(((tec - t_ref).abs().max() / t_ref.abs().max()).item())
        out["unfactored_equals_sweep"] = bool(torch.equal(tec, tec_from_quads(rays, grid, quads, check_bounds=False)))
        del fp0
    dobs = ib.forward_equation(rays, K, ib.TriCubic(w["xvec"], w["yvec"], w["zvec"], w["m_true"]), 0)
    dobs = dobs + 0.01 * torch.randn(dobs.shape, dtype=dobs.dtype, device=dobs.device,
                                generator=torch.Generator(device="cuda").manual_seed(1234))
    CdCt = torch.full_like(dobs, 1e-4)
    fp.tec_quads(quads, out=tec)
    bufs = {}
    out["residual"] = timeit(lambda: residual(tec, dobs, CdCt, 0, want_coef=False, want_perm=True, out=bufs), args.reps)
    g, S, _, perm = residual(tec, dobs, CdCt, 0, want_coef=True, want_perm=True)
    coef = residual(tec, dobs, CdCt, 0, want_coef=True, want_perm=False)[2]
    acc = torch.empty((nx, ny, nz), dtype=torch.float64, device="cuda")
    for runs in ("1", "0"):
        if ("runs" + runs) in skip:
            continue
        os.environ["IONO_BP_RUNS"] = runs
        t0 = time.time()
        bp = BackProjector(rays, m_tci)
        torch.cuda.synchronize()
        out["bp_runs%s_build_s" % runs] = time.time() - t0
        out["bp_runs%s_bytes" % runs] = bp.nbytes
        out["bp_nnz"] = bp.nnz
        out["apply_runs%s" % runs] = timeit(lambda: bp.apply_permuted(perm, scale=ne, out=acc), args.reps)
        if runs == "1":
            ref = acc.clone()
            out["apply_gradient"] = timeit(lambda: _lib.call("iono_backprojector_apply_gradient_f64", bp.handle,
                                                             _lib.ptr(perm), _lib.ptr(m), K / 1e13, _lib.ptr(acc), 0, 16,
                                                             _lib.stream_ptr()), args.reps)
            out["apply_gradient_relerr"] = float(((acc - ref).abs().max() / ref.abs().max()).item())
            out["bp_rows"] = int(_lib.load().iono_backprojector_n_rows(bp.handle))
            for name, env in (("w8_c3", {"IONO_BP_CTAS": "3"}), ("w6_c5", {"IONO_BP_WARPS": "6", "IONO_BP_CTAS": "5"}),
                              ("w4_c8", {"IONO_BP_WARPS": "4", "IONO_BP_CTAS": "8"})):
                os.environ.update(env)
                out["apply_runs1_" + name] = timeit(lambda: bp.apply_permuted(perm, scale=ne, out=acc), args.reps)
                assert torch.equal(acc, ref)
                for k in env:
                    del os.environ[k]
        else:
            out["runs_vs_plain_equal"] = bool(torch.equal(ref, acc))
        del bp
    os.environ.pop("IONO_BP_RUNS", None)
    if "padj" not in skip:
        # the forward projector applied transposed (+ the chain-rule finish over its voxels)
        accf = torch.zeros((nx, ny, nz), dtype=torch.float64, device="cuda")
        gradp = torch.zeros_like(accf)
        out["fp_voxels"] = fp.n_voxels
        for wv, stg in (("16", "4"), ("20", "3"), ("24", "2"), ("24", "3"), ("28", "2"), ("28", "3"), ("32", "2")):
            os.environ.update({"IONO_PADJ_WARPS": wv, "IONO_PADJ_STAGES": stg})
            out["prepared_adjoint_w%s_s%s" % (wv, stg)] = timeit(
                lambda: (fp.adjoint(perm, accf), fp.finish_gradient(accf, m, K / 1e13, gradp)), args.reps)
        del os.environ["IONO_PADJ_WARPS"], os.environ["IONO_PADJ_STAGES"]
        out["prepared_adjoint"] = timeit(
            lambda: (fp.adjoint(perm, accf), fp.finish_gradient(accf, m, K / 1e13, gradp)), args.reps)
        out["prepared_adjoint_kernel_only"] = timeit(lambda: fp.adjoint(perm, accf), args.reps)
        accf.zero_()
        if "runs1" not in skip:
            out["prepared_adjoint_vs_binned_relerr"] = float(((gradp - ref).abs().max() / ref.abs().max()).item())
        del accf, gradp
    if "scatter" not in skip:
        os.environ["IONO_ADJOINT_RUNS"] = "0"
        out["scatter_adjoint_plain"] = timeit(lambda: backproject(rays, grid, coef, (nx, ny, nz), check_bounds=False, out=acc),
                                              max(3, args.reps // 4))
        plain = acc.clone()
        del os.environ["IONO_ADJOINT_RUNS"]
        for wv in ("16", "20", "24"):
            os.environ["IONO_ADJOINT_RUNS_WARPS"] = wv
            out["scatter_adjoint_runs_w" + wv] = timeit(lambda: backproject(rays, grid, coef, (nx, ny, nz), check_bounds=False,
                                                                            out=acc), max(3, args.reps // 2))
        del os.environ["IONO_ADJOINT_RUNS_WARPS"]
        out["scatter_runs_vs_plain_relerr"] = float(((acc - plain).abs().max() / plain.abs().max()).item())
        if "runs1" not in skip:
            out["scatter_vs_binned_relerr"] = float(((acc * ne - ref).abs().max() / ref.abs().max()).item())
    if "session" not in skip:
        del fp
        for graph in (True, False):
            ses = DeviceSession(rays, K, m_tci, 0, dobs, CdCt, use_graph=graph, keep_rays=True)
            out["session_step_graph%d" % graph] = timeit(lambda: ses.misfit_and_gradient(m), args.reps)
            out["session_launches"] = ses.launches_per_call.get("step")
            del ses
    hbm = 6546.6
    bf, ba = R * (4 * Ns * 8 + 8) + V * 8, R * (4 * Ns * 8 + 8) + 2 * V * 8
    for k in list(out):
        if isinstance(out[k], dict) and "ms" in out[k]:
            if k.startswith("prepared_adjoint"):
                out[k]["frac"] = ba / out[k]["ms"] / 1e6 / hbm
            elif k.startswith(("sweep", "prepared")):
                out[k]["frac"] = bf / out[k]["ms"] / 1e6 / hbm
            elif k.startswith(("apply", "scatter_adjoint", "prepared_adjoint")):
                out[k]["frac"] = ba / out[k]["ms"] / 1e6 / hbm
            elif k.startswith("session"):
                out[k]["frac"] = (bf + ba) / out[k]["ms"] / 1e6 / hbm
    print(json.dumps(out))


if __name__ == "__main__":
    main()
