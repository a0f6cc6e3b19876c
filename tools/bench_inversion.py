#!/usr/bin/env python
"""BASELINE.json configs[3]: iterative tomographic inversion (L-BFGS, 50 iterations) on a
512x512x256 grid with a synthetic turbulent ionosphere, rays of the LOFAR-like case
(62 x 100 x 200, Ns = nz = 256).  Prints one JSON line with the time per iteration."""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ionotomo_b200 as ib
from ionotomo_b200.ionosphere.synthetic import make_workload
from ionotomo_b200.inversion.session import DeviceSession
from ionotomo_b200.inversion.solver import lbfgs_solve

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, nargs=3, default=[512, 512, 256])
ap.add_argument("--nt", type=int, default=100)
ap.add_argument("--iters", type=int, default=50)
ap.add_argument("--scatter", action="store_true")
ap.add_argument("--adjoint", default=None, choices=[None, "binned", "prepared"], help="default: the session's (prepared)")
ap.add_argument("--sweep", action="store_true", help="stateless forward sweep instead of the prepared forward projector")
ap.add_argument("--metric", default=None, choices=[None, "simpson"])
args = ap.parse_args()
nx, ny, nz = args.grid
w = make_workload(Nt=args.nt, nx=nx, ny=ny, nz=nz)
m_true = ib.TriCubic(w["xvec"], w["yvec"], w["zvec"], w["m_true"])
rays = ib.cast_ray((w["origins"], w["directions"]), ib.Fermat(m_true), w["tmax"], w["Ns"])
del w["origins"], w["directions"]
dobs = ib.forward_equation(rays, w["K_ne"], m_true, 0)
dobs = dobs + 0.01 * torch.randn(dobs.shape, dtype=dobs.dtype, device=dobs.device,
                                generator=torch.Generator(device="cuda").manual_seed(1234))
CdCt = torch.full_like(dobs, 1e-4)
free, total = torch.cuda.mem_get_info()
need = rays.shape[0] * rays.shape[1] * rays.shape[2] * w["Ns"] * 8 * 40
if not args.scatter and need > 0.9 * free:
    raise SystemExit("not enough free HBM for the operator assembly: need ~%.0f GB, free %.0f GB" % (need / 1e9, free / 1e9))
torch.cuda.synchronize()
t0 = time.time()
prob = DeviceSession(rays, w["K_ne"], ib.TriCubic(w["xvec"], w["yvec"], w["zvec"], w["m_prior"]), 0, dobs, CdCt,
                     forward="sweep" if args.sweep else "prepared", adjoint="scatter" if args.scatter else args.adjoint,
                     keep_rays=args.sweep or args.scatter)
del rays
torch.cuda.empty_cache()
torch.cuda.synchronize()
t_build = time.time() - t0
t0 = time.time()
m, info = lbfgs_solve(prob, w["m_prior"], n_iter=args.iters, metric=args.metric)
torch.cuda.synchronize()
dt = time.time() - t0
# share of the iteration spent in the ray kernels: the same number of forward / gradient evaluations, alone
nf, ng = info["n_forward"], info["n_gradient"]
prob.forward(m)
torch.cuda.synchronize()
t1 = time.time()
for _ in range(10):
    prob.forward(m)
torch.cuda.synchronize()
t_fwd = (time.time() - t1) / 10
t1 = time.time()
for _ in range(10):
    prob.gradient_after_forward()
torch.cuda.synchronize()
t_adj = (time.time() - t1) / 10
import numpy as np
it_s = np.array(info["iter_seconds"])
steady = float(np.median(it_s[3:])) if len(it_s) > 6 else float(np.median(it_s))     # past the one-off graph captures
evals_f, evals_g = (nf - 2) / max(1, len(it_s)), (ng - 1) / max(1, len(it_s))
ray_share = (evals_f * t_fwd + evals_g * t_adj) / steady
err0 = float((w["m_prior"] - w["m_true"]).abs().mean())
err1 = float((m - w["m_true"]).abs().mean())
print(json.dumps({
    "config": "L-BFGS inversion, %dx%dx%d grid, %d rays x %d samples" % (nx, ny, nz, prob.ray_shape[0] * prob.ray_shape[1] * prob.ray_shape[2], w["Ns"]),
    "forward_ms": t_fwd * 1e3, "adjoint_ms": t_adj * 1e3, "share_of_time_in_forward_and_adjoint": ray_share,
    "active_voxels": info["active_voxels"], "host_syncs_per_iteration": info["host_syncs_per_iteration"],
    "iterations": len(info["S"]) - 1, "seconds": dt, "s_per_iteration": dt / max(1, len(info["S"]) - 1),
    "steady_ms_per_iteration": steady * 1e3, "forwards_per_iteration": evals_f, "gradients_per_iteration": evals_g,
    "n_forward": info["n_forward"], "n_gradient": info["n_gradient"], "operator_build_s": t_build,
    "operator_gb": prob.operator_bytes / 1e9, "adjoint": prob.adjoint_kind,
    "forward": "sweep" if args.sweep else "prepared", "forward_operator_gb": (prob.fp.nbytes / 1e9) if prob.fp else 0.0,
    "misfit_first": info["S"][0], "misfit_last": info["S"][-1], "mean_abs_model_error": [err0, err1],
    "peak_hbm_gb": torch.cuda.max_memory_allocated() / 1e9}))
