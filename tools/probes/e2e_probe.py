#!/usr/bin/env python
"""Probe: where does the host-array pass (misfit_and_gradient) spend its time?"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import ionotomo_b200 as ib
from ionotomo_b200.ionosphere.synthetic import make_workload
from ionotomo_b200.inversion.host_stream import misfit_and_gradient
w = make_workload()
tci = ib.TriCubic(w["xvec"], w["yvec"], w["zvec"], w["m_true"])
rays = ib.cast_ray((w["origins"], w["directions"]), ib.Fermat(tci), w["tmax"], w["Ns"])
dobs = ib.forward_equation(rays, w["K_ne"], tci, 0)
rays_h = torch.empty(rays.shape, dtype=torch.float64, pin_memory=True); rays_h.copy_(rays)
pin = lambda t: torch.empty(t.shape, dtype=torch.float64, pin_memory=True).copy_(t).numpy()
m_h, dobs_h = pin(w["m_prior"]), pin(dobs)
C_h = pin(torch.full_like(dobs, 1e-4))
del rays
mt = ib.TriCubic(w["xvec"], w["yvec"], w["zvec"], m_h)
for bt in (None, 10):
    misfit_and_gradient(rays_h, w["K_ne"], mt, 0, dobs_h, C_h, block_times=bt, copy_results=False)
    torch.cuda.synchronize(); t0 = time.time()
    for _ in range(3):
        misfit_and_gradient(rays_h, w["K_ne"], mt, 0, dobs_h, C_h, block_times=bt, copy_results=False)
    torch.cuda.synchronize()
    tm = {}
    misfit_and_gradient(rays_h, w["K_ne"], mt, 0, dobs_h, C_h, block_times=bt, timings=tm)
    print({k: round(v, 1) for k, v in tm.items()})
    print("block_times", bt, "%.1f ms per pass" % ((time.time() - t0) / 3 * 1e3), flush=True)
import cProfile, pstats
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    misfit_and_gradient(rays_h, w["K_ne"], mt, 0, dobs_h, C_h, copy_results=False)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
