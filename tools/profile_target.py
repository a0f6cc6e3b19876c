#!/usr/bin/env python
"""Small fixed workload for ncu: a few forward and adjoint ray sweeps at the LOFAR-like
grid with NT time steps (default 25), launch knobs from the IONO_SWEEP_* environment."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ionotomo_b200 as ib
from ionotomo_b200.ionosphere.synthetic import make_workload
from ionotomo_b200.inversion.forward_equation import tec_from_ne, _ne_from_m
from ionotomo_b200.inversion.gradient import backproject

Nt = int(os.environ.get("NT", 25))
order = os.environ.get("ORDER", "time")
iso = os.environ.get("ISO")
w = make_workload(Nt=Nt, isotropic_spacing=float(iso) if iso else None)
tci = ib.TriCubic(w["xvec"], w["yvec"], w["zvec"], w["m_true"])
rays = ib.cast_ray((w["origins"], w["directions"]), ib.Fermat(tci), w["tmax"], w["Ns"])
ne = _ne_from_m(tci.device_M(), w["K_ne"])
coef = torch.randn(rays.shape[:3], dtype=torch.float64, device="cuda")
for _ in range(int(os.environ.get("REPS", 3))):
    tec = tec_from_ne(rays, tci.grid(), ne, order=order, check_bounds=False)
    acc = backproject(rays, tci.grid(), coef, tuple(ne.shape), order=order, check_bounds=False)
torch.cuda.synchronize()
print("ok", float(tec.sum()), float(acc.sum()))
