#!/usr/bin/env python
"""ncu target: build the voxel-binned back-projector for NT time steps and apply it a few times."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ionotomo_b200 as ib
from ionotomo_b200.ionosphere.synthetic import make_workload
from ionotomo_b200.inversion.forward_equation import _ne_from_m

Nt = int(os.environ.get("NT", 25))
w = make_workload(Nt=Nt)
tci = ib.TriCubic(w["xvec"], w["yvec"], w["zvec"], w["m_true"])
rays = ib.cast_ray((w["origins"], w["directions"]), ib.Fermat(tci), w["tmax"], w["Ns"])
ne = _ne_from_m(tci.device_M(), w["K_ne"])
coef = torch.randn(rays.shape[:3], dtype=torch.float64, device="cuda")
bp = ib.BackProjector(rays, tci)
for _ in range(3):
    acc = bp.apply(coef, scale=ne)
torch.cuda.synchronize()
print("ok nnz", bp.nnz, float(acc.sum()))
