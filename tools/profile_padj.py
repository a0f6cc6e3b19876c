#!/usr/bin/env python
"""ncu target: one launch of the prepared adjoint (forward projector transposed) at the benchmark size.
    ncu --set full --import-source on -k regex:prepared_adjoint -c 1 -o gpurun_out/padj python tools/profile_padj.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import ionotomo_b200 as ib
from ionotomo_b200.ionosphere.synthetic import make_workload
from ionotomo_b200.inversion.forward_equation import ForwardProjector

w = make_workload(Na=62, Nt=int(os.environ.get("NT", "100")), Nd=200, nx=256, ny=256, nz=128, device="cuda")
m_tci = ib.TriCubic(w["xvec"], w["yvec"], w["zvec"], w["m_prior"])
rays = ib.cast_ray((w["origins"], w["directions"]), ib.Fermat(m_tci), w["tmax"], w["Ns"])
fp = ForwardProjector(rays, m_tci)
del rays
Na, Nt, Nd = fp.ray_shape
perm = torch.randn(Na * Nt * Nd, dtype=torch.float64, device="cuda")
acc = torch.zeros(fp.shape, dtype=torch.float64, device="cuda")
for _ in range(2):
    fp.adjoint(perm, acc)
torch.cuda.synchronize()
print("ok", float(acc.abs().max()))
