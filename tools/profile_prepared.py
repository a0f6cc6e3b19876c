#!/usr/bin/env python
"""ncu target: the two per-geometry operators of an inversion step at NT time steps -- prepared forward
(ForwardProjector) and voxel-binned adjoint (BackProjector; IONO_BP_RUNS=1 for the run-compressed form) --
built once, applied a few times.

    NT=100 ncu --set full --clock-control none --import-source on -k regex:"prepared_forward|backproject_w" \
        -c 6 -o gpurun_out/prof_prepared python tools/profile_prepared.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ionotomo_b200 as ib
from ionotomo_b200.ionosphere.synthetic import make_workload
from ionotomo_b200.inversion.forward_equation import _ne_from_m, tec_from_ne

Nt = int(os.environ.get("NT", 25))
w = make_workload(Nt=Nt)
tci = ib.TriCubic(w["xvec"], w["yvec"], w["zvec"], w["m_true"])
rays = ib.cast_ray((w["origins"], w["directions"]), ib.Fermat(tci), w["tmax"], w["Ns"])
ne = _ne_from_m(tci.device_M(), w["K_ne"])
coef = torch.randn(rays.shape[:3], dtype=torch.float64, device="cuda")
fp = ib.ForwardProjector(rays, tci)
bp = ib.BackProjector(rays, tci)
for _ in range(int(os.environ.get("REPS", 3))):
    tec = fp.tec(ne)
    acc = bp.apply(coef, scale=ne)
torch.cuda.synchronize()
same = torch.equal(tec, tec_from_ne(rays, tci.grid(), ne, check_bounds=False))
print("ok fp %.2f GB, bp %.2f GB (nnz %d), prepared == sweep: %s" % (fp.nbytes / 1e9, bp.nbytes / 1e9, bp.nnz, same))
