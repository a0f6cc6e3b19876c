#!/usr/bin/env python
"""ncu target (round 2): one launch of each hot kernel at the benchmark size -- stateless sweep (quad layout),
prepared forward (quad layout), fused residual kernel, prepared (transposed) adjoint, run-compressed binned adjoint,
stateless run-aggregated adjoint (SCATTER=1).
    NT=100 ncu --set full --clock-control none --import-source on -k regex:"ray_sweep|prepared_forward|backproject_w|residual" \
        -o gpurun_out/prof python tools/profile_r2.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import ionotomo_b200 as ib
from ionotomo_b200.ionosphere.synthetic import make_workload
from ionotomo_b200.inversion.forward_equation import ForwardProjector, ne_quads_from_m, tec_from_quads
from ionotomo_b200.inversion.gradient import BackProjector, backproject, residual

nt = int(os.environ.get("NT", "100"))
w = make_workload(Na=62, Nt=nt, Nd=200, nx=256, ny=256, nz=128, device="cuda")
m_tci = ib.TriCubic(w["xvec"], w["yvec"], w["zvec"], w["m_prior"])
rays = ib.cast_ray((w["origins"], w["directions"]), ib.Fermat(m_tci), w["tmax"], w["Ns"])
ne, quads = ne_quads_from_m(m_tci.device_M(), w["K_ne"])
fp = ForwardProjector(rays, m_tci)
bp = BackProjector(rays, m_tci)
dobs = torch.zeros(rays.shape[:3], dtype=torch.float64, device="cuda")
C = torch.full_like(dobs, 1e-4)
from ionotomo_b200 import _lib
acc_full = torch.zeros((256, 256, 128), dtype=torch.float64, device="cuda")
grad_p = torch.zeros_like(acc_full)
for _ in range(int(os.environ.get("PASSES", "2"))):
    tec = tec_from_quads(rays, m_tci.grid(), quads, check_bounds=False)
    _lib.call("iono_forwardprojector_quads_from_m_f64", fp.handle, _lib.ptr(m_tci.device_M()), w["K_ne"] / 1e13,
              _lib.ptr(quads), _lib.stream_ptr())
    tec = fp.tec_quads(quads)
    g, S, coef, perm = residual(tec, dobs, C, 0, want_coef=True, want_perm=True)
    _lib.call("iono_backprojector_ne_rows_f64", bp.handle, _lib.ptr(m_tci.device_M()), w["K_ne"] / 1e13, _lib.ptr(ne),
              _lib.stream_ptr())
    acc = bp.apply_permuted(perm, scale=ne)
    # the forward operator transposed (session default adjoint) + the chain-rule finish over its voxels
    fp.adjoint(perm, acc_full)
    fp.finish_gradient(acc_full, m_tci.device_M(), w["K_ne"] / 1e13, grad_p)
if os.environ.get("SCATTER", "0") == "1":
    backproject(rays, m_tci.grid(), coef, tuple(ne.shape), check_bounds=False)
torch.cuda.synchronize()
print("ok", float(S))
