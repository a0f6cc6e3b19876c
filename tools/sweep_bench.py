#!/usr/bin/env python
"""Developer tool: time the forward / adjoint ray sweeps at the LOFAR-like size for a list
of launch configurations (env knobs read by libionob200 at launch time)."""
import itertools
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ionotomo_b200 as ib
from ionotomo_b200.ionosphere.synthetic import make_workload
from ionotomo_b200.inversion.forward_equation import tec_from_ne, _ne_from_m
from ionotomo_b200.inversion.gradient import backproject

HBM = 6546.6e9


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    ev[0].record()
    for i in range(n):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(n)]
    return min(ts), sum(ts) / n


def main():
    Nt = int(os.environ.get("NT", 100))
    iso = os.environ.get("ISO")
    w = make_workload(Nt=Nt, isotropic_spacing=float(iso) if iso else None)
    tci = ib.TriCubic(w["xvec"], w["yvec"], w["zvec"], w["m_true"])
    rays = ib.cast_ray((w["origins"], w["directions"]), ib.Fermat(tci), w["tmax"], w["Ns"])
    del w["origins"], w["directions"]
    R = rays.shape[0] * rays.shape[1] * rays.shape[2]
    V = tci.nx * tci.ny * tci.nz
    ne = _ne_from_m(tci.device_M(), w["K_ne"])
    coef = torch.randn(rays.shape[:3], dtype=torch.float64, device="cuda")
    bytes_f = R * (4 * w["Ns"] * 8 + 8) + V * 8
    bytes_a = R * (4 * w["Ns"] * 8 + 8) + 2 * V * 8
    print("rays", tuple(rays.shape), "dx %.3f dy %.3f dz %.3f" % (w["dx_km"], w["dy_km"], w["dz_km"]), flush=True)
    t_cast = timeit(lambda: ib.cast_ray((rays[..., 0:3, 0].contiguous(), rays[..., 0:3, 1].contiguous() - rays[..., 0:3, 0].contiguous()), ib.Fermat(tci), w["tmax"], w["Ns"]), n=2, warm=1) if os.environ.get("CAST") else None
    configs = json.loads(os.environ.get("CONFIGS", "[]")) or [
        dict(order=o, warps=wp, stages=s, chunk=c, bulk=b)
        for o, wp, s, c, b in [("time", 16, 3, 64, 1), ("natural", 16, 3, 64, 1), ("antenna", 16, 3, 64, 1),
                               ("time", 16, 2, 128, 1), ("time", 16, 4, 64, 1), ("time", 12, 3, 64, 1),
                               ("time", 8, 3, 64, 1), ("time", 16, 3, 64, 0)]]
    if os.environ.get("BINNED", "1") != "0":
        import time
        torch.cuda.synchronize(); t0 = time.time()
        bp = ib.BackProjector(rays, tci)
        torch.cuda.synchronize(); t_build = time.time() - t0
        acc = torch.empty(tuple(ne.shape), dtype=torch.float64, device="cuda")
        tb = timeit(lambda: bp.apply(coef, scale=ne, out=acc))
        ref = backproject(rays, tci.grid(), coef, tuple(ne.shape), check_bounds=False) * ne
        err = float((acc - ref).abs().max() / ref.abs().max())
        print(json.dumps(dict(binned_adjoint_ms=round(tb[0], 3), binned_frac=round(bytes_a / (tb[0] * 1e-3) / HBM, 3),
                              nnz=bp.nnz, nnz_per_ray=round(bp.nnz / R, 1), gbytes=round(bp.nbytes / 1e9, 2),
                              actual_gbs=round((bp.nbytes + 16 * V) / (tb[0] * 1e-3) / 1e9, 1),
                              build_s=round(t_build, 3), rel_diff_vs_scatter=err)), flush=True)
        del bp, ref
    for c in configs:
        os.environ["IONO_SWEEP_WARPS"] = str(c["warps"])
        os.environ["IONO_SWEEP_STAGES"] = str(c["stages"])
        os.environ["IONO_SWEEP_CHUNK"] = str(c["chunk"])
        if c.get("bulk", 1):
            os.environ.pop("IONO_SWEEP_NO_BULK", None)
        else:
            os.environ["IONO_SWEEP_NO_BULK"] = "1"
        tf = timeit(lambda: tec_from_ne(rays, tci.grid(), ne, order=c["order"], check_bounds=False))
        ta = timeit(lambda: backproject(rays, tci.grid(), coef, tuple(ne.shape), order=c["order"], check_bounds=False))
        print(json.dumps(dict(c, fwd_ms=round(tf[0], 3), fwd_frac=round(bytes_f / (tf[0] * 1e-3) / HBM, 3),
                              adj_ms=round(ta[0], 3), adj_frac=round(bytes_a / (ta[0] * 1e-3) / HBM, 3),
                              fwd_rays_s="%.3e" % (R / (tf[0] * 1e-3)), adj_rays_s="%.3e" % (R / (ta[0] * 1e-3)))),
              flush=True)


if __name__ == "__main__":
    main()
