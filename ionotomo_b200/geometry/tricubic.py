"""True tricubic (C1, Lekien-Marsden) interpolation of a grid field and rays bent by its gradient -- BASELINE
config 5 "where the reference implements it": the reference's notebooks
(``notebooks/TricubicInterpolation.ipynb[cell 0]:138-299,1192-1257``, ``notebooks/DeriveTricubic.ipynb[cell 0]:87-141``,
``notebooks/FermatClass.ipynb[cell 0]:60-96``); the shipped package interpolates trilinearly and never bends a ray.
See ``csrc/iono_tricubic.cuh`` for the conventions (derivatives by 4th-order central differences over the local
spacing, scaled by the cell size; RK4 in z)."""
import ctypes

import numpy as np
import torch

from .. import _lib


class TricubicField(object):
    """The 8 derivative grids of ``tci.M`` on the device + point-wise evaluation with gradient."""

    def __init__(self, tci, field=None):
        self.tci = tci
        self.grid = tci.grid()
        f = tci.device_M() if field is None else _lib.to_device(field).reshape(tci.nx, tci.ny, tci.nz)
        self.derivs = torch.empty((8,) + tuple(f.shape), dtype=torch.float64, device=f.device)
        _lib.call("iono_tricubic_derivs_f64", self.grid.handle, _lib.ptr(f.contiguous()), _lib.ptr(self.derivs),
                  _lib.stream_ptr())

    def interp(self, x, y, z, grad=False, bounds_error=True):
        """``f(x,y,z)`` (and the physical gradient ``(...,3)`` with ``grad=True``); NumPy in -> NumPy out."""
        want_numpy = not isinstance(x, torch.Tensor)
        xd, yd, zd = (_lib.to_device(v).reshape(-1) for v in (x, y, z))
        shape = tuple(np.shape(x)) if want_numpy else tuple(x.shape)
        out = torch.empty_like(xd)
        g = torch.empty((xd.numel(), 3), dtype=torch.float64, device=xd.device) if grad else None
        oob = torch.zeros(1, dtype=torch.int64, device=xd.device)
        _lib.call("iono_tricubic_interp_f64", self.grid.handle, _lib.ptr(self.derivs), _lib.ptr(xd), _lib.ptr(yd),
                  _lib.ptr(zd), xd.numel(), _lib.ptr(out), _lib.ptr(g) if g is not None else None,
                  ctypes.c_void_p(oob.data_ptr()), _lib.stream_ptr())
        if bounds_error and int(oob.item()) != 0:
            raise ValueError("One of the requested xi is out of bounds (%d points outside the grid)" % int(oob.item()))
        out = out.reshape(shape)
        if grad:
            g = g.reshape(shape + (3,))
            return (out.cpu().numpy(), g.cpu().numpy()) if want_numpy else (out, g)
        return out.cpu().numpy() if want_numpy else out


def bent_rays(ne_tci, origins, directions, tmax, N, frequency=120e6, substeps=4):
    """Rays through the refractive index ``n = sqrt(1 - 8.98^2 ne / nu^2)`` of ``ne_tci`` (``Fermat.ne2n``,
    fermat.py:36-46), bent by its tricubic gradient: ``(..., 3)`` origins/directions -> ``(..., 4, N)`` rows
    x, y, z, s at ``z = linspace(z0, tmax, N)`` (the layout of ``cast_ray``).  Raises ``ValueError`` if a ray
    leaves the grid."""
    want_numpy = not isinstance(origins, torch.Tensor)
    o, d = _lib.to_device(origins), _lib.to_device(directions)
    assert o.shape == d.shape and o.shape[-1] == 3
    lead = tuple(o.shape[:-1])
    nrays = int(np.prod(lead)) if lead else 1
    ne = ne_tci.device_M()
    n_field = torch.empty_like(ne)
    _lib.call("iono_ne_to_refractive_index_f64", _lib.ptr(ne), ne.numel(), float(frequency), _lib.ptr(n_field),
              _lib.stream_ptr())
    fld = TricubicField(ne_tci, n_field)
    rays = torch.empty(lead + (4, int(N)), dtype=torch.float64, device=o.device)
    oob = torch.zeros(1, dtype=torch.int64, device=o.device)
    _lib.call("iono_bent_rays_f64", fld.grid.handle, _lib.ptr(fld.derivs), _lib.ptr(o), _lib.ptr(d), nrays, float(tmax),
              int(N), int(substeps), _lib.ptr(rays), ctypes.c_void_p(oob.data_ptr()), _lib.stream_ptr())
    if int(oob.item()) != 0:
        raise ValueError("%d rays left the grid" % int(oob.item()))
    return rays.cpu().numpy() if want_numpy else rays
