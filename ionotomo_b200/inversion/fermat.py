"""``Fermat``: ray integration through the model frame.

Mirrors ``inversion/fermat.py:5-174`` of the reference.  The shipped code is
straight-ray in effect (``euler_ode`` hard-codes grad n = 0, ``fermat.py:53-55``;
every production caller passes ``straight_line_approx=True``), with the
z coordinate as independent variable (``type='z'``): the closed form
``x = x0 + (px/pz)(z - z0)``, ``s = (z - z0)/pz`` on ``z = linspace(z0, tmax, N)``
is evaluated on the GPU by ``iono_cast_rays_straight_f64``.  ``type='s'`` (arc length as the independent
variable, ``fermat.py:74-82``): ``s = linspace(0, tmax, N)``, position ``= origin + p s``
(``iono_cast_rays_arclength_f64``).
"""
import numpy as np
import torch

from .. import _lib


class Fermat(object):
    def __init__(self, ne_tci=None, frequency=120e6, type='z', straight_line_approx=True):
        if type not in ('z', 's'):
            raise ValueError("type must be 'z' (z as the independent variable) or 's' (arc length)")
        if type == 's' and not straight_line_approx:
            # fermat.py:74-82 with the interpolated index: dx/ds = p/n along a fixed direction; no caller of the
            # reference uses it (calc_rays.py:142 always passes type='z')
            raise NotImplementedError("type='s' is built for straight_line_approx=True")
        self.type = type
        self.frequency = frequency  # Hz
        self.straight_line_approx = straight_line_approx
        self.ne_tci = ne_tci

    def cast(self, origins, directions, tmax, N):
        """All rays at once: (..., 3) origins/directions -> (..., 4, N) rows x, y, z, s."""
        lib = _lib.load()
        want_numpy = not isinstance(origins, torch.Tensor)
        o = _lib.to_device(origins)
        d = _lib.to_device(directions)
        assert o.shape == d.shape and o.shape[-1] == 3
        lead = tuple(o.shape[:-1])
        nrays = int(np.prod(lead)) if lead else 1
        rays = torch.empty(lead + (4, int(N)), dtype=torch.float64, device=o.device)
        _lib.call("iono_cast_rays_straight_f64" if self.type == 'z' else "iono_cast_rays_arclength_f64", _lib.ptr(o),
                  _lib.ptr(d), nrays, float(tmax), int(N), _lib.ptr(rays), _lib.stream_ptr())
        if not self.straight_line_approx:
            # the shipped "curved" mode: same geometry, s becomes the optical path int n dz/pz
            # (euler_ode with grad n == 0, fermat.py:53-55,57-66)
            import ctypes
            assert self.ne_tci is not None, "straight_line_approx=False needs ne_tci"
            ne = self.ne_tci.device_M()
            n_field = torch.empty_like(ne)
            _lib.call("iono_ne_to_refractive_index_f64", _lib.ptr(ne), ne.numel(), float(self.frequency),
                      _lib.ptr(n_field), _lib.stream_ptr())
            oob = torch.zeros(1, dtype=torch.int64, device=rays.device)
            _lib.call("iono_optical_path_f64", self.ne_tci.grid().handle, _lib.ptr(n_field), _lib.ptr(rays), nrays,
                      int(N), ctypes.c_void_p(oob.data_ptr()), _lib.stream_ptr())
            if int(oob.item()) != 0:
                raise ValueError("One of the requested xi is out of bounds (%d quadrature points outside the grid)"
                                 % int(oob.item()))
        return rays.cpu().numpy() if want_numpy else rays

    def integrate_ray(self, origin, direction, tmax, N=100):
        """One ray: returns ``x, y, z, s`` each of length ``N`` (fermat.py:150-174)."""
        rays = self.cast(np.asarray(origin, dtype=np.float64).reshape(1, 3),
                         np.asarray(direction, dtype=np.float64).reshape(1, 3), tmax, N)
        return rays[0, 0], rays[0, 1], rays[0, 2], rays[0, 3]
