"""Model-covariance application on the grid: ``Covariance.smooth`` of the reference
(``ionosphere/covariance.py:8-63,383-385``): a numerical stencil of the covariance kernel is
convolved with a grid function (``scipy.ndimage.convolve(phi, c_stencil, mode='nearest')``) --
the ``Cm . (G^T r)`` step that follows the adjoint in the reference's solvers.

The reference builds its stencil from symbolic GP kernels (``utils/gaussian_process.py``, out of
scope); here the stencil is passed in, or built from the closed-form separable Matern-p kernel
the reference uses by default (``MaternPSep(3, d, l=20, sigma=1, p=0)`` per axis = exponential).
"""
import numpy as np
import torch

from .. import _lib


def exponential_sep_stencil(dx, dy, dz, sigma=1., l=20., threshold=0.05, max_m=29):
    """Stencil of the default kernel of ``Covariance.__init__`` (covariance.py:21-23): the product over
    the three axes of Matern-1/2 (exponential) kernels ``sigma^2 exp(-|r_d|/l)``, grown from 5 points
    per axis until its edge falls below ``threshold`` of its centre (``create_c_stencil``, :46-62)."""
    m = 5
    while True:
        ax = [np.linspace(-d * (m >> 1), d * (m >> 1), m) for d in (dx, dy, dz)]
        X, Y, Z = np.meshgrid(*ax, indexing='ij')
        c = (sigma ** 2 * np.exp(-np.abs(X) / l)) * (sigma ** 2 * np.exp(-np.abs(Y) / l)) \
            * (sigma ** 2 * np.exp(-np.abs(Z) / l))
        if np.min(c) / np.max(c) <= threshold or m + 2 > max_m:
            return c
        m += 2


class Covariance(object):
    def __init__(self, c_stencil=None, dx=None, dy=None, dz=None, tci=None, sigma=1., l=20.):
        if tci is not None:
            dx, dy, dz = (tci.xvec[1] - tci.xvec[0], tci.yvec[1] - tci.yvec[0], tci.zvec[1] - tci.zvec[0])
        self.dx, self.dy, self.dz = dx, dy, dz
        if c_stencil is None and dx is not None:
            c_stencil = exponential_sep_stencil(dx, dy, dz, sigma=sigma, l=l)
        self.c_stencil = None if c_stencil is None else np.ascontiguousarray(c_stencil, dtype=np.float64)
        if self.c_stencil is not None:
            m = self.c_stencil.shape[0]
            assert self.c_stencil.shape == (m, m, m) and m % 2 == 1

    def smooth(self, phi):
        """``Cm . phi`` by the stencil (covariance.py:383-385). NumPy in -> NumPy out, CUDA in -> CUDA out."""
        want_numpy = not isinstance(phi, torch.Tensor)
        p = _lib.to_device(phi)
        assert p.dim() == 3
        w = _lib.to_device(self.c_stencil, p.device)
        out = torch.empty_like(p)
        _lib.call("iono_convolve3d_nearest_f64", _lib.ptr(p), p.shape[0], p.shape[1], p.shape[2], _lib.ptr(w),
                  int(self.c_stencil.shape[0]), _lib.ptr(out), _lib.stream_ptr())
        return out.cpu().numpy() if want_numpy else out
