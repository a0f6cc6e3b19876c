"""Linear-operator view of the ray integrals: ``RayOp`` / ``TECForwardEquation`` of the reference's
TensorFlow generation (``tomography/linear_operators.py:7-98``):

    h[i1..ir] = simps( w * interp(M (.) x ; rays), dx )          (minus the reference antenna's ray)

``rays`` is ``(..., 3, N)`` (x, y, z rows); the abscissa ``dx`` defaults to the cumulative
point-to-point distance along each ray (``linear_operators.py:25-28``).  Out-of-range points are
extrapolated from the edge cell instead of raising, like the TF interpolator
(``tomography/interpolation.py:186-188``).  ``matmul(x, adjoint=True)`` applies the exact
transpose (what ``tf.gradients`` gave the reference).
"""
import numpy as np
import torch

from .. import _lib
from ..geometry.tri_cubic import TriCubic
from ..inversion.forward_equation import tec_from_ne
from ..inversion.gradient import backproject


class RayOp(object):
    def __init__(self, grid, M, rays, dx=None, weight=None, transpose=False):
        self.want_numpy = not isinstance(rays, torch.Tensor)
        r = _lib.to_device(rays)
        assert r.shape[-2] == 3
        self.lead = tuple(r.shape[:-2])
        N = r.shape[-1]
        r = r.reshape(-1, 3, N)
        if dx is None:
            seg = torch.sqrt(((r[..., 1:] - r[..., :-1]) ** 2).sum(-2))
            s = torch.cat([torch.zeros_like(seg[..., :1]), torch.cumsum(seg, -1)], -1)
        else:
            s = _lib.to_device(dx, r.device).reshape(1, N).expand(r.shape[0], N)
        # pack as the (Na, Nt, Nd, 4, N) layout of the sweep kernels with Nt = Nd = 1
        self.rays4 = torch.cat([r, s[:, None, :]], 1).reshape(r.shape[0], 1, 1, 4, N).contiguous()
        self.M = _lib.to_device(M, r.device)
        xv, yv, zv = (_lib.host_f64(g) for g in grid)
        self.tci = TriCubic(xv, yv, zv, self.M)
        self.weight = None if weight is None else _lib.to_device(weight, r.device).reshape(-1)
        self.transpose = transpose

    def domain_shape(self):
        return tuple(self.M.shape)

    def range_shape(self):
        return self.lead

    def shape(self):
        return self.range_shape() + self.domain_shape()

    def _forward(self, x):
        f = (self.M * _lib.to_device(x, self.M.device).reshape(self.M.shape)).contiguous()
        h = tec_from_ne(self.rays4, self.tci.grid(), f, order="natural", check_bounds=False).reshape(-1)
        if self.weight is not None:
            h = h * self.weight    # a per-ray weight commutes with the (linear) quadrature
        return h

    def _adjoint(self, y):
        c = _lib.to_device(y, self.M.device).reshape(-1)
        if self.weight is not None:
            c = c * self.weight
        acc = backproject(self.rays4, self.tci.grid(), c.reshape(-1, 1, 1).contiguous(), tuple(self.M.shape),
                          order="natural", check_bounds=False)
        return acc * self.M

    def matmul(self, x, adjoint=False, adjoint_arg=False):
        out = self._adjoint(x) if (adjoint != self.transpose) else self._forward(x).reshape(self.lead)
        return out.cpu().numpy() if self.want_numpy else out


class TECForwardEquation(RayOp):
    """``RayOp`` minus the ray of reference antenna ``i0`` along the first ray axis
    (linear_operators.py:75-98)."""

    def __init__(self, i0, grid, M, rays, dx=None, weight=None, transpose=False):
        super(TECForwardEquation, self).__init__(grid, M, rays, dx, weight, transpose)
        self.i0 = int(i0)

    def matmul(self, x, adjoint=False, adjoint_arg=False):
        if adjoint != self.transpose:
            y = _lib.to_device(x, self.M.device).reshape(self.lead).clone()
            y[self.i0] -= y.sum(0)                      # transpose of  A x - (A x)[i0]
            out = self._adjoint(y)
        else:
            h = self._forward(x).reshape(self.lead)
            out = h - h[self.i0:self.i0 + 1]
        return out.cpu().numpy() if self.want_numpy else out
