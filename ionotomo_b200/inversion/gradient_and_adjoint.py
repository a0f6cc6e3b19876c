"""``Cm.G^t.Cd^-1.(g - dobs) + (m - m_prior)`` with a Gaussian model covariance: the reference's
"adjoint B", ``compute_adjoint`` / ``compute_adjoint_dask`` of inversion/gradient_and_adjoint.py:106-167.

The five nested loops of ``do_adjoint`` (:12-103) become one kernel (``iono_gaussian_adjoint_f64``, one warp
per ray); ``ne`` along the rays comes from the point-wise interpolation and exponential kernels.  All
directions are processed in one launch (the reference sums ``do_adjoint`` over directions, :162-163).
"""
import torch

from .. import _lib
from .forward_equation import _ne_from_m


def gaussian_adjoint(rays, dd, K_ne, m_tci, sigma_m, Nkernel, size_cell, out=None):
    """``sum_ray dd[ray] * simps(sigma_m^2 exp(-r^2/(2 L_m^2)) ne(s), s)`` per voxel, device tensors in and
    out (gradient_and_adjoint.py:12-100 without the final plane subtraction).  Raises ``ValueError`` when a
    ray sample lies outside the grid, like the ``m_tci.interp`` call at :37."""
    Na, Nt, Nd, _, Ns = rays.shape
    m_ray = m_tci.interp(rays[:, :, :, 0, :], rays[:, :, :, 1, :], rays[:, :, :, 2, :])    # (Na,Nt,Nd,Ns)
    ne_ray = _ne_from_m(m_ray.contiguous(), K_ne)
    m_dev = m_tci.device_M()
    acc = out if out is not None else torch.empty(tuple(m_dev.shape), dtype=torch.float64, device=rays.device)
    _lib.call("iono_gaussian_adjoint_f64", m_tci.grid().handle, _lib.ptr(rays), Na, Nt, Nd, Ns, _lib.ptr(ne_ray),
              _lib.ptr(dd), float(sigma_m), float(Nkernel) * float(size_cell), int(Nkernel), 1, _lib.ptr(acc),
              _lib.stream_ptr())
    return acc


def compute_adjoint(rays, g, dobs, i0, K_ne, m_tci, m_prior, CdCt, sigma_m, Nkernel, size_cell,
                    bug_compat=False):
    """Same signature as the reference (gradient_and_adjoint.py:137).  NumPy in -> NumPy out, CUDA tensors in ->
    CUDA tensor out.  ``bug_compat=True`` also applies ``grad -= grad[i0, :, :]`` (:102), which uses the
    antenna index on the grid's x axis."""
    want_numpy = not isinstance(rays, torch.Tensor)
    rays_dev = _lib.to_device(rays)
    g_d, dobs_d, C_d = _lib.to_device(g), _lib.to_device(dobs), _lib.to_device(CdCt)
    dd = ((g_d - dobs_d) / (C_d + 1e-15)).contiguous()
    acc = gaussian_adjoint(rays_dev, dd, K_ne, m_tci, sigma_m, Nkernel, size_cell)
    if bug_compat:
        acc = acc - acc[int(i0), :, :]
    acc = acc + m_tci.device_M() - _lib.to_device(m_prior)
    return acc.cpu().numpy() if want_numpy else acc


compute_adjoint_dask = compute_adjoint
