"""Phase-domain forward operators of the reference's generation B
(``inversion/iterative_newton.py:86-184``): ``forward_equation(model, tci, rays, freqs, K, i0)``
and ``prior_penalty_mu(model, model_prior, tci, rays, freqs, K, i0)``.

Same gather + Simpson sweep as the TEC forward with a different integrand, evaluated for all
frequencies in one pass over the rays (``iono_phase_integrals_f64``).

Note on parity: the reference calls ``tci.interp`` on 4-D coordinate arrays, where
``np.array([x,y,z]).T`` reverses all axes (geometry/tri_cubic.py:69-70) and scrambles which
sample belongs to which ray.  By default each ray is integrated over its own samples (the evident
intent, one fused sweep); ``reference_axis_scramble=True`` reproduces the reference's own output:
point-wise interpolation of all samples (``iono_tci_interp_f64``, SciPy's arithmetic), the same
transpose-and-reshape, then Simpson along each ray's ``s`` (``iono_simps_rows_f64``) -- pinned to
``tests/golden/forward_*.npz['phase', 'penalty']``, which the reference's module produced.
"""
import ctypes

import numpy as np
import torch

from .. import _lib

SPEED_OF_LIGHT = 299792458.  # iterative_newton.py:15


def _integrals(tci, ne_dev, dmu_dev, rays_dev, freqs, order, check_bounds):
    Na, Nt, Nd, four, Ns = rays_dev.shape
    assert four == 4
    f_h = np.ascontiguousarray(_lib.host_f64(freqs), dtype=np.float64).reshape(-1)
    Nf = f_h.size
    assert 1 <= Nf <= 8, "up to 8 frequencies per call"
    out = torch.empty((Na, Nt, Nd, Nf), dtype=torch.float64, device=rays_dev.device)
    oob = torch.zeros(1, dtype=torch.int64, device=rays_dev.device)
    _lib.call("iono_phase_integrals_f64", tci.grid().handle, _lib.ptr(ne_dev),
              _lib.ptr(dmu_dev) if dmu_dev is not None else None, _lib.ptr(rays_dev), Na, Nt, Nd, Ns,
              f_h.ctypes.data, Nf, _lib.ORDERS[order], _lib.ptr(out), ctypes.c_void_p(oob.data_ptr()),
              _lib.stream_ptr())
    if check_bounds and int(oob.item()) != 0:
        raise ValueError("One of the requested xi is out of bounds (%d ray samples outside the grid)"
                         % int(oob.item()))
    return out, f_h


def _ne_from_mu(mu, K, shape, device):
    mu_d = _lib.to_device(mu, device).reshape(shape)
    ne = torch.empty_like(mu_d)
    _lib.call("iono_ne_from_m_f64", _lib.ptr(mu_d), mu_d.numel(), float(K), _lib.ptr(ne), _lib.stream_ptr())
    return mu_d, ne


def _scrambled_samples(tci, field_dev, rays_d, check_bounds):
    """``tci.interp(rays[...,0,:], rays[...,1,:], rays[...,2,:])`` exactly as the reference evaluates it on 4-D
    inputs (geometry/tri_cubic.py:69-70): values in the order of the fully transposed coordinate array, reshaped
    to ``(Na,Nt,Nd,Ns)`` without undoing the transpose."""
    Na, Nt, Nd, _, Ns = rays_d.shape
    x = rays_d[:, :, :, 0, :].contiguous().reshape(-1)
    y = rays_d[:, :, :, 1, :].contiguous().reshape(-1)
    z = rays_d[:, :, :, 2, :].contiguous().reshape(-1)
    v = torch.empty_like(x)
    oob = torch.zeros(1, dtype=torch.int64, device=rays_d.device)
    _lib.call("iono_tci_interp_f64", tci.grid().handle, _lib.ptr(field_dev), _lib.ptr(x), _lib.ptr(y), _lib.ptr(z),
              x.numel(), 0, _lib.ptr(v), ctypes.c_void_p(oob.data_ptr()), _lib.stream_ptr())
    if check_bounds and int(oob.item()) != 0:
        raise ValueError("One of the requested xi is out of bounds in dimension 0")
    return v.reshape(Na, Nt, Nd, Ns).permute(3, 2, 1, 0).contiguous().reshape(Na, Nt, Nd, Ns)


def _integrals_scrambled(tci, ne_dev, dmu_dev, rays_d, freqs, check_bounds):
    Na, Nt, Nd, _, Ns = rays_d.shape
    f_h = np.ascontiguousarray(_lib.host_f64(freqs), dtype=np.float64).reshape(-1)
    Nf = f_h.size
    ne_rays = _scrambled_samples(tci, ne_dev, rays_d, check_bounds)
    dmu_rays = _scrambled_samples(tci, dmu_dev, rays_d, check_bounds) if dmu_dev is not None else None
    out = torch.empty((Na, Nt, Nd, Nf), dtype=torch.float64, device=rays_d.device)
    for l in range(Nf):
        c = -1.0 / (1.2404e-2 * f_h[l] ** 2)
        _lib.call("iono_simps_rows_f64", _lib.ptr(ne_rays), _lib.ptr(dmu_rays) if dmu_rays is not None else None,
                  _lib.ptr(rays_d), Na * Nt * Nd, Ns, 2 if dmu_rays is not None else 1, float(c),
                  ctypes.c_void_p(out.data_ptr() + 8 * l), Nf, _lib.stream_ptr())
    return out, f_h


def forward_equation(model, tci, rays, freqs, K=1e11, i0=0, order="time", check_bounds=True,
                     reference_axis_scramble=False):
    """Phase ``(Na, Nt, Nd, Nf)`` from ``model = (mu, clock, const)`` (iterative_newton.py:86-127).
    Like the reference this leaves ``tci.M = K*exp(mu)``.  ``reference_axis_scramble=True``: the reference's
    own output, sample scramble included (module docstring)."""
    want_numpy = not isinstance(rays, torch.Tensor)
    rays_d = _lib.to_device(rays)
    Na, Nt, Nd, _, Ns = rays_d.shape
    mu, clock, const = model
    shape = (tci.nx, tci.ny, tci.nz)
    _, ne = _ne_from_mu(mu, K, shape, rays_d.device)
    tci.M = ne if isinstance(tci.M, torch.Tensor) else ne.cpu().numpy()      # iterative_newton.py:106
    if reference_axis_scramble:
        I, f_h = _integrals_scrambled(tci, ne, None, rays_d, freqs, check_bounds)
    else:
        I, f_h = _integrals(tci, ne, None, rays_d, freqs, order, check_bounds)
    clock_d = _lib.to_device(clock).reshape(Na, Nt)
    const_d = _lib.to_device(const).reshape(Na)
    g = torch.empty_like(I)
    _lib.call("iono_phase_assemble_f64", _lib.ptr(I), Na, Nt, Nd, f_h.size, int(i0), f_h.ctypes.data,
              _lib.ptr(clock_d), _lib.ptr(const_d), 0, _lib.ptr(g), _lib.stream_ptr())
    return g.cpu().numpy() if want_numpy else g


def prior_penalty_mu(model, model_prior, tci, rays, freqs, K=1e11, i0=0, order="time", check_bounds=True,
                     reference_axis_scramble=False):
    """First-order prior penalty ``(Na, Nt, Nd, Nf)`` (iterative_newton.py:138-184).
    Leaves ``tci.M = mu_prior - mu`` like the reference (:159)."""
    want_numpy = not isinstance(rays, torch.Tensor)
    rays_d = _lib.to_device(rays)
    Na, Nt, Nd, _, Ns = rays_d.shape
    mu = model[0]
    mu_prior = model_prior[0]
    shape = (tci.nx, tci.ny, tci.nz)
    mu_d, ne = _ne_from_mu(mu, K, shape, rays_d.device)
    dmu = (_lib.to_device(mu_prior, rays_d.device).reshape(shape) - mu_d).contiguous()
    tci.M = dmu if isinstance(tci.M, torch.Tensor) else dmu.cpu().numpy()
    if reference_axis_scramble:
        J, f_h = _integrals_scrambled(tci, ne, dmu, rays_d, freqs, check_bounds)
    else:
        J, f_h = _integrals(tci, ne, dmu, rays_d, freqs, order, check_bounds)
    r = torch.empty_like(J)
    _lib.call("iono_phase_assemble_f64", _lib.ptr(J), Na, Nt, Nd, f_h.size, int(i0), f_h.ctypes.data,
              None, None, 1, _lib.ptr(r), _lib.stream_ptr())
    return r.cpu().numpy() if want_numpy else r


def data_residuals(dobs, g):
    """Unweighted data residuals (iterative_newton.py:129-136)."""
    return dobs - g


def neg_log_like(g, dobs, CdCt):
    """``sum((dobs-g)^2/CdCt)/2`` -- the data term of iterative_newton.py:17-40 (``full=False``)."""
    g_d, dobs_d, C_d = _lib.to_device(g), _lib.to_device(dobs), _lib.to_device(CdCt)
    dd = dobs_d - g_d
    return float((dd * dd / C_d).sum() / 2.)
