"""Phase-domain forward operators of the reference's generation B
(``inversion/iterative_newton.py:86-184``): ``forward_equation(model, tci, rays, freqs, K, i0)``
and ``prior_penalty_mu(model, model_prior, tci, rays, freqs, K, i0)``.

Same gather + Simpson sweep as the TEC forward with a different integrand, evaluated for all
frequencies in one pass over the rays (``iono_phase_integrals_f64``).

Note on parity: the reference calls ``tci.interp`` on 4-D coordinate arrays, where
``np.array([x,y,z]).T`` reverses all axes (geometry/tri_cubic.py:69-70) and scrambles which
sample belongs to which ray.  This implementation integrates each ray over its own samples
(the evident intent); the oracle can reproduce the reference's scrambled values for pinning
(``reference_axis_scramble=True``), see DESIGN.md.
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


def forward_equation(model, tci, rays, freqs, K=1e11, i0=0, order="time", check_bounds=True):
    """Phase ``(Na, Nt, Nd, Nf)`` from ``model = (mu, clock, const)`` (iterative_newton.py:86-127).
    Like the reference this leaves ``tci.M = K*exp(mu)``."""
    want_numpy = not isinstance(rays, torch.Tensor)
    rays_d = _lib.to_device(rays)
    Na, Nt, Nd, _, Ns = rays_d.shape
    mu, clock, const = model
    shape = (tci.nx, tci.ny, tci.nz)
    _, ne = _ne_from_mu(mu, K, shape, rays_d.device)
    tci.M = ne if isinstance(tci.M, torch.Tensor) else ne.cpu().numpy()      # iterative_newton.py:106
    I, f_h = _integrals(tci, ne, None, rays_d, freqs, order, check_bounds)
    clock_d = _lib.to_device(clock).reshape(Na, Nt)
    const_d = _lib.to_device(const).reshape(Na)
    g = torch.empty_like(I)
    _lib.call("iono_phase_assemble_f64", _lib.ptr(I), Na, Nt, Nd, f_h.size, int(i0), f_h.ctypes.data,
              _lib.ptr(clock_d), _lib.ptr(const_d), 0, _lib.ptr(g), _lib.stream_ptr())
    return g.cpu().numpy() if want_numpy else g


def prior_penalty_mu(model, model_prior, tci, rays, freqs, K=1e11, i0=0, order="time", check_bounds=True):
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
