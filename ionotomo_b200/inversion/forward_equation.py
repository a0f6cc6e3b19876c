"""TEC forward model: ``forward_equation`` of ``inversion/forward_equation.py:36-67``.

``dtec[a,t,d] = TEC[a,t,d] - TEC[i0,t,d]`` with ``TEC = simps(interp(ne; x,y,z), s)`` along
each ray and ``ne = K_ne * exp(m) / 1e13`` per voxel.  One warp per ray on the GPU
(``iono_tec_forward_f64``); results return in the kind of ``rays`` (NumPy or CUDA tensor).
"""
import ctypes

import torch

from .. import _lib

TECU = 1e13  # inversion/forward_equation.py:12


def _ne_from_m(m_dev, K_ne):
    lib = _lib.load()
    ne = torch.empty_like(m_dev)
    _lib.call("iono_ne_from_m_f64", _lib.ptr(m_dev), m_dev.numel(), float(K_ne) / TECU, _lib.ptr(ne),
              _lib.stream_ptr())
    return ne


def tec_from_ne(rays_dev, grid, ne_dev, order="time", check_bounds=True):
    """Absolute TEC per ray, ``(Na, Nt, Nd)`` CUDA tensor (do_forward_equation,
    forward_equation.py:13-33).  Raises ``ValueError`` like SciPy's
    ``bounds_error=True`` interpolator if a sample leaves the grid."""
    lib = _lib.load()
    Na, Nt, Nd, four, Ns = rays_dev.shape
    assert four == 4
    tec = torch.empty((Na, Nt, Nd), dtype=torch.float64, device=rays_dev.device)
    oob = torch.zeros(1, dtype=torch.int64, device=rays_dev.device)
    _lib.call("iono_tec_forward_f64", grid.handle, _lib.ptr(ne_dev), _lib.ptr(rays_dev), Na, Nt, Nd, Ns,
              _lib.ORDERS[order], _lib.ptr(tec), ctypes.c_void_p(oob.data_ptr()),
              _lib.stream_ptr())
    if check_bounds and int(oob.item()) != 0:
        raise ValueError("One of the requested xi is out of bounds (%d ray samples outside the grid)"
                         % int(oob.item()))
    return tec


def forward_equation(rays, K_ne, m_tci, i0, order="time", check_bounds=True, return_tec=False):
    """For each ray do the forward equation using reference antenna ``i0``
    (forward_equation.py:36-51).  ``m_tci`` is not modified."""
    lib = _lib.load()
    want_numpy = not isinstance(rays, torch.Tensor)
    rays_dev = _lib.to_device(rays)
    Na, Nt, Nd, _, Ns = rays_dev.shape
    ne = _ne_from_m(m_tci.device_M(), K_ne)
    tec = tec_from_ne(rays_dev, m_tci.grid(), ne, order=order, check_bounds=check_bounds)
    dtec = torch.empty_like(tec)
    _lib.call("iono_dtec_f64", _lib.ptr(tec), Na, Nt, Nd, int(i0), _lib.ptr(dtec), _lib.stream_ptr())
    if return_tec:
        return (dtec.cpu().numpy(), tec.cpu().numpy()) if want_numpy else (dtec, tec)
    return dtec.cpu().numpy() if want_numpy else dtec


# the reference's dask variant computes the same numbers (tests/test_forward_equation.py:26-27)
forward_equation_dask = forward_equation
