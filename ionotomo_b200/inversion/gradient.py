"""Voxel gradient of the misfit: ``compute_gradient`` / ``compute_gradient_dask`` of
``inversion/gradient.py:22-100``.

The GPU adjoint is the *exact transpose* of the dTEC forward:

    grad[v] = ne[v] * sum_ray c[ray] * sum_s w_s(ray) * phi_v(x_s),
    c[ray]  = dd[ray] - [ray is (i0,t,d)] * sum_i dd[i,t,d],  dd = (g - dobs)/(CdCt + 1e-15)

i.e. dS/dm of ``S = sum((g-dobs)^2/(CdCt+1e-15))/2`` -- what the reference's own
finite-difference protocol checks (tests/test_inversion.py:71-87) and what an
L-BFGS/CG driver needs.  The reference's chord-length variant (gradient.py:15-20 over
geometry/ray_dirac.py) allocates a dense (rays x voxels) array and cannot run beyond toy
sizes; its ``gradient -= gradient[i0,...]`` (gradient.py:55) indexes grid-x, not the
antenna axis, and is not reproduced.
"""
import ctypes

import torch

from .. import _lib
from .forward_equation import _ne_from_m


def adjoint_coefficients(g, dobs, CdCt, i0):
    lib = _lib.load()
    Na, Nt, Nd = g.shape
    coef = torch.empty_like(g)
    _lib.call("iono_adjoint_coef_f64", _lib.ptr(g), _lib.ptr(dobs), _lib.ptr(CdCt), Na, Nt, Nd, int(i0),
              _lib.ptr(coef), _lib.stream_ptr())
    return coef


def residual(tec, dobs, CdCt, i0, want_coef=True, want_perm=False, out=None):
    """Everything between the forward and the adjoint in one launch (``iono_residual_f64``):
    ``dtec = tec - tec[i0]``, the misfit ``S`` (0-d CUDA tensor), the adjoint coefficients in the natural
    ``(Na,Nt,Nd)`` order (``want_coef``) and/or in the back-projector's internal (antenna, direction, time)
    order (``want_perm``, for ``BackProjector.apply_permuted``).  ``out``: optional dict of preallocated
    buffers ``dtec, coef, coef_perm, scratch, S``.  Returns ``(dtec, S, coef, coef_perm)``."""
    lib = _lib.load()
    Na, Nt, Nd = tec.shape
    out = out or {}
    dev = tec.device
    dtec = out.get("dtec") if out.get("dtec") is not None else torch.empty_like(tec)
    coef = (out.get("coef") if out.get("coef") is not None else torch.empty_like(tec)) if want_coef else None
    perm = (out.get("coef_perm") if out.get("coef_perm") is not None
            else torch.empty(Na * Nt * Nd, dtype=torch.float64, device=dev)) if want_perm else None
    scratch = out.get("scratch")
    if scratch is None:
        scratch = torch.empty(int(lib.iono_residual_scratch_elems(Na, Nt, Nd)), dtype=torch.float64, device=dev)
    S = out.get("S") if out.get("S") is not None else torch.empty(1, dtype=torch.float64, device=dev)
    _lib.call("iono_residual_f64", _lib.ptr(tec), _lib.ptr(dobs), _lib.ptr(CdCt), Na, Nt, Nd, int(i0), _lib.ptr(dtec),
              _lib.ptr(coef) if coef is not None else None, _lib.ptr(perm) if perm is not None else None,
              _lib.ptr(scratch), _lib.ptr(S), _lib.stream_ptr())
    return dtec, S[0], coef, perm


def backproject(rays_dev, grid, coef, shape, order="time", check_bounds=True, out=None):
    """``acc[v] = sum_ray coef[ray] sum_s w_s phi_v(x_s)`` (before the ``ne[v]`` factor and
    before any cross-GPU sum)."""
    lib = _lib.load()
    Na, Nt, Nd, _, Ns = rays_dev.shape
    acc = out if out is not None else torch.empty(shape, dtype=torch.float64, device=rays_dev.device)
    oob = torch.zeros(1, dtype=torch.int64, device=rays_dev.device)
    _lib.call("iono_tec_adjoint_f64", grid.handle, _lib.ptr(rays_dev), Na, Nt, Nd, Ns, _lib.ptr(coef),
              _lib.ORDERS[order], 1, _lib.ptr(acc), ctypes.c_void_p(oob.data_ptr()),
              _lib.stream_ptr())
    if check_bounds and int(oob.item()) != 0:
        raise ValueError("One of the requested xi is out of bounds (%d ray samples outside the grid)"
                         % int(oob.item()))
    return acc


class BackProjector(object):
    """Voxel-binned form of the adjoint for a fixed ray geometry (``iono_backprojector_*``).

    Build once per solve (the reference computes its rays once per solve too,
    inversion_pipeline.py:195-197), then every ``compute_gradient(..., backprojector=bp)`` is an
    atomics-free gather.  ``apply(coef)`` equals ``backproject(rays, grid, coef)`` to rounding.
    """

    def __init__(self, rays, tci, check_bounds=True):
        lib = _lib.load()
        rays_dev = _lib.to_device(rays)
        Na, Nt, Nd, four, Ns = rays_dev.shape
        assert four == 4
        self.shape = (tci.nx, tci.ny, tci.nz)
        self.ray_shape = (Na, Nt, Nd)
        self.device = rays_dev.device
        self._grid = tci.grid()          # keep the grid handle alive
        oob = torch.zeros(1, dtype=torch.int64, device=rays_dev.device)
        h = ctypes.c_void_p()
        _lib.call("iono_backprojector_create", self._grid.handle, _lib.ptr(rays_dev), Na, Nt, Nd, Ns,
                  ctypes.byref(h), ctypes.c_void_p(oob.data_ptr()), _lib.stream_ptr())
        self.handle = h
        self.nnz = int(lib.iono_backprojector_nnz(h))
        self.nbytes = int(lib.iono_backprojector_bytes(h))
        if check_bounds and int(oob.item()) != 0:
            raise ValueError("One of the requested xi is out of bounds (%d ray samples outside the grid)"
                             % int(oob.item()))

    def apply(self, coef, scale=None, out=None):
        """``out[v] = scale[v] * sum_ray A[v,ray] coef[ray]`` (``scale`` optional)."""
        coef = _lib.to_device(coef)
        assert tuple(coef.shape) == self.ray_shape
        acc = out if out is not None else torch.empty(self.shape, dtype=torch.float64, device=coef.device)
        _lib.call("iono_backprojector_apply_f64", self.handle, _lib.ptr(coef),
                  _lib.ptr(scale) if scale is not None else None, _lib.ptr(acc), _lib.stream_ptr())
        return acc

    def apply_permuted(self, coef_perm, scale=None, out=None, c0=0, c1=16):
        """``apply`` for coefficients already in the internal (antenna, direction, time) order (``residual(...,
        want_perm=True)``): no permutation pass.  ``c0, c1``: sixteenths of the operator (in order from 0)."""
        assert coef_perm.numel() == self.ray_shape[0] * self.ray_shape[1] * self.ray_shape[2]
        acc = out if out is not None else torch.empty(self.shape, dtype=torch.float64, device=coef_perm.device)
        _lib.call("iono_backprojector_apply_permuted_f64", self.handle, _lib.ptr(coef_perm),
                  _lib.ptr(scale) if scale is not None else None, _lib.ptr(acc), int(c0), int(c1), _lib.stream_ptr())
        return acc

    def chunk_voxels(self, c):
        """Flat voxel index below which ``out`` is final once chunks ``[0, c)`` have been applied."""
        return int(_lib.load().iono_backprojector_chunk_voxels(self.handle, int(c)))

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.load().iono_backprojector_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


def compute_gradient(rays, g, dobs, i0, K_ne, m_tci, m_prior, CdCt, sigma_m, Nkernel, size_cell, cov_obj=None,
                     order="time", check_bounds=True, reduce_fn=None, backprojector=None):
    """Same signature as the reference (gradient.py:66).  ``m_prior``, ``sigma_m``,
    ``Nkernel``, ``size_cell``, ``cov_obj`` are accepted for compatibility; the reference
    computes the prior term and discards it (gradient.py:56-58).

    ``reduce_fn(acc)`` (optional) is applied to the backprojection before the ``ne[v]``
    factor -- the hook for the cross-GPU allreduce when rays are sharded.
    ``backprojector`` (optional ``BackProjector`` built from the same rays and grid axes) replaces
    the atomic scatter by the pre-assembled voxel-binned gather.
    """
    lib = _lib.load()
    want_numpy = not isinstance(rays, torch.Tensor)
    rays_dev = _lib.to_device(rays)
    g_d, dobs_d, C_d = _lib.to_device(g), _lib.to_device(dobs), _lib.to_device(CdCt)
    m_dev = m_tci.device_M()
    coef = adjoint_coefficients(g_d, dobs_d, C_d, i0)
    if backprojector is not None:
        acc = backprojector.apply(coef)
    else:
        acc = backproject(rays_dev, m_tci.grid(), coef, tuple(m_dev.shape), order=order,
                          check_bounds=check_bounds)
    if reduce_fn is not None:
        acc = reduce_fn(acc)
    ne = _ne_from_m(m_dev, K_ne)
    _lib.call("iono_mul_f64", _lib.ptr(ne), _lib.ptr(acc), acc.numel(), _lib.ptr(acc), _lib.stream_ptr())
    return acc.cpu().numpy() if want_numpy else acc


compute_gradient_dask = compute_gradient


def compute_gradient_chord(rays, g, dobs, i0, K_ne, m_tci, m_prior, CdCt, sigma_m, Nkernel, size_cell,
                           cov_obj=None, bug_compat=False):
    """The reference's own generation-A gradient (gradient.py:22-62): chord lengths of each ray
    through the cell-centred voxel boxes (``get_ray_dirac``), times ``ne[v]``, times the weighted
    residual -- **not** the transpose of the forward (no Simpson weights, no reference-antenna
    term).  ``bug_compat=True`` also applies ``gradient -= gradient[i0, ...]`` (gradient.py:55),
    which indexes grid-x rather than the antenna axis."""
    want_numpy = not isinstance(rays, torch.Tensor)
    rays_dev = _lib.to_device(rays)
    Na, Nt, Nd, _, Ns = rays_dev.shape
    g_d, dobs_d, C_d = _lib.to_device(g), _lib.to_device(dobs), _lib.to_device(CdCt)
    dd = ((g_d - dobs_d) / (C_d + 1e-15)).contiguous()
    m_dev = m_tci.device_M()
    acc = torch.empty(tuple(m_dev.shape), dtype=torch.float64, device=rays_dev.device)
    _lib.call("iono_chord_adjoint_f64", m_tci.grid().handle, _lib.ptr(rays_dev), Na, Nt, Nd, Ns, _lib.ptr(dd), 1,
              _lib.ptr(acc), _lib.stream_ptr())
    ne = _ne_from_m(m_dev, K_ne)
    _lib.call("iono_mul_f64", _lib.ptr(ne), _lib.ptr(acc), acc.numel(), _lib.ptr(acc), _lib.stream_ptr())
    if bug_compat:
        acc = acc - acc[int(i0), ...]
    return acc.cpu().numpy() if want_numpy else acc


def misfit(g, dobs, CdCt):
    """``S = sum((g-dobs)^2/(CdCt+1e-15))/2`` (line_search.py:48-49) as a 0-d CUDA tensor."""
    lib = _lib.load()
    g_d, dobs_d, C_d = _lib.to_device(g), _lib.to_device(dobs), _lib.to_device(CdCt)
    scratch = torch.empty(int(lib.iono_misfit_scratch_elems()), dtype=torch.float64, device=g_d.device)
    out = torch.empty(1, dtype=torch.float64, device=g_d.device)
    _lib.call("iono_misfit_f64", _lib.ptr(g_d), _lib.ptr(dobs_d), _lib.ptr(C_d), g_d.numel(),
              _lib.ptr(scratch), _lib.ptr(out), _lib.stream_ptr())
    return out[0]
