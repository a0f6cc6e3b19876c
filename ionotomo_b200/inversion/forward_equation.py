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


def quads_alloc(shape, device):
    """Buffer for the quad layout of an ``(nx, ny, nz)`` field (32-byte aligned: torch allocations are)."""
    q = torch.empty(tuple(shape) + (4,), dtype=torch.float64, device=device)
    assert q.data_ptr() % 32 == 0
    return q


def ne_quads_from_m(m_dev, K_ne, ne_out=None, quads_out=None, want_ne=True):
    """``ne = K_ne exp(m)/1e13`` (forward_equation.py:41-43) in one launch as the plain grid (optional)
    and as the quad records the forward gathers from.  Returns ``(ne or None, quads)``."""
    nx, ny, nz = m_dev.shape
    q = quads_out if quads_out is not None else quads_alloc(m_dev.shape, m_dev.device)
    ne = (ne_out if ne_out is not None else torch.empty_like(m_dev)) if want_ne else None
    _lib.call("iono_ne_quads_from_m_f64", _lib.ptr(m_dev), nx, ny, nz, float(K_ne) / TECU,
              _lib.ptr(ne) if ne is not None else None, _lib.ptr(q), _lib.stream_ptr())
    return ne, q


def tec_from_quads(rays_dev, grid, quads, order="time", check_bounds=True, out=None, oob=None):
    """``tec_from_ne`` on the quad layout of ne (``ne_quads_from_m``)."""
    Na, Nt, Nd, four, Ns = rays_dev.shape
    assert four == 4
    tec = out if out is not None else torch.empty((Na, Nt, Nd), dtype=torch.float64, device=rays_dev.device)
    if oob is None:
        oob = torch.zeros(1, dtype=torch.int64, device=rays_dev.device)
    _lib.call("iono_tec_forward_quads_f64", grid.handle, _lib.ptr(quads), _lib.ptr(rays_dev), Na, Nt, Nd, Ns,
              _lib.ORDERS[order], _lib.ptr(tec), ctypes.c_void_p(oob.data_ptr()), _lib.stream_ptr())
    if check_bounds and int(oob.item()) != 0:
        raise ValueError("One of the requested xi is out of bounds (%d ray samples outside the grid)"
                         % int(oob.item()))
    return tec


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


class ForwardProjector(object):
    """Prepared form of the TEC forward for a fixed ray geometry (``iono_forwardprojector_*``) -- and of its
    adjoint: the same record stream applied transposed.

    Build once per solve: cell index, in-cell fractions and Simpson weight of every sample are computed
    once and streamed afterwards (36 B per sample of HBM; 28 B when the weights factor into a common pattern x a
    per-ray factor, ``self.factored``), so ``tec(ne)`` is gathers and fmas only.  With per-sample weights the
    result is bit-identical to ``tec_from_ne`` on the same rays, with factored weights equal to ~1e-14 relative.
    Raises ``ValueError`` at construction when a sample lies outside the grid, as every forward on those rays would.
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
        _lib.call("iono_forwardprojector_create", self._grid.handle, _lib.ptr(rays_dev), Na, Nt, Nd, Ns,
                  ctypes.byref(h), ctypes.c_void_p(oob.data_ptr()), _lib.stream_ptr())
        self.handle = h
        self.nbytes = int(lib.iono_forwardprojector_bytes(h))
        self.factored = bool(lib.iono_forwardprojector_factored(h))   # Simpson weights stored as pattern x per-ray factor
        if check_bounds and int(oob.item()) != 0:
            raise ValueError("One of the requested xi is out of bounds (%d ray samples outside the grid)"
                             % int(oob.item()))

    def tec(self, ne_dev, out=None):
        """``TEC[a,t,d] = simps(interp(ne; ray), s)`` for the ``(nx,ny,nz)`` CUDA tensor ``ne_dev``."""
        assert tuple(ne_dev.shape) == self.shape and ne_dev.is_contiguous()
        tec = out if out is not None else torch.empty(self.ray_shape, dtype=torch.float64, device=ne_dev.device)
        _lib.call("iono_forwardprojector_apply_f64", self.handle, _lib.ptr(ne_dev), _lib.ptr(tec),
                  _lib.stream_ptr())
        return tec

    def tec_quads(self, quads, out=None):
        """``tec`` on the quad layout of ne (``ne_quads_from_m``): no per-call rewrite of the grid."""
        assert tuple(quads.shape) == self.shape + (4,) and quads.is_contiguous()
        tec = out if out is not None else torch.empty(self.ray_shape, dtype=torch.float64, device=quads.device)
        _lib.call("iono_forwardprojector_apply_quads_f64", self.handle, _lib.ptr(quads), _lib.ptr(tec),
                  _lib.stream_ptr())
        return tec

    # ---- the transpose ------------------------------------------------------------------------------
    @property
    def n_voxels(self):
        return int(_lib.load().iono_forwardprojector_n_voxels(self.handle))

    def voxels(self):
        """Flat indices of the grid nodes the operator touches (corners of visited cells), ascending, int32."""
        out = torch.empty(max(self.n_voxels, 1), dtype=torch.int32, device=self.device)
        _lib.call("iono_forwardprojector_voxels", self.handle, ctypes.c_void_p(out.data_ptr()), _lib.stream_ptr())
        return out[:self.n_voxels]

    def adjoint(self, coef_perm, acc):
        """``acc[v] += sum_ray coef[ray] A[ray, v]`` with ``A`` the matrix ``tec`` applies; ``coef_perm``: the
        coefficients in (antenna, direction, time) order as ``residual(..., want_perm=True)`` writes them; ``acc``:
        the ``(nx,ny,nz)`` accumulator (zero where the operator touches it, see ``finish_*``)."""
        Na, Nt, Nd = self.ray_shape
        assert coef_perm.numel() == Na * Nt * Nd and coef_perm.is_contiguous()
        assert tuple(acc.shape) == self.shape and acc.is_contiguous()
        _lib.call("iono_forwardprojector_adjoint_f64", self.handle, _lib.ptr(coef_perm), _lib.ptr(acc), _lib.stream_ptr())
        return acc

    def finish_gradient(self, acc, m, k, grad):
        """``grad[v] = k exp(m[v]) acc[v]`` and ``acc[v] = 0`` on the touched voxels (``grad`` elsewhere untouched)."""
        _lib.call("iono_forwardprojector_finish_gradient_f64", self.handle, _lib.ptr(acc), _lib.ptr(m), float(k),
                  _lib.ptr(grad), _lib.stream_ptr())
        return grad

    def finish_compact(self, acc, out, dst=None):
        """``out[dst[i]] = acc[voxel_i]`` (``dst`` None: ``out[i]``) and ``acc[voxel_i] = 0``."""
        _lib.call("iono_forwardprojector_finish_compact_f64", self.handle, _lib.ptr(acc),
                  ctypes.c_void_p(dst.data_ptr()) if dst is not None else None, _lib.ptr(out), _lib.stream_ptr())
        return out

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.load().iono_forwardprojector_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


def forward_equation(rays, K_ne, m_tci, i0, order="time", check_bounds=True, return_tec=False, projector=None):
    """For each ray do the forward equation using reference antenna ``i0``
    (forward_equation.py:36-51).  ``m_tci`` is not modified.  ``projector`` (optional
    ``ForwardProjector`` built from the same rays and grid axes) replaces the stateless sweep."""
    lib = _lib.load()
    want_numpy = not isinstance(rays, torch.Tensor)
    rays_dev = _lib.to_device(rays)
    Na, Nt, Nd, _, Ns = rays_dev.shape
    ne = _ne_from_m(m_tci.device_M(), K_ne)
    if projector is not None:
        assert projector.ray_shape == (Na, Nt, Nd)
        tec = projector.tec(ne)
    else:
        tec = tec_from_ne(rays_dev, m_tci.grid(), ne, order=order, check_bounds=check_bounds)
    dtec = torch.empty_like(tec)
    _lib.call("iono_dtec_f64", _lib.ptr(tec), Na, Nt, Nd, int(i0), _lib.ptr(dtec), _lib.stream_ptr())
    if return_tec:
        return (dtec.cpu().numpy(), tec.cpu().numpy()) if want_numpy else (dtec, tec)
    return dtec.cpu().numpy() if want_numpy else dtec


# the reference's dask variant computes the same numbers (tests/test_forward_equation.py:26-27)
forward_equation_dask = forward_equation
