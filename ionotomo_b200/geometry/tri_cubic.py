"""``TriCubic``: the reference's grid container + interpolator, device-backed.

Mirrors ``geometry/tri_cubic.py:13-132`` of the reference (same constructor,
properties, ``interp`` / ``extrapolate`` / ``copy`` / ``inner`` /
``get_model_coordinates`` and the module-level ``bisection``).  Despite the name
the reference interpolates *trilinearly* (it wraps SciPy's
``RegularGridInterpolator`` with the default ``method='linear'``,
``tri_cubic.py:22``), and so does this class -- on the GPU, through
``iono_tci_interp_f64``.

``M`` may be a NumPy array (as in the reference) or a float64 CUDA tensor (kept
resident; nothing is copied per call).  Results come back in the kind of the
query points (NumPy in -> NumPy out, CUDA tensor in -> CUDA tensor out).
"""
import ctypes
import weakref

import numpy as np
import torch

from .. import _lib


class _GridHandle(object):
    """Owns an ``iono_grid_t`` (device cell tables of the three axes)."""

    def __init__(self, xvec, yvec, zvec):
        lib = _lib.load()
        _lib.require_cuda()
        self.xvec = np.ascontiguousarray(xvec, dtype=np.float64)
        self.yvec = np.ascontiguousarray(yvec, dtype=np.float64)
        self.zvec = np.ascontiguousarray(zvec, dtype=np.float64)
        self.device = torch.cuda.current_device()
        h = ctypes.c_void_p()
        _lib.call("iono_grid_create", self.xvec.ctypes.data, self.yvec.ctypes.data, self.zvec.ctypes.data,
                  self.xvec.size, self.yvec.size, self.zvec.size, ctypes.byref(h))
        self.handle = h
        self.uniform = bool(lib.iono_grid_is_uniform(h))

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.load().iono_grid_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


_grid_cache = weakref.WeakValueDictionary()


def grid_handle(xvec, yvec, zvec):
    """Handle for these axes on the current device (shared between copies of a TriCubic)."""
    _lib.require_cuda()
    key = (torch.cuda.current_device(), np.asarray(xvec, dtype=np.float64).tobytes(),
           np.asarray(yvec, dtype=np.float64).tobytes(), np.asarray(zvec, dtype=np.float64).tobytes())
    h = _grid_cache.get(key)
    if h is None:
        h = _GridHandle(xvec, yvec, zvec)
        _grid_cache[key] = h
    return h


class TriCubic(object):
    def __init__(self, xvec=None, yvec=None, zvec=None, M=None, filename=None):
        if filename is not None:
            self.load(filename)
        else:
            self.xvec = xvec
            self.yvec = yvec
            self.zvec = zvec
            self.M = M

    # --- axes (host, as in the reference: tri_cubic.py:24-44) ---------------
    @property
    def xvec(self):
        return self._xvec

    @xvec.setter
    def xvec(self, val):
        self._xvec = np.array(_lib.host_f64(val))
        self.nx = int(np.size(self._xvec))
        self._grid = None

    @property
    def yvec(self):
        return self._yvec

    @yvec.setter
    def yvec(self, val):
        self._yvec = np.array(_lib.host_f64(val))
        self.ny = int(np.size(self._yvec))
        self._grid = None

    @property
    def zvec(self):
        return self._zvec

    @zvec.setter
    def zvec(self, val):
        self._zvec = np.array(_lib.host_f64(val))
        self.nz = int(np.size(self._zvec))
        self._grid = None

    # --- values (tri_cubic.py:45-59) ----------------------------------------
    @property
    def M(self):
        return self._M

    @M.setter
    def M(self, val):
        if isinstance(val, torch.Tensor):
            assert bool(torch.isfinite(val).all())
            if val.dim() == 1:
                val = val.reshape(len(self.xvec), len(self.yvec), len(self.zvec))
        else:
            val = np.asarray(val)
            assert not np.any(np.isnan(val)) and not np.any(np.isinf(val))
            if len(val.shape) == 1:
                val = val.reshape((len(self.xvec), len(self.yvec), len(self.zvec)))
        assert val.shape[0] == len(self.xvec)
        assert val.shape[1] == len(self.yvec)
        assert val.shape[2] == len(self.zvec)
        self._M = val

    def get_shaped_array(self):
        """Older accessor still used by the reference's line search (line_search.py:46)."""
        return self._M

    # --- device views ---------------------------------------------------------
    def grid(self):
        if getattr(self, "_grid", None) is None or self._grid.device != torch.cuda.current_device():
            self._grid = grid_handle(self._xvec, self._yvec, self._zvec)
        return self._grid

    def device_M(self):
        """Contiguous float64 CUDA tensor of ``M`` (the tensor itself if already one)."""
        return _lib.to_device(self._M)

    # --- interpolation (tri_cubic.py:69-75) ------------------------------------
    def _sample(self, x, y, z, extrapolate):
        lib = _lib.load()
        want_numpy = not isinstance(x, torch.Tensor)
        shape = tuple(np.shape(x)) if want_numpy else tuple(x.shape)
        xd = _lib.to_device(x).reshape(-1)
        yd = _lib.to_device(y).reshape(-1)
        zd = _lib.to_device(z).reshape(-1)
        assert xd.numel() == yd.numel() == zd.numel()
        out = torch.empty_like(xd)
        oob = torch.zeros(1, dtype=torch.int64, device=xd.device)
        Md = self.device_M()
        _lib.call("iono_tci_interp_f64", self.grid().handle, _lib.ptr(Md), _lib.ptr(xd), _lib.ptr(yd),
                  _lib.ptr(zd), xd.numel(), int(extrapolate), _lib.ptr(out),
                  ctypes.c_void_p(oob.data_ptr()), _lib.stream_ptr())
        if not extrapolate and int(oob.item()) != 0:
            # scipy RegularGridInterpolator(bounds_error=True) behaviour (tri_cubic.py:22)
            raise ValueError("One of the requested xi is out of bounds (%d points)" % int(oob.item()))
        out = out.reshape(shape)
        return out.cpu().numpy() if want_numpy else out

    def interp(self, x, y, z):
        return self._sample(x, y, z, False)

    def extrapolate(self, x, y, z):
        return self._sample(x, y, z, True)

    # --- misc -------------------------------------------------------------------
    def inner(self, M, inplace=False):
        """Inner product of ``self.M`` with ``M``: triple Simpson (tri_cubic.py:61-67)."""
        a = torch.as_tensor(_lib.host_f64(self._M)) if not isinstance(self._M, torch.Tensor) else self._M
        b = torch.as_tensor(_lib.host_f64(M)) if not isinstance(M, torch.Tensor) else M
        b = b.to(a.device)
        if inplace and isinstance(M, torch.Tensor):
            b *= a
            P = b
        elif inplace:
            M *= _lib.host_f64(self._M)
            P = torch.as_tensor(M)
        else:
            P = a * b
        P = P.to(torch.float64)
        r = _simps_avg_torch(P, torch.as_tensor(self._zvec, device=P.device))
        r = _simps_avg_torch(r, torch.as_tensor(self._yvec, device=P.device))
        r = _simps_avg_torch(r, torch.as_tensor(self._xvec, device=P.device))
        return float(r)

    def copy(self, **kwargs):
        M = self._M.clone() if isinstance(self._M, torch.Tensor) else self._M.copy()
        return TriCubic(self._xvec.copy(), self._yvec.copy(), self._zvec.copy(), M, **kwargs)

    def load(self, filename, **kwargs):
        """HDF5 group ``TCI/{xvec,yvec,zvec,M}`` (tri_cubic.py:81-88); needs h5py."""
        import h5py
        with h5py.File(filename, 'r') as f:
            xvec = f["TCI/xvec"][:]
            yvec = f["TCI/yvec"][:]
            zvec = f["TCI/zvec"][:]
            M = f["TCI/M"][:, :, :]
        self.__init__(xvec, yvec, zvec, M, **kwargs)

    def save(self, filename):
        """Same on-disk layout as the reference (tri_cubic.py:89-99); needs h5py."""
        import h5py
        with h5py.File(filename, 'w') as f:
            f.create_dataset("TCI/xvec", data=self._xvec, dtype=np.double)
            f.create_dataset("TCI/yvec", data=self._yvec, dtype=np.double)
            f.create_dataset("TCI/zvec", data=self._zvec, dtype=np.double)
            f.create_dataset("TCI/M", data=_lib.host_f64(self._M), dtype=np.double)

    def get_model_coordinates(self):
        X, Y, Z = np.meshgrid(self._xvec, self._yvec, self._zvec, indexing='ij')
        return X.flatten(order='C'), Y.flatten(order='C'), Z.flatten(order='C')


def _simps_avg_torch(y, x):
    """``simps(y, x, axis=-1, even='avg')`` on torch tensors; 1-D ``x`` (host-side helper
    for ``TriCubic.inner`` only -- the ray integrals use the CUDA kernels)."""
    h = x[1:] - x[:-1]
    N = y.shape[-1]

    def basic(start, stop):
        h0 = h[start:stop:2]
        h1 = h[start + 1:stop + 1:2]
        hsum = h0 + h1
        return (hsum / 6.0 * (y[..., start:stop:2] * (2 - h1 / h0)
                              + y[..., start + 1:stop + 1:2] * hsum * hsum / (h0 * h1)
                              + y[..., start + 2:stop + 2:2] * (2 - h0 / h1))).sum(-1)
    if N % 2 == 0:
        val = 0.5 * h[-1] * (y[..., -1] + y[..., -2]) + 0.5 * h[0] * (y[..., 1] + y[..., 0])
        return (basic(0, N - 3) + basic(1, N - 2)) / 2.0 + val / 2.0
    return basic(0, N - 2)


def bisection(array, value):
    """Index ``j`` with ``array[j] <= value < array[j+1]``; -1 / ``len(array)`` when out of
    range; exact hits on the first / last node give 0 / n-1 (tri_cubic.py:105-132).
    Host-side scalar helper, ``np.searchsorted`` based."""
    array = np.asarray(array)
    n = len(array)
    if value < array[0]:
        return -1
    if value > array[n - 1]:
        return n
    if value == array[n - 1]:
        return n - 1
    return int(np.searchsorted(array, value, side='right') - 1)
