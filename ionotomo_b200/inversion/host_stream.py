"""Forward + gradient for HOST-resident inputs, with the ray array streamed through the GPU.

The reference's callers hold ``rays`` as a NumPy array (``geometry/calc_rays.py:78-92``) and
call ``forward_equation`` then ``compute_gradient`` on it every iteration
(``tests/test_inversion.py:30-39``).  At the LOFAR scale that array is 5 GB, so a drop-in
call is bound by the host link.  ``misfit_and_gradient`` makes one pass: time blocks
``rays[:, t0:t1]`` are copied host -> device on a side stream (``iono_copy2d_h2d``) into a
double buffer while the previous block is integrated and back-projected.  A time block is
self-contained for the exact adjoint because the reference-antenna coupling
(``tec - tec[i0]``, ``forward_equation.py:50``) only links rays of the same (time, direction).
"""
import ctypes

import numpy as np
import torch

from .. import _lib
from .forward_equation import _ne_from_m, tec_from_ne
from .gradient import adjoint_coefficients


def pin(a):
    """A pinned float64 torch view of a NumPy array (copies once into pinned memory)."""
    t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64))
    return t if t.is_pinned() else t.pin_memory()


_pinned_out = {}
_staging = {}


def _pinned_like(key, shape):
    """Cached pinned host buffer for results (device -> host copies at full PCIe speed)."""
    t = _pinned_out.get(key)
    if t is None or tuple(t.shape) != tuple(shape):
        t = torch.empty(tuple(shape), dtype=torch.float64, pin_memory=True)
        _pinned_out[key] = t
    return t


def misfit_and_gradient(rays, K_ne, m_tci, i0, dobs, CdCt, order="time", block_times=None, check_bounds=True,
                        reduce_fn=None, timings=None, copy_results=True, reduce_scalar=None):
    """``(dtec, S, gradient)`` as NumPy/float from host arrays.

    ``rays`` may be a NumPy array or a (preferably pinned) CPU tensor of shape
    ``(Na, Nt, Nd, 4, Ns)``.  Equivalent to ``forward_equation`` + misfit + ``compute_gradient``
    of this package (and to the reference's ``func_and_gradient`` sketch,
    tests/test_inversion.py:30-39, with the exact adjoint).

    Results are read back into cached *pinned* host buffers.  With ``copy_results=False`` the
    returned arrays are views of those buffers (valid until the next call) -- what an optimiser
    loop wants; the default returns private copies.

    With sharded rays pass ``reduce_fn`` (sum of the voxel accumulator over ranks, applied in place
    to a CUDA tensor) AND ``reduce_scalar`` (0-d CUDA tensor -> float summed over ranks, e.g.
    ``sharding.sharded_misfit``): the returned ``(S, gradient)`` pair is then the global one on
    every rank; ``dtec`` always covers the local rays only.  Without ``reduce_scalar`` the misfit
    is the shard-local one.
    """
    import time as _time
    _t0 = _time.time()
    _lib.require_cuda()
    lib = _lib.load()
    rays_h = rays if isinstance(rays, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(rays, dtype=np.float64))
    assert rays_h.device.type == "cpu" and rays_h.dtype == torch.float64 and rays_h.is_contiguous()
    Na, Nt, Nd, four, Ns = rays_h.shape
    dev = torch.device("cuda", torch.cuda.current_device())
    row_bytes = Nd * 4 * Ns * 8
    if block_times is None:   # ~256 MB blocks
        block_times = max(1, min(Nt, (256 << 20) // max(1, Na * row_bytes)))
    main = torch.cuda.current_stream()
    # staging buffers, copy stream and events are cached across calls: allocating and freeing
    # 2 x 256 MB per pass (cudaMalloc/cudaFree synchronise) costs more than the compute
    key = (dev.index, Na * block_times * Nd * 4 * Ns)
    st = _staging.get(key)
    if st is None:
        st = {"side": torch.cuda.Stream(),
              "bufs": [torch.empty(key[1], dtype=torch.float64, device=dev) for _ in range(2)],
              "ready": [torch.cuda.Event() for _ in range(2)], "free": [torch.cuda.Event() for _ in range(2)]}
        _staging.clear()
        _staging[key] = st
    side, bufs, ready, free = st["side"], st["bufs"], st["ready"], st["free"]
    side.wait_stream(main)
    blocks = [(t0, min(Nt, t0 + block_times)) for t0 in range(0, Nt, block_times)]

    def upload(b):
        t0, t1 = blocks[b]
        w = (t1 - t0) * row_bytes
        with torch.cuda.stream(side):
            if b >= 2:
                side.wait_event(free[b % 2])
            src = rays_h.data_ptr() + t0 * row_bytes
            _lib.call("iono_copy2d_h2d", ctypes.c_void_p(bufs[b % 2].data_ptr()), w, ctypes.c_void_p(src),
                      Nt * row_bytes, w, Na, ctypes.c_void_p(side.cuda_stream))
            ready[b % 2].record(side)

    m_dev = m_tci.device_M()
    ne = _ne_from_m(m_dev, K_ne)
    dobs_d, C_d = _lib.to_device(dobs), _lib.to_device(CdCt)
    dtec = torch.empty((Na, Nt, Nd), dtype=torch.float64, device=dev)
    acc = torch.zeros(tuple(m_dev.shape), dtype=torch.float64, device=dev)
    oob = torch.zeros(1, dtype=torch.int64, device=dev)
    oob_tot = torch.zeros(1, dtype=torch.int64, device=dev)
    grid = m_tci.grid()
    if timings is not None:
        torch.cuda.synchronize()
        timings["setup_ms"] = (_time.time() - _t0) * 1e3
        _t0 = _time.time()
    if blocks:
        upload(0)
    for b, (t0, t1) in enumerate(blocks):
        if b + 1 < len(blocks):
            upload(b + 1)
        main.wait_event(ready[b % 2])
        tb = t1 - t0
        # the 2-D copy packs the block densely, also the last, shorter one
        rb = bufs[b % 2][:Na * tb * Nd * 4 * Ns].reshape(Na, tb, Nd, 4, Ns)
        tec = tec_from_ne(rb, grid, ne, order=order, check_bounds=False)
        d = torch.empty_like(tec)
        _lib.call("iono_dtec_f64", _lib.ptr(tec), Na, tb, Nd, int(i0), _lib.ptr(d), _lib.stream_ptr())
        dtec[:, t0:t1] = d
        coef = adjoint_coefficients(d, dobs_d[:, t0:t1].contiguous(), C_d[:, t0:t1].contiguous(), i0)
        _lib.call("iono_tec_adjoint_f64", grid.handle, _lib.ptr(rb), Na, tb, Nd, Ns, _lib.ptr(coef),
                  _lib.ORDERS[order], 0, _lib.ptr(acc), ctypes.c_void_p(oob.data_ptr()), _lib.stream_ptr())
        oob_tot += oob
        free[b % 2].record(main)
    if timings is not None:
        timings["enqueue_ms"] = (_time.time() - _t0) * 1e3
        torch.cuda.synchronize()
        timings["stream_ms"] = (_time.time() - _t0) * 1e3
        _t0 = _time.time()
    if reduce_fn is not None:
        acc = reduce_fn(acc)
    from .gradient import misfit
    S = misfit(dtec, dobs_d, C_d)
    _lib.call("iono_mul_f64", _lib.ptr(ne), _lib.ptr(acc), acc.numel(), _lib.ptr(acc), _lib.stream_ptr())
    if check_bounds and int(oob_tot.item()) != 0:
        raise ValueError("One of the requested xi is out of bounds (%d ray samples outside the grid)"
                         % int(oob_tot.item()))
    main.wait_stream(side)
    dtec_h = _pinned_like("dtec", dtec.shape)
    grad_h = _pinned_like("grad", acc.shape)
    dtec_h.copy_(dtec, non_blocking=True)
    grad_h.copy_(acc, non_blocking=True)
    S = float(reduce_scalar(S)) if reduce_scalar is not None else float(S)   # synchronises
    torch.cuda.current_stream().synchronize()
    if timings is not None:
        timings["finish_ms"] = (_time.time() - _t0) * 1e3
    if copy_results:   # private copies, so that the next call cannot overwrite them
        return dtec_h.numpy().copy(), S, grad_h.numpy().copy()
    return dtec_h.numpy(), S, grad_h.numpy()


class HostSession(object):
    """Host-level iteration API: geometry, data and the prepared operators stay on the GPU for the whole solve;
    per call only the model goes in and ``(dtec, S, gradient)`` come out.

    The reference computes its rays once per solve (``inversion_pipeline.py:195-197``) and then calls
    ``func_and_gradient(m)`` every iteration (``tests/test_inversion.py:30-39``); this is that call for host
    (NumPy) callers.  ``rays`` may be the materialised host array, or ``None`` with ``origins`` / ``directions``
    (``(Na,Nt,Nd,3)``) given: the rays are then generated on the device (``cast_ray``), 60 MB of input
    instead of 5 GB.  All arrays returned are views of pinned host buffers owned by the session, valid until
    the next call; ``m_host`` is a pinned buffer the caller may fill in place (pass ``m=None`` then).

    Adjoint (``adjoint=`` keyword of the session, default chosen here): full-grid gradients on one GPU use the
    voxel-ordered ``"binned"`` operator, because its pieces finish the gradient in voxel order and the 67 MB download
    rides behind the kernel (4.2 ms per call at the LOFAR case; 4.9 ms with the faster ``"prepared"`` adjoint, whose
    gradient is only complete at the end); ``active_only`` and multi-GPU sessions use ``"prepared"``.

    ``active_only=True``: the model and the gradient travel as vectors over the ACTIVE voxels
    (``self.active_voxels``, flat indices of the voxels some ray touches; the gradient is identically zero
    elsewhere, so an optimiser only ever changes those entries of ``m``): a fifth of the bytes at the LOFAR case.
    The session keeps the full model on the device, initialised from ``m_tci.M``.

    With ``torch.distributed`` initialised and the rays sharded over the ranks (direction or time blocks) the
    HOST program is rank ``root``'s: its model is broadcast to the other GPUs over NVLink, every rank returns the
    ``dtec`` of its own rays, and ``S`` and the gradient (global) are copied to the host on ``root`` only (the other
    ranks return ``None`` for the gradient) -- one model upload and one gradient download per step for the whole
    job, whatever the number of GPUs.
    """

    def __init__(self, rays, K_ne, m_tci, i0, dobs, CdCt, origins=None, directions=None, tmax=1000., Ns=None,
                 active_only=False, root=0, overlap_copies=True, **session_kw):
        import ctypes
        import torch.distributed as dist
        from ..geometry.calc_rays import cast_ray
        from .fermat import Fermat
        from .session import DeviceSession
        _lib.require_cuda()
        if rays is None:
            assert origins is not None and directions is not None
            rays = cast_ray((_lib.to_device(origins), _lib.to_device(directions)), Fermat(m_tci), tmax,
                            Ns if Ns is not None else m_tci.nz)
        if session_kw.get("adjoint") is None and session_kw.get("forward", "prepared") == "prepared":
            # full-grid gradients on one GPU: the voxel-binned operator finishes the grid in voxel order, which lets the
            # 67 MB download run behind the kernel piece by piece (_pipelined_step); otherwise the faster transposed
            # forward operator
            import torch.distributed as _d
            one_gpu = not (_d.is_available() and _d.is_initialized() and _d.get_world_size() > 1)
            session_kw["adjoint"] = "binned" if (one_gpu and not active_only and overlap_copies) else "prepared"
        self.session = DeviceSession(rays, K_ne, m_tci, i0, dobs, CdCt, **session_kw)
        s = self.session
        self.root = int(root)
        self.is_root = (not s.sharded) or s.world == 1 or s.rank == self.root
        self._dist = dist if (s.sharded and s.world > 1) else None
        self.active_only = bool(active_only)
        self.active_voxels = None
        if self.active_only:
            idx = s.active_voxels()
            assert idx is not None, "active_only needs a prepared adjoint (it knows the voxels the rays touch)"
            self._idx = idx
            self.active_voxels = idx.cpu().numpy().astype(np.int64)
            n = int(idx.numel())
            s.m.copy_(m_tci.device_M())
            self._m_act = torch.empty(n, dtype=torch.float64, device=s.device)
            self._g_act = torch.empty(n, dtype=torch.float64, device=s.device)
            self._one = torch.ones(1, dtype=torch.float64, device=s.device)
            self._zero_grid = None
            self.m_host = torch.empty(n, dtype=torch.float64, pin_memory=True)
            self.m_host.copy_(s.m.reshape(-1)[idx.long()])
            self.grad_host = torch.empty(n, dtype=torch.float64, pin_memory=True)
        else:
            self.m_host = torch.empty(s.shape, dtype=torch.float64, pin_memory=True)
            self.grad_host = torch.empty(s.shape, dtype=torch.float64, pin_memory=True)
        self.dtec_host = torch.empty(s.ray_shape, dtype=torch.float64, pin_memory=True)
        self.S_host = torch.empty(1, dtype=torch.float64, pin_memory=True)
        # one GPU: copies overlapped with the kernels (see _pipelined_step)
        self.overlap_copies = bool(overlap_copies)
        self.n_pieces = 4
        self._side = torch.cuda.Stream()
        self._grad_full_host = self.grad_host
        if self.active_only and not s.sharded and s.bp is not None:
            # active voxels whose gradient is final after each piece of the operator
            bounds = [0] + [s.bp.chunk_voxels(c) for c in range(16 // self.n_pieces, 16, 16 // self.n_pieces)] + \
                [s.grad.numel()]
            self._act_bounds = np.searchsorted(self.active_voxels, np.array(bounds), side="left")
        self.h2d_bytes_per_call = self.m_host.numel() * 8 if self.is_root else 0
        self.d2h_bytes_per_call = (self.dtec_host.numel() + (self.grad_host.numel() + 1 if self.is_root else 0)) * 8

    def _upload(self, m):
        import ctypes
        s = self.session
        if self.is_root:
            if m is not None:
                src = torch.as_tensor(m, dtype=torch.float64).reshape(self.m_host.shape)
                if src.data_ptr() != self.m_host.data_ptr():
                    self.m_host.copy_(src)
            if self.active_only:
                self._m_act.copy_(self.m_host, non_blocking=True)
            else:
                s.m.copy_(self.m_host, non_blocking=True)
        if self._dist is not None:                       # the root's model reaches the other GPUs over NVLink
            self._dist.broadcast(self._m_act if self.active_only else s.m, src=self.root, group=s.group)
        if self.active_only:
            # m[active] = m_act
            _lib.call("iono_scatter_set_f64", _lib.ptr(self._m_act), ctypes.c_void_p(self._idx.data_ptr()),
                          self._m_act.numel(), _lib.ptr(s.m), _lib.stream_ptr())

    def misfit_and_gradient(self, m=None):
        """``(dtec, S, gradient)`` for the model ``m`` (NumPy ``(nx,ny,nz)``, or the vector over
        ``active_voxels`` with ``active_only``; ``None``: ``self.m_host`` as filled by the caller)."""
        import ctypes
        s = self.session
        self._upload(m)
        if self.overlap_copies and not s.sharded and s.bp is not None:
            return self._pipelined_step()
        s.misfit_and_gradient(None)
        self.dtec_host.copy_(s.dtec, non_blocking=True)
        if self.is_root:
            if self.active_only:
                _lib.call("iono_gather_f64", _lib.ptr(s.grad), ctypes.c_void_p(self._idx.data_ptr()),
                          self._g_act.numel(), _lib.ptr(self._g_act), _lib.stream_ptr())
                self.grad_host.copy_(self._g_act, non_blocking=True)
            else:
                self.grad_host.copy_(s.grad, non_blocking=True)
        self.S_host.copy_(s.S, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self.dtec_host.numpy(), float(self.S_host[0]), (self.grad_host.numpy() if self.is_root else None)

    def _pipelined_step(self):
        """One GPU, binned adjoint: the device -> host copies ride behind the kernels.  dTEC and the misfit leave
        on a side stream while the back-projection runs; the operator is applied in ``self.n_pieces`` pieces of
        its voxel-sorted entry stream (``iono_backprojector_apply_*`` chunks) and every finished voxel range of
        the gradient is copied out while the next piece is computed.  Same numbers as the one-shot step."""
        import ctypes
        s = self.session
        main = torch.cuda.current_stream()
        side = self._side
        s.forward(None)                                  # quad records, forward, dTEC + misfit + coefficients (graph)
        ev = torch.cuda.Event()
        ev.record(main)
        with torch.cuda.stream(side):
            side.wait_event(ev)
            self.dtec_host.copy_(s.dtec, non_blocking=True)
            self.S_host.copy_(s.S, non_blocking=True)
        s.n_gradient += 1
        if s.use_quads:
            _lib.call("iono_backprojector_ne_rows_f64", s.bp.handle, _lib.ptr(s.m), s.K_ne / 1e13, _lib.ptr(s.ne_rows),
                      _lib.stream_ptr())
        V = s.grad.numel()
        flat_d, flat_h = s.grad.view(-1), self._grad_full_host.view(-1)
        step = 16 // self.n_pieces
        done = 0
        for c0 in range(0, 16, step):
            s.bp.apply_permuted(s.coef_perm, scale=s.ne_rows, out=s.grad, c0=c0, c1=c0 + step)
            upto = V if c0 + step == 16 else s.bp.chunk_voxels(c0 + step)
            if upto > done:
                ev = torch.cuda.Event()
                ev.record(main)
                with torch.cuda.stream(side):
                    side.wait_event(ev)
                    if self.active_only:
                        a, b = int(self._act_bounds[c0 // step]), int(self._act_bounds[c0 // step + 1])
                        if b > a:
                            _lib.call("iono_gather_f64", _lib.ptr(s.grad), ctypes.c_void_p(self._idx[a:b].data_ptr()),
                                      b - a, _lib.ptr(self._g_act[a:b]), ctypes.c_void_p(side.cuda_stream))
                            self.grad_host[a:b].copy_(self._g_act[a:b], non_blocking=True)
                    else:
                        flat_h[done:upto].copy_(flat_d[done:upto], non_blocking=True)
                done = upto
        side.synchronize()
        main.synchronize()
        return self.dtec_host.numpy(), float(self.S_host[0]), self.grad_host.numpy()

    def forward(self, m=None):
        """``(dtec, S)`` only (line searches)."""
        s = self.session
        self._upload(m)
        s.forward(None)
        self.dtec_host.copy_(s.dtec, non_blocking=True)
        self.S_host.copy_(s.S, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self.dtec_host.numpy(), float(self.S_host[0])

    def close(self):
        self.session.close()
