"""Device-resident session for one inversion: fixed ray geometry + data, repeated
``(misfit, gradient)`` evaluations for changing models.

This is how the reference's drivers use the path: rays are computed ONCE per solve
(``inversion_pipeline.py:195-197``), then every iteration calls the forward and the gradient with a
new model (``tests/test_inversion.py:30-39`` ``func_and_gradient(m)``; ``bfgs_dask.py:207-340``;
``iterative_newton.py:954-1017``).  The session therefore assembles the two prepared operators once
(``ForwardProjector``, ``BackProjector``), keeps every buffer of the step allocated, and replays
the step as ONE CUDA graph:

    quad records of ne = K exp(m)/1e13      iono_ne_quads_from_m_f64
    TEC per ray                             iono_forwardprojector_apply_quads_f64 (or the stateless sweep)
    dTEC, misfit, adjoint coefficients      iono_residual_f64
    gradient = ne * A^T coef                iono_backprojector_apply_permuted_f64 (or the scatter adjoint)

With rays sharded over ranks pass ``reduce_grad`` / ``reduce_scalar`` (see ``ionotomo_b200.sharding``);
collectives are not captured, the graph is then split around them.
"""
import ctypes

import torch

from .. import _lib
from .forward_equation import ForwardProjector, ne_quads_from_m, quads_alloc, tec_from_quads
from .gradient import BackProjector, backproject, residual


class DeviceSession(object):
    def __init__(self, rays, K_ne, m_tci, i0, dobs, CdCt, forward="prepared", adjoint="binned", order="time",
                 use_graph=True, keep_rays=None, check_bounds=True):
        lib = _lib.load()
        self.rays = _lib.to_device(rays)
        Na, Nt, Nd, four, Ns = self.rays.shape
        assert four == 4
        self.ray_shape = (Na, Nt, Nd)
        self.Ns = Ns
        self.K_ne = float(K_ne)
        self.i0 = int(i0)
        self.order = order
        self.grid = m_tci.grid()
        self.shape = (m_tci.nx, m_tci.ny, m_tci.nz)
        dev = self.rays.device
        self.device = dev
        self.dobs = _lib.to_device(dobs).reshape(self.ray_shape).contiguous()
        self.CdCt = _lib.to_device(CdCt).reshape(self.ray_shape).contiguous()
        assert forward in ("prepared", "sweep") and adjoint in ("binned", "scatter")
        self.fp = ForwardProjector(self.rays, m_tci, check_bounds=check_bounds) if forward == "prepared" else None
        self.bp = BackProjector(self.rays, m_tci, check_bounds=check_bounds) if adjoint == "binned" else None
        if keep_rays is None:
            keep_rays = not (self.fp is not None and self.bp is not None)
        if not keep_rays:
            assert self.fp is not None and self.bp is not None, "the stateless kernels read the rays"
            self.rays = None           # both operators are prepared: the 4 x Ns doubles per ray are not read again
        f64 = dict(dtype=torch.float64, device=dev)
        self.m = torch.empty(self.shape, **f64)                 # static input of the graph
        self.ne = torch.empty(self.shape, **f64)
        self.quads = quads_alloc(self.shape, dev)
        self.tec = torch.empty(self.ray_shape, **f64)
        self.dtec = torch.empty(self.ray_shape, **f64)
        self.coef = torch.empty(self.ray_shape, **f64) if self.bp is None else None
        self.coef_perm = torch.empty(Na * Nt * Nd, **f64) if self.bp is not None else None
        self.scratch = torch.empty(int(lib.iono_residual_scratch_elems()), **f64)
        self.S = torch.zeros(1, **f64)
        self.grad = torch.empty(self.shape, **f64)
        self.oob = torch.zeros(1, dtype=torch.int64, device=dev)
        self.use_graph = bool(use_graph)
        self._graphs = {}
        self.launches_per_call = {}
        self.n_forward = 0
        self.n_gradient = 0

    # ---- the step, as enqueued work on the current stream ------------------------------------
    def _enqueue_forward(self):
        ne_quads_from_m(self.m, self.K_ne, ne_out=self.ne, quads_out=self.quads)
        if self.fp is not None:
            self.fp.tec_quads(self.quads, out=self.tec)
        else:
            tec_from_quads(self.rays, self.grid, self.quads, order=self.order, check_bounds=False, out=self.tec,
                           oob=self.oob)

    def _enqueue_residual(self):
        residual(self.tec, self.dobs, self.CdCt, self.i0, want_coef=self.bp is None, want_perm=self.bp is not None,
                 out=dict(dtec=self.dtec, coef=self.coef, coef_perm=self.coef_perm, scratch=self.scratch, S=self.S))

    def _enqueue_adjoint(self):
        if self.bp is not None:
            self.bp.apply_permuted(self.coef_perm, scale=self.ne, out=self.grad)
        else:
            backproject(self.rays, self.grid, self.coef, self.shape, order=self.order, check_bounds=False,
                        out=self.grad)
            _lib.call("iono_mul_f64", _lib.ptr(self.ne), _lib.ptr(self.grad), self.grad.numel(), _lib.ptr(self.grad),
                      _lib.stream_ptr())

    def _run(self, key, fn):
        """Run ``fn`` (which only enqueues kernels on the current stream) eagerly the first time -- kernel
        attributes get set, lazy allocations happen -- then capture it once and replay the graph."""
        if not self.use_graph:
            return fn()
        g = self._graphs.get(key)
        if g is None:
            l0 = _lib.launch_count
            fn()                                    # warm-up, eager
            self.launches_per_call[key] = _lib.launch_count - l0
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
            _lib.launch_count -= self.launches_per_call[key]      # the capture launched nothing
            self._graphs[key] = g
            return
        g.replay()
        _lib.launch_count += self.launches_per_call[key]

    def _set_model(self, m):
        if m is not None and m.data_ptr() != self.m.data_ptr():
            self.m.copy_(_lib.to_device(m).reshape(self.shape), non_blocking=True)

    # ---- public ------------------------------------------------------------------------------
    def forward(self, m=None):
        """``dtec`` (view of the session's buffer, overwritten by the next call) and misfit for model ``m``."""
        self._set_model(m)
        self.n_forward += 1

        def fn():
            self._enqueue_forward()
            self._enqueue_residual()
        self._run("forward", fn)
        return self.dtec, self.S[0]

    def misfit_and_gradient(self, m=None):
        """``(S, grad)``: 0-d CUDA tensor and ``(nx,ny,nz)`` CUDA tensor (the session's buffers; copy them if
        they must survive the next call).  ``self.dtec`` holds the forward."""
        self._set_model(m)
        self.n_forward += 1
        self.n_gradient += 1

        def fn():
            self._enqueue_forward()
            self._enqueue_residual()
            self._enqueue_adjoint()
        self._run("step", fn)
        return self.S[0], self.grad

    def gradient_after_forward(self):
        """Gradient for the model of the last ``forward`` call (reuses its ne and coefficients)."""
        self.n_gradient += 1
        self._run("adjoint", self._enqueue_adjoint)
        return self.grad

    @property
    def operator_bytes(self):
        return (self.fp.nbytes if self.fp is not None else 0) + (self.bp.nbytes if self.bp is not None else 0)
