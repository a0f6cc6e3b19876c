"""Device-resident session for one inversion: fixed ray geometry + data, repeated
``(misfit, gradient)`` evaluations for changing models -- on one GPU or with the rays sharded over the
GPUs of a node (one process per GPU).

This is how the reference's drivers use the path: rays are computed ONCE per solve
(``inversion_pipeline.py:195-197``), then every iteration calls the forward and the gradient with a
new model (``tests/test_inversion.py:30-39`` ``func_and_gradient(m)``; ``bfgs_dask.py:207-340``;
``iterative_newton.py:954-1017``).  The session therefore assembles the prepared operator once
(``ForwardProjector``, applied in both directions; optionally the voxel-binned ``BackProjector``), keeps every
buffer of the step allocated, and replays the step as ONE CUDA graph:

    quad records of ne = K exp(m)/1e13      iono_forwardprojector_quads_from_m_f64 (touched records only)
    TEC per ray                             iono_forwardprojector_apply_quads_f64  (or the stateless sweep)
    dTEC, misfit, adjoint coefficients      iono_residual_f64
    gradient = ne * A^T coef                adjoint="binned":   iono_backprojector_ne_rows_f64 +
                                                                iono_backprojector_apply_permuted_f64 (bitwise
                                                                reproducible, second copy of the matrix in voxel order)
                                            adjoint="prepared": iono_forwardprojector_adjoint_f64 + finish_gradient
                                                                (the forward's records transposed, run-aggregated
                                                                reductions; reproducible to rounding)
                                            adjoint="scatter":  the stateless kernel on the rays

Sharded (``torch.distributed`` initialised, world size > 1; shard the direction or time axis so that the
reference antenna is local): the forward needs no exchange; the back-projector of every rank writes into a
COMPACT accumulator that numbers only the voxels some rank's rays touch, the misfit rides along as its last
element, and ``iono_peer_reduce_expand_f64`` sums, scales and expands it over NVLink peer memory inside the
same graph (``reducer="peer"``), or ``torch.distributed.all_reduce`` does between two graphs
(``reducer="nccl"``; gloo in the CPU tests of the host logic).
"""
import ctypes
import os

import torch
import torch.distributed as dist

from .. import _lib
from .. import sharding
from .forward_equation import ForwardProjector, TECU, ne_quads_from_m, quads_alloc, tec_from_quads
from .gradient import BackProjector, backproject, residual


class DeviceSession(object):
    def __init__(self, rays, K_ne, m_tci, i0, dobs, CdCt, forward="prepared", adjoint=None, order="time",
                 use_graph=True, keep_rays=None, check_bounds=True, group=None, reducer="peer", compact=None):
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
        if adjoint is None:
            # the transposed forward operator is the fastest adjoint and needs no second operator in HBM; "binned" is
            # the bitwise-reproducible one (and the one HostSession pipelines its downloads behind)
            adjoint = "prepared" if forward == "prepared" else "scatter"
        assert forward in ("prepared", "sweep") and adjoint in ("binned", "prepared", "scatter")
        assert adjoint != "prepared" or forward == "prepared", "adjoint='prepared' transposes the prepared forward"
        self.adjoint_kind = adjoint
        self.group = group
        self.rank, self.world = sharding.world() if group is None else (dist.get_rank(group), dist.get_world_size(group))
        # compact=True on a single process: the sharded step without the link (what one rank does; profiling)
        self.sharded = self.world > 1 or bool(compact)
        if self.world == 1 and self.sharded:
            reducer = "nccl"
        if self.sharded:
            assert adjoint in ("binned", "prepared"), "sharded rays: the compact accumulator needs a prepared adjoint"
            assert reducer in ("peer", "nccl")
        self.fp = ForwardProjector(self.rays, m_tci, check_bounds=check_bounds) if forward == "prepared" else None
        self.bp = BackProjector(self.rays, m_tci, check_bounds=check_bounds) if adjoint == "binned" else None
        both_prepared = self.fp is not None and (self.bp is not None or adjoint == "prepared")
        if keep_rays is None:
            keep_rays = not both_prepared
        if not keep_rays:
            assert both_prepared, "the stateless kernels read the rays"
            self.rays = None           # both operators are prepared: the 4 x Ns doubles per ray are not read again
        f64 = dict(dtype=torch.float64, device=dev)
        self.m = torch.empty(self.shape, **f64)                 # static input of the graph
        self.ne = torch.empty(self.shape, **f64) if adjoint == "scatter" else None
        self.ne_rows = torch.empty(self.shape, **f64) if (self.bp is not None and not self.sharded) else None
        # adjoint="prepared": full-grid accumulator of the reductions; zero between steps (finish_* clears what it reads)
        self.acc_full = torch.zeros(self.shape, **f64) if adjoint == "prepared" else None
        # quad records (4 x the grid): two 256-bit loads per sample instead of eight 64-bit ones
        V = self.shape[0] * self.shape[1] * self.shape[2]
        # prepared forward: only the records the rays read are rewritten per step, so the layout pays at any size
        # that fits (512x512x256: forward 2.86 -> 2.15 ms); the stateless sweep rewrites the whole grid per call
        self.use_quads = V <= ((1 << 27) if forward == "prepared" else (1 << 24))
        if os.environ.get("IONO_SESSION_QUADS") in ("0", "1"):      # measurement knob
            self.use_quads = os.environ["IONO_SESSION_QUADS"] == "1"
        self.quads = quads_alloc(self.shape, dev) if self.use_quads else None
        if not self.use_quads:
            if self.ne is None:
                self.ne = self.ne_rows if self.ne_rows is not None else torch.zeros(self.shape, **f64)
            elif self.ne_rows is not None:
                self.ne_rows = self.ne      # one buffer: the forward fills it, the apply scales with it
        self.tec = torch.empty(self.ray_shape, **f64)
        self.dtec = torch.empty(self.ray_shape, **f64)
        self.coef = torch.empty(self.ray_shape, **f64) if adjoint == "scatter" else None
        self.coef_perm = torch.empty(Na * Nt * Nd, **f64) if adjoint != "scatter" else None
        self.scratch = torch.empty(int(lib.iono_residual_scratch_elems(Na, Nt, Nd)), **f64)
        self.S = torch.zeros(1, **f64)
        self.S_local = self.S
        self.grad = torch.zeros(self.shape, **f64)     # voxels no ray touches stay zero for the whole session
        self.oob = torch.zeros(1, dtype=torch.int64, device=dev)
        self.use_graph = bool(use_graph)
        self._graphs = {}
        self.launches_per_call = {}
        self.n_forward = 0
        self.n_gradient = 0
        self.reducer = None
        self.reducer_kind = None
        if self.sharded:
            self._setup_sharded(reducer)

    # ---- sharded set-up: common numbering of the voxels any rank touches ---------------------------
    def _setup_sharded(self, reducer):
        lib = _lib.load()
        if self.bp is not None:
            n_rows = int(lib.iono_backprojector_n_rows(self.bp.handle))
            rows = torch.empty(max(n_rows, 1), dtype=torch.int32, device=self.device)
            _lib.call("iono_backprojector_row_voxels", self.bp.handle, ctypes.c_void_p(rows.data_ptr()),
                      _lib.stream_ptr())
            rows = rows[:n_rows]
        else:
            rows = self.fp.voxels()      # (a superset of the binned operator's rows: corners with weight exactly 0)
        V = self.shape[0] * self.shape[1] * self.shape[2]
        self.row_dst, self.union_voxels, self.n_union = sharding.union_index(rows, V, self.group)
        L = self.n_union + 1                                    # + the misfit
        self.reducer_kind = reducer
        if reducer == "peer":
            # peer memory needs CUDA IPC + peer access between all GPUs of the job; if any rank cannot map its peers
            # every rank falls back to the NCCL reduction (decided collectively, so nobody waits on a flag forever)
            from ..peer import PeerReducer
            ok = torch.ones(1, dtype=torch.int32, device=self.device)
            try:
                self.reducer = PeerReducer(L, self.group)
            except _lib.IonoError as exc:
                self.reducer = None
                self._peer_error = str(exc)
                ok.zero_()
            if self.world > 1:
                dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
            if int(ok.item()) == 0:
                if self.reducer is not None:
                    self.reducer = None          # (its buffers stay mapped until the process exits)
                reducer = self.reducer_kind = "nccl"
        if reducer == "peer":
            self.acc_c = self.reducer.acc_t
        else:
            from ..peer import LocalExpander
            self.reducer = LocalExpander(L, self.device)
            self.acc_c = self.reducer.acc_t
        self.S_local = self.acc_c[self.n_union:self.n_union + 1]   # the residual kernel writes the shard's misfit here

    # ---- the step, as enqueued work on the current stream ------------------------------------
    def _enqueue_forward(self):
        if not self.use_quads:
            # large grids: plain layout (IONO_FWD_LAYOUT policy of the C side keeps it plain beyond 2^24 voxels)
            from .forward_equation import tec_from_ne
            if self.fp is not None and self.bp is not None:
                _lib.call("iono_backprojector_ne_rows_f64", self.bp.handle, _lib.ptr(self.m), self.K_ne / TECU,
                          _lib.ptr(self.ne), _lib.stream_ptr())      # the operator's rows = every corner the rays read
                self.fp.tec(self.ne, out=self.tec)
            else:
                _lib.call("iono_ne_from_m_f64", _lib.ptr(self.m), self.m.numel(), self.K_ne / TECU, _lib.ptr(self.ne),
                          _lib.stream_ptr())
                if self.fp is not None:
                    self.fp.tec(self.ne, out=self.tec)
                else:
                    self.tec.copy_(tec_from_ne(self.rays, self.grid, self.ne, order=self.order, check_bounds=False))
            return
        if self.fp is not None:
            _lib.call("iono_forwardprojector_quads_from_m_f64", self.fp.handle, _lib.ptr(self.m), self.K_ne / TECU,
                      _lib.ptr(self.quads), _lib.stream_ptr())
            self.fp.tec_quads(self.quads, out=self.tec)
        else:
            ne_quads_from_m(self.m, self.K_ne, ne_out=self.ne, quads_out=self.quads, want_ne=self.ne is not None)
            tec_from_quads(self.rays, self.grid, self.quads, order=self.order, check_bounds=False, out=self.tec,
                           oob=self.oob)
        if self.fp is not None and self.ne is not None:       # prepared forward + scatter adjoint: the plain grid too
            _lib.call("iono_ne_from_m_f64", _lib.ptr(self.m), self.m.numel(), self.K_ne / TECU, _lib.ptr(self.ne),
                      _lib.stream_ptr())

    def _enqueue_residual(self):
        if self.sharded and self.reducer_kind == "nccl" and self.world > 1:
            # the in-place all_reduce leaves the SUM in this rank's accumulator: the rows only other ranks touch must
            # be cleared again.  (Peer reducer: the accumulator is only ever read by the peers; the apply overwrites
            # this rank's rows with plain stores and the others stay zero from the allocation.)
            _lib.call("iono_zero_f64", _lib.ptr(self.acc_c), self.acc_c.numel(), _lib.stream_ptr())
        residual(self.tec, self.dobs, self.CdCt, self.i0, want_coef=self.coef is not None,
                 want_perm=self.coef_perm is not None,
                 out=dict(dtec=self.dtec, coef=self.coef, coef_perm=self.coef_perm, scratch=self.scratch,
                          S=self.S_local))

    def _enqueue_adjoint(self):
        if self.adjoint_kind == "prepared":
            self.fp.adjoint(self.coef_perm, self.acc_full)
            if self.sharded:
                self.fp.finish_compact(self.acc_full, self.acc_c, self.row_dst)
            else:
                self.fp.finish_gradient(self.acc_full, self.m, self.K_ne / TECU, self.grad)
        elif self.sharded:
            _lib.call("iono_backprojector_apply_compact_f64", self.bp.handle, _lib.ptr(self.coef_perm),
                      ctypes.c_void_p(self.row_dst.data_ptr()), _lib.ptr(self.acc_c), 0, 16, _lib.stream_ptr())
        elif self.bp is not None:
            # chain-rule factor for the touched rows only, then the apply with `scale` (evaluating exp inside the
            # apply kernel costs 0.2 ms at the LOFAR case: it runs once per segment, not once per row)
            if self.use_quads:              # (plain layout: the forward has just filled ne for these rows)
                _lib.call("iono_backprojector_ne_rows_f64", self.bp.handle, _lib.ptr(self.m), self.K_ne / TECU,
                          _lib.ptr(self.ne_rows), _lib.stream_ptr())
            self.bp.apply_permuted(self.coef_perm, scale=self.ne_rows, out=self.grad)
        else:
            backproject(self.rays, self.grid, self.coef, self.shape, order=self.order, check_bounds=False,
                        out=self.grad)
            _lib.call("iono_mul_f64", _lib.ptr(self.ne), _lib.ptr(self.grad), self.grad.numel(), _lib.ptr(self.grad),
                      _lib.stream_ptr())

    def _enqueue_reduce(self):
        """Sharded: sum over ranks, chain-rule factor, expansion to the grid, summed misfit."""
        self.reducer.reduce_expand(self.union_voxels, self.n_union, self.m, self.K_ne / TECU, self.grad, self.S)

    def _run(self, key, fn):
        """Run ``fn`` (which only enqueues kernels on the current stream) eagerly the first time -- kernel
        attributes get set, lazy allocations happen -- then capture it once and replay the graph."""
        if not self.use_graph:
            return fn()
        g = self._graphs.get(key)
        if g is None:
            l0 = _lib.launch_count
            fn()                                    # first call: eager
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
        """``(dtec, S)``: the forward for model ``m`` (view of the session's buffer, overwritten by the next call;
        the local rays when sharded) and the misfit summed over all ranks (0-d CUDA tensor)."""
        self._set_model(m)
        self.n_forward += 1

        def fn():
            self._enqueue_forward()
            self._enqueue_residual()
        self._run("forward", fn)
        if self.sharded:
            self.S.copy_(self.S_local)
            if self.world > 1:
                dist.all_reduce(self.S, group=self.group)
        return self.dtec, self.S[0]

    def misfit_and_gradient(self, m=None):
        """``(S, grad)``: 0-d CUDA tensor and ``(nx,ny,nz)`` CUDA tensor, both global when sharded (the session's
        buffers; copy them if they must survive the next call).  ``self.dtec`` holds the forward of the local rays."""
        self._set_model(m)
        self.n_forward += 1
        self.n_gradient += 1
        if self.sharded and self.reducer_kind == "nccl":
            def fa():
                self._enqueue_forward()
                self._enqueue_residual()
                self._enqueue_adjoint()
            self._run("step_a", fa)
            if self.world > 1:
                dist.all_reduce(self.acc_c, group=self.group)
            self._run("step_b", self._enqueue_reduce)
        else:
            def fn():
                self._enqueue_forward()
                self._enqueue_residual()
                self._enqueue_adjoint()
                if self.sharded:
                    self._enqueue_reduce()
            self._run("step", fn)
        return self.S[0], self.grad

    def gradient_after_forward(self):
        """Gradient for the model of the last ``forward`` call (reuses its coefficients)."""
        self.n_gradient += 1
        if self.sharded and self.reducer_kind == "nccl":
            self._run("adjoint_a", self._enqueue_adjoint)
            if self.world > 1:
                dist.all_reduce(self.acc_c, group=self.group)
            self._run("step_b", self._enqueue_reduce)
        else:
            def fn():
                self._enqueue_adjoint()
                if self.sharded:
                    self._enqueue_reduce()
            self._run("adjoint", fn)
        return self.grad

    def active_voxels(self):
        """Flat int32 indices (device) of the voxels the gradient can be non-zero on -- the union over ranks when
        sharded -- or ``None`` when no prepared operator knows them (stateless kernels: the whole grid)."""
        if self.sharded:
            return self.union_voxels[:self.n_union]
        if self.bp is not None:
            nr = int(_lib.load().iono_backprojector_n_rows(self.bp.handle))
            idx = torch.empty(max(nr, 1), dtype=torch.int32, device=self.device)
            _lib.call("iono_backprojector_row_voxels", self.bp.handle, ctypes.c_void_p(idx.data_ptr()), _lib.stream_ptr())
            return idx[:nr].contiguous()
        if self.adjoint_kind == "prepared":
            return self.fp.voxels().contiguous()
        return None

    @property
    def operator_bytes(self):
        return (self.fp.nbytes if self.fp is not None else 0) + (self.bp.nbytes if self.bp is not None else 0)

    def close(self):
        """Release the peer mappings (collective when sharded with the peer reducer)."""
        self._graphs = {}
        if self.reducer is not None and hasattr(self.reducer, "close"):
            self.reducer.close()
            self.reducer = None
