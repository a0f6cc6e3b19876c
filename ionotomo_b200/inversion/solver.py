"""Device-resident inversion driver: L-BFGS on the log-density model ``m`` with the misfit
``S(m) = sum((g(m) - dobs)^2 / (CdCt + 1e-15)) / 2`` -- the loop the reference sketches in
``tests/test_inversion.py:30-39`` (``fmin_l_bfgs_b(func_and_gradient, m0, ...)``) and builds as a
dask graph in ``inversion/bfgs_dask.py:207-340``.

Every forward and every gradient is one CUDA-graph replay of a ``DeviceSession``; the optimiser's own
algebra runs in the library's vector kernels on the ACTIVE voxels only (the ones some ray touches --
the gradient is identically zero elsewhere), with the history kept as one matrix:

* L-BFGS in the compact form of Byrd, Nocedal & Schnabel (1994): ``H g = gamma g + S u + gamma Y v`` with
  ``u, v`` from the small matrices ``S^T Y``, ``Y^T Y`` and the products ``S^T g``, ``Y^T g`` -- all
  products of an iteration come from three ``iono_multi_dot_f64`` passes over the history and return to
  the host in ONE copy per iteration; the direction is one ``iono_lincomb_f64`` pass.  (The reference
  evaluates the same recursion one scalar product and one axpy per dask task, bfgs_dask.py:165-194.)
* ``metric="simpson"`` makes every product the reference's inner product (``scalarProduct`` =
  triple Simpson integral over the grid, bfgs_dask.py:165-167 / ``TriCubic.inner``); default Euclidean.
* The first step length comes from the reference's secant probe (``line_search.py:55-66``: one extra
  forward at ``ep = 1e-3``); later iterations start from the unit step with Armijo backtracking.

With sharded rays (``DeviceSession`` under ``torch.distributed``) misfit and gradient are global on every
rank, so every rank takes identical steps.
"""
import ctypes

import numpy as np
import torch

from .. import _lib
from ..geometry.tri_cubic import TriCubic
from .forward_equation import ForwardProjector, forward_equation
from .gradient import BackProjector, adjoint_coefficients, backproject, misfit, _ne_from_m
from .session import DeviceSession


class InversionProblem(object):
    """Fixed geometry + data; evaluates misfit and gradient for a model array on the device.
    (Kernel-by-kernel form kept for callers of round 1's API; ``lbfgs_solve`` drives a ``DeviceSession``.)"""

    def __init__(self, rays, K_ne, m_tci, i0, dobs, CdCt, order="time", binned=True, reduce_fn=None,
                 reduce_scalar=None, prepared=False):
        self.rays = _lib.to_device(rays)
        self.K_ne = float(K_ne)
        self.i0 = int(i0)
        self.xvec, self.yvec, self.zvec = m_tci.xvec, m_tci.yvec, m_tci.zvec
        self.grid = m_tci.grid()
        self.dobs = _lib.to_device(dobs)
        self.CdCt = _lib.to_device(CdCt)
        self.order = order
        self.reduce_fn = reduce_fn
        self.reduce_scalar = reduce_scalar
        self.bp = BackProjector(self.rays, m_tci) if binned else None
        self.fp = ForwardProjector(self.rays, m_tci) if prepared else None
        self.n_forward = 0
        self.n_gradient = 0

    def _tci(self, m):
        t = TriCubic.__new__(TriCubic)
        t._xvec, t._yvec, t._zvec = self.xvec, self.yvec, self.zvec
        t.nx, t.ny, t.nz = len(self.xvec), len(self.yvec), len(self.zvec)
        t._grid = self.grid
        t._M = m
        return t

    def forward(self, m):
        self.n_forward += 1
        return forward_equation(self.rays, self.K_ne, self._tci(m), self.i0, order=self.order, check_bounds=False,
                                projector=self.fp)

    def misfit(self, g):
        S = misfit(g, self.dobs, self.CdCt)
        if self.reduce_scalar is not None:
            return self.reduce_scalar(S)
        return float(S)

    def gradient(self, m, g):
        """dS/dm given the forward ``g = forward(m)`` (exact adjoint)."""
        self.n_gradient += 1
        coef = adjoint_coefficients(g, self.dobs, self.CdCt, self.i0)
        ne = _ne_from_m(m, self.K_ne)
        if self.bp is not None:
            grad = self.bp.apply(coef, scale=ne)
        else:
            grad = backproject(self.rays, self.grid, coef, tuple(m.shape), order=self.order, check_bounds=False)
            grad *= ne
        if self.reduce_fn is not None:
            grad = self.reduce_fn(grad)
        return grad


def simpson_grid_weights(xvec, yvec, zvec):
    """Weights ``w`` with ``sum(w * a * b) == TriCubic.inner``-style triple ``simps`` of ``a*b`` over the grid
    (geometry/tri_cubic.py:61-67, bfgs_dask.py:165-167): outer product of the 1-D old-SciPy ``even='avg'``
    weights of the three axes."""
    def w1(x):
        x = np.asarray(x, dtype=np.float64)
        n = x.size
        eye = np.eye(n)
        return np.array([_simps_avg_1d(eye[i], x) for i in range(n)])
    return np.einsum("i,j,k->ijk", w1(xvec), w1(yvec), w1(zvec))


def _simps_avg_1d(y, x):
    """Old ``scipy.integrate.simps(y, x, even='avg')`` for one 1-D array (tomography/integrate.py:50-153)."""
    def basic(y, x, start, stop):
        s = 0.0
        for a in range(start, stop, 2):
            h0, h1 = x[a + 1] - x[a], x[a + 2] - x[a + 1]
            s += (h0 + h1) / 6. * (y[a] * (2 - h1 / h0) + y[a + 1] * (h0 + h1) ** 2 / (h0 * h1) + y[a + 2] * (2 - h0 / h1))
        return s
    n = len(y)
    if n < 2:
        return 0.0
    if n % 2 == 1:
        return basic(y, x, 0, n - 2)
    first = basic(y, x, 0, n - 3) + 0.5 * (x[-1] - x[-2]) * (y[-1] + y[-2])
    last = basic(y, x, 1, n - 2) + 0.5 * (x[1] - x[0]) * (y[1] + y[0])
    return 0.5 * (first + last)


class _Vectors(object):
    """History matrix + work rows on the active voxels, and the kernels that act on them.
    Rows: s in slots 0..k-1, y in slots k..2k-1 (a pair occupies physical slot p and k+p; pairs are kept in a ring,
    ``order`` lists the slots from oldest to newest), then G (current gradient), D (direction), T (new gradient)."""

    def __init__(self, n, history, device, weights=None):
        lib = _lib.load()
        self.n, self.k = int(n), int(history)
        assert 2 * self.k + 3 <= 32, "history <= 14"
        self.ld = (self.n + 3) // 4 * 4
        self.H = torch.zeros((2 * self.k + 3, self.ld), dtype=torch.float64, device=device)
        self.G, self.D, self.T = 2 * self.k, 2 * self.k + 1, 2 * self.k + 2
        self.scratch = torch.empty(int(lib.iono_multi_dot_scratch_elems()), dtype=torch.float64, device=device)
        self.dots = torch.zeros((3, 32), dtype=torch.float64, device=device)
        self.dots_h = torch.zeros((3, 32), dtype=torch.float64).pin_memory()
        # coefficient vectors go up asynchronously from pinned slots; a ring, so that a slot is not rewritten while its
        # copy may still be in flight (an iteration issues <= 4 of them between two host synchronisations)
        self.coef_h = torch.zeros((8, 40), dtype=torch.float64).pin_memory()
        self.coef = torch.zeros((8, 40), dtype=torch.float64, device=device)
        self._slot = 0
        self.w = weights

    def row(self, r):
        return self.H[r]

    def _w(self, weighted=True):
        return _lib.ptr(self.w) if (self.w is not None and weighted) else None

    def multi_dot(self, rows, x_row, slot, weighted=True):
        """dots[slot, r] = <H[r], H[x_row]> for r < rows."""
        _lib.call("iono_multi_dot_f64", _lib.ptr(self.H), self.ld, int(rows), _lib.ptr(self.H[x_row]),
                  self._w(weighted), self.n, _lib.ptr(self.scratch), _lib.ptr(self.dots[slot]), _lib.stream_ptr())

    def multi_dot3(self, rows, x0, x1, x2):
        """dots[j, r] = <H[r], H[x_j]> for r < rows, j = 0, 1, 2: one pass over the history."""
        _lib.call("iono_multi_dot3_f64", _lib.ptr(self.H), self.ld, int(rows), int(x0), int(x1), int(x2), self._w(),
                  self.n, _lib.ptr(self.scratch), _lib.ptr(self.dots), _lib.stream_ptr())

    def fetch_dots(self):
        self.dots_h.copy_(self.dots, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self.dots_h.numpy()

    def lincomb(self, coefs, x_row, out_row, rows=None, first_row=0):
        """H[out_row] = coefs[0] * H[x_row] + sum_r coefs[r+1] * H[first_row + r]."""
        c = np.asarray(coefs, dtype=np.float64)
        rows = len(c) - 1 if rows is None else rows
        k = self._slot
        self._slot = (k + 1) % 8
        self.coef_h[k, :len(c)] = torch.from_numpy(c)
        self.coef[k].copy_(self.coef_h[k], non_blocking=True)
        _lib.call("iono_lincomb_f64", _lib.ptr(self.H[first_row]) if rows > 0 else None, self.ld, int(rows),
                  _lib.ptr(self.coef[k]), _lib.ptr(self.H[x_row]) if x_row is not None else None, self.n,
                  _lib.ptr(self.H[out_row]), _lib.stream_ptr())


def lbfgs_solve(problem, m0, n_iter=50, history=10, c1=1e-4, max_backtracks=20, callback=None, metric=None):
    """Minimise the misfit from ``m0`` (float64 ``(nx, ny, nz)``, NumPy or CUDA tensor).

    ``problem``: a ``DeviceSession`` (or an ``InversionProblem`` of round 1's API, which is wrapped).
    Returns ``(m, info)`` with ``info['S']`` the misfit history (``n_iter + 1`` values at most),
    ``info['n_forward']``, ``info['n_gradient']``, ``info['host_syncs_per_iteration']``, ``info['iter_seconds']``
    (wall time of every iteration; the first ones include the one-off CUDA-graph captures of the session).
    ``metric``: ``None`` (Euclidean) or ``"simpson"`` (the reference's grid inner product).
    """
    import time as _time
    ses = problem if isinstance(problem, DeviceSession) else _session_from_problem(problem)
    dev = ses.device
    m = _lib.to_device(m0).reshape(ses.shape).clone()
    # active voxels: rows of the operator (union over ranks when sharded)
    idx = ses.active_voxels()
    if idx is None:
        idx = torch.arange(m.numel(), dtype=torch.int32, device=dev)
    n = int(idx.numel())
    w = None
    if metric == "simpson":
        xv, yv, zv = ses.grid.xvec, ses.grid.yvec, ses.grid.zvec
        w = torch.as_tensor(simpson_grid_weights(xv, yv, zv)).to(dev).reshape(-1)[idx.long()].contiguous()
    elif metric is not None:
        raise ValueError("metric must be None or 'simpson'")
    V = _Vectors(n, history, dev, w)
    k = history
    step_dev = torch.zeros(1, dtype=torch.float64, device=dev)
    step_pin = torch.zeros(1, dtype=torch.float64).pin_memory()
    ses.m.copy_(m)

    def gather(src, row):
        _lib.call("iono_gather_f64", _lib.ptr(src), ctypes.c_void_p(idx.data_ptr()), n, _lib.ptr(V.row(row)),
                  _lib.stream_ptr())

    def trial(step):
        """Forward at m + step * d, written straight into the session's model buffer; returns the misfit
        (the one host synchronisation of a trial)."""
        step_pin[0] = step
        step_dev.copy_(step_pin, non_blocking=True)
        _lib.call("iono_scatter_axpy_f64", _lib.ptr(m), _lib.ptr(step_dev), _lib.ptr(V.row(V.D)),
                  ctypes.c_void_p(idx.data_ptr()), n, _lib.ptr(ses.m), _lib.stream_ptr())
        dtec, S_t = ses.forward(None)
        return float(S_t)

    S, grad = ses.misfit_and_gradient(None)
    S = float(S)
    gather(grad, V.G)
    S_hist = [S]
    order = []                 # physical slots of the stored pairs, oldest first
    SY = np.zeros((k, k))      # SY[p, q] = <s_p, y_q>, YY[p, q] = <y_p, y_q>  (physical slots)
    YY = np.zeros((k, k))
    step0 = None
    syncs, iter_seconds = [], []
    t_prev = _time.time()
    V.multi_dot(2 * k + 1, V.G, 0)       # <every row, g> incl. <g, g>
    d = V.fetch_dots()
    for it in range(n_iter):
        n_sync = 0
        c = len(order)
        o = np.array(order, dtype=int)
        p1, p2, gg = d[0, o].copy(), d[0, k + o].copy(), float(d[0, V.G])
        use_history = c > 0
        if use_history:
            SYc, YYc = SY[np.ix_(o, o)], YY[np.ix_(o, o)]
            R = np.triu(SYc)
            Dg = np.diag(np.diag(SYc))
            gamma = SYc[-1, -1] / YYc[-1, -1]
            try:
                Rinv_p1 = np.linalg.solve(R, p1)
                u = np.linalg.solve(R.T, (Dg + gamma * YYc) @ Rinv_p1 - gamma * p2)
                v = -Rinv_p1
                gd = -(gamma * gg + u @ p1 + gamma * (v @ p2))
            except np.linalg.LinAlgError:
                use_history = False
            if use_history and not (gd < 0 and np.isfinite(gd)):
                use_history = False            # not a descent direction: restart with steepest descent
        if use_history:
            coefs = np.zeros(2 * k + 1)
            coefs[0] = -gamma
            coefs[1 + o] = -u
            coefs[1 + k + o] = -gamma * v
            V.lincomb(coefs, V.G, V.D, rows=2 * k)
            step = 1.0
        else:
            order = []
            V.lincomb([-1.0], V.G, V.D, rows=0)
            gd = -gg
            step = step0
        if w is not None:                       # the Armijo test needs the Euclidean directional derivative
            V.multi_dot(V.G + 1, V.D, 1, weighted=False)
            gd = float(V.fetch_dots()[1, V.G])
            n_sync += 1
        if step is None:
            # secant estimate of the step along -grad from one probe (reference line_search.py:55-66)
            gmax = float(V.row(V.G)[:n].abs().max())
            ep = 1e-3 / max(gmax, 1e-300)
            g0 = ses.dtec.clone()
            trial(ep)
            Gm = (ses.dtec - g0) / ep
            dd = (g0 - ses.dobs) / (ses.CdCt + 1e-15)
            nd = torch.stack([torch.sum(dd * Gm), torch.sum(Gm * Gm / (ses.CdCt + 1e-15))])
            if ses.sharded and ses.world > 1:
                torch.distributed.all_reduce(nd, group=ses.group)
            num, den = float(nd[0]), float(nd[1])
            step = step0 = abs(num / den) if den > 0 else 1.0
            n_sync += 3
        accepted = False
        for _ in range(max_backtracks):
            S_new = trial(step)
            n_sync += 1
            if S_new == S_new and S_new <= S + c1 * step * gd:
                accepted = True
                break
            step *= 0.5
        if not accepted:
            ses.m.copy_(m)
            break
        grad = ses.gradient_after_forward()
        gather(grad, V.T)
        # new pair in a free slot (or in the oldest pair's): s = step * d, y = g_new - g
        if len(order) == k:
            p = order.pop(0)
        else:
            p = [q for q in range(k) if q not in order][0]
        V.lincomb([step], V.D, p, rows=0)
        V.lincomb([1.0, -1.0], V.T, k + p, rows=1, first_row=V.G)
        V.H[V.G].copy_(V.H[V.T])
        # m += step * d on the active voxels (the same arithmetic that produced the accepted trial model)
        _lib.call("iono_scatter_axpy_f64", _lib.ptr(m), _lib.ptr(step_dev), _lib.ptr(V.row(V.D)),
                  ctypes.c_void_p(idx.data_ptr()), n, _lib.ptr(m), _lib.stream_ptr())
        S = S_new
        S_hist.append(S)
        # every product the next iteration needs in one pass over the history and ONE copy to the host:
        #   dots[0] = <rows, g_new>   dots[1] = <rows, y_new>   dots[2] = <rows, s_new>
        V.multi_dot3(2 * k + 1, V.G, k + p, p)
        d = V.fetch_dots()
        n_sync += 1
        sy, yy = d[1, p], d[1, k + p]
        if sy > 1e-12 * yy:
            SY[:, p] = d[1, :k]                 # <s_q, y_new>
            SY[p, :] = d[2, k:2 * k]            # <s_new, y_q>
            YY[:, p] = d[1, k:2 * k]
            YY[p, :] = d[1, k:2 * k]
            order.append(p)
        syncs.append(n_sync)
        t_now = _time.time()                    # (fetch_dots above synchronised the stream)
        iter_seconds.append(t_now - t_prev)
        t_prev = t_now
        if callback is not None:
            callback(it, m, S)
    return m, {"S": S_hist, "n_forward": ses.n_forward, "n_gradient": ses.n_gradient, "active_voxels": n,
               "host_syncs_per_iteration": float(np.mean(syncs)) if syncs else 0.0, "iter_seconds": iter_seconds}


def _session_from_problem(problem):
    """Round 1's ``InversionProblem`` -> session on the same rays and data."""
    tci = problem._tci(torch.zeros((len(problem.xvec), len(problem.yvec), len(problem.zvec)), dtype=torch.float64,
                                   device=problem.rays.device))
    ses = DeviceSession(problem.rays, problem.K_ne, tci, problem.i0, problem.dobs, problem.CdCt,
                        forward="prepared" if problem.fp is not None else "sweep",
                        adjoint="binned" if problem.bp is not None else "scatter", order=problem.order)
    return ses
