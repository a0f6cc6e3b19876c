"""Device-resident inversion driver: L-BFGS on the log-density model ``m`` with the misfit
``S(m) = sum((g(m) - dobs)^2 / (CdCt + 1e-15)) / 2`` -- the loop the reference sketches in
``tests/test_inversion.py:30-39`` (``fmin_l_bfgs_b(func_and_gradient, m0, ...)``) and builds as a
dask graph in ``inversion/bfgs_dask.py:207-340``; every forward and every gradient runs on the
GPU and nothing but scalars returns to the host between iterations.

The first step length comes from the reference's secant probe (``line_search.py:55-66``: one
extra forward at ``ep = 1e-3``); later iterations start from the L-BFGS-scaled unit step with
Armijo backtracking.  With sharded rays pass ``reduce_fn`` / ``reduce_scalar`` (see
``ionotomo_b200.sharding``): the search direction is then identical on every rank.
"""
import torch

from .. import _lib
from ..geometry.tri_cubic import TriCubic
from .forward_equation import ForwardProjector, forward_equation
from .gradient import BackProjector, adjoint_coefficients, backproject, misfit, _ne_from_m


class InversionProblem(object):
    """Fixed geometry + data; evaluates misfit and gradient for a model array on the device."""

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


def lbfgs_solve(problem, m0, n_iter=50, history=10, c1=1e-4, max_backtracks=20, callback=None):
    """Minimise the misfit from ``m0`` (float64 CUDA tensor ``(nx, ny, nz)``).

    Returns ``(m, info)`` with ``info['S']`` the misfit history (``n_iter + 1`` values at most),
    ``info['n_forward']``, ``info['n_gradient']``.
    """
    m = _lib.to_device(m0).clone()
    g = problem.forward(m)
    S = problem.misfit(g)
    grad = problem.gradient(m, g)
    hist_s, hist_y, hist_rho = [], [], []
    S_hist = [S]
    step0 = None
    for it in range(n_iter):
        # two-loop recursion
        q = grad.clone()
        alphas = []
        for s, y, rho in zip(reversed(hist_s), reversed(hist_y), reversed(hist_rho)):
            a = rho * float(torch.sum(s * q))
            alphas.append(a)
            q.add_(y, alpha=-a)
        if hist_s:
            gamma = float(torch.sum(hist_s[-1] * hist_y[-1]) / torch.sum(hist_y[-1] * hist_y[-1]))
            q.mul_(gamma)
        for (s, y, rho), a in zip(zip(hist_s, hist_y, hist_rho), reversed(alphas)):
            b = rho * float(torch.sum(y * q))
            q.add_(s, alpha=a - b)
        d = -q
        gd = float(torch.sum(grad * d))
        if not gd < 0:          # not a descent direction: restart with steepest descent
            hist_s, hist_y, hist_rho = [], [], []
            d = -grad
            gd = float(torch.sum(grad * d))
        if not hist_s:
            # secant estimate of the step along -grad from one probe (reference line_search.py:55-66)
            if step0 is None:
                ep = 1e-3 / max(float(grad.abs().max()), 1e-300)
                g_probe = problem.forward(m + ep * d)
                Gm = (g_probe - g) / ep
                dd = (g - problem.dobs) / (problem.CdCt + 1e-15)
                num = float(torch.sum(dd * Gm))
                den = float(torch.sum(Gm * Gm / (problem.CdCt + 1e-15)))
                if problem.reduce_scalar is not None:
                    num = problem.reduce_scalar(torch.tensor(num, dtype=torch.float64, device=m.device))
                    den = problem.reduce_scalar(torch.tensor(den, dtype=torch.float64, device=m.device))
                step0 = abs(num / den) if den > 0 else 1.0
            step = step0
        else:
            step = 1.0
        accepted = False
        for _ in range(max_backtracks):
            m_new = m + step * d
            g_new = problem.forward(m_new)
            S_new = problem.misfit(g_new)
            if S_new == S_new and S_new <= S + c1 * step * gd:
                accepted = True
                break
            step *= 0.5
        if not accepted:
            break
        grad_new = problem.gradient(m_new, g_new)
        s_vec = m_new - m
        y_vec = grad_new - grad
        sy = float(torch.sum(s_vec * y_vec))
        if sy > 1e-12 * float(torch.sum(y_vec * y_vec)):
            hist_s.append(s_vec)
            hist_y.append(y_vec)
            hist_rho.append(1.0 / sy)
            if len(hist_s) > history:
                hist_s.pop(0); hist_y.pop(0); hist_rho.pop(0)
        m, g, S, grad = m_new, g_new, S_new, grad_new
        S_hist.append(S)
        if callback is not None:
            callback(it, m, S)
    return m, {"S": S_hist, "n_forward": problem.n_forward, "n_gradient": problem.n_gradient}
