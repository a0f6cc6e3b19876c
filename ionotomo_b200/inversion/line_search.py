"""Line search along the negative gradient: ``line_search`` / ``vertex`` of
``inversion/line_search.py:13-100`` (secant estimate of the step from one probe at
``ep=1e-3``, halving until the misfit drops and at least three evaluations, parabola vertex
through the last three).  Every forward runs on the GPU; only scalars come back."""
import numpy as np
import torch

from .. import _lib
from ..geometry.tri_cubic import TriCubic
from .forward_equation import forward_equation
from .gradient import misfit


def vertex(x1, x2, x3, y1, y2, y3):
    """Vertex (x, y) of the parabola through three points (line_search.py:13-43)."""
    denom = (x1 - x2) * (x1 - x3) * (x2 - x3)
    A = (x3 * (y2 - y1) + x2 * (y1 - y3) + x1 * (y3 - y2)) / denom
    B = (x3 * x3 * (y1 - y2) + x2 * x2 * (y3 - y1) + x1 * x1 * (y2 - y3)) / denom
    C = (x2 * x3 * (x2 - x3) * y1 + x3 * x1 * (x3 - x1) * y2 + x1 * x2 * (x1 - x2) * y3) / denom
    return -B / (2 * A), C - B * B / (4 * A)


def line_search(rays, K_ne, m_tci, i0, gradient, g, dobs, CdCt, figname=None, order="time", verbose=False):
    """Returns ``(epsilon_n, S, S/S0 - 1)`` like the reference (line_search.py:45-100)."""
    rays_d = _lib.to_device(rays)
    M = m_tci.device_M()
    grad = _lib.to_device(gradient).reshape(M.shape)
    g_d, dobs_d, C_d = _lib.to_device(g), _lib.to_device(dobs), _lib.to_device(CdCt)
    xv, yv, zv = m_tci.xvec, m_tci.yvec, m_tci.zvec

    def fwd(eps):
        return forward_equation(rays_d, K_ne, TriCubic(xv, yv, zv, M - eps * grad), i0, order=order)

    S0 = float(misfit(g_d, dobs_d, C_d))
    dd = (g_d - dobs_d) / (C_d + 1e-15)
    ep = 1e-3
    Gm = (g_d - fwd(ep)) / ep
    numerator = 2. * float((dd * Gm).sum())
    denominator = float((Gm * Gm / (C_d + 1e-15)).sum())
    epsilon_n = abs(numerator / denominator)
    epsilon_n0 = epsilon_n
    ep_a, S_a = [], []
    S = S0
    it = 0
    while S >= S0 or it < 3:
        epsilon_n /= 2.
        S = float(misfit(fwd(epsilon_n), dobs_d, C_d))
        ep_a.append(epsilon_n)
        S_a.append(S)
        if not np.isnan(S):
            it += 1
        if len(ep_a) > 200:
            break
    epsilon_n, S_p = vertex(*ep_a[-3:], *S_a[-3:])
    S = float(misfit(fwd(epsilon_n), dobs_d, C_d))
    if verbose:
        print("S0: {} | Estimated epsilon_n: {}".format(S0, epsilon_n0))
        print("Parabolic minimum | epsilon_n = {}, S = {}".format(epsilon_n, S_p))
        print("Actual | S = {}".format(S))
        print("Misfit Reduction: {:.2f}%".format(S / S0 * 100. - 100.))
    return epsilon_n, S, (S / S0 - 1.)
