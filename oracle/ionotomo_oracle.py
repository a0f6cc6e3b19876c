"""CPU oracle for the IonoTomo ray-integral forward model and its adjoint.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it; nothing under ``ionotomo_b200/`` does.

It is a plain NumPy (fp64) restatement of the reference's algorithm for the hot
path.  Every function cites the reference file:line it follows (paths relative
to ``/root/reference/src/ionotomo/``).  The arithmetic the reference delegates
to SciPy (un-pinned: ``/root/reference/pip-requirements.txt:2``) is restated
from SciPy's published algorithms:

* ``scipy.interpolate.RegularGridInterpolator(method='linear')`` -> `rgi_linear`
* ``scipy.integrate.simps(y, x, even='avg')`` (removed from SciPy >= 1.14; the
  reference's own port is ``tomography/integrate.py:50-153``) -> `simps_avg`
* ``scipy.integrate.odeint`` on the straight-ray RHS (``inversion/fermat.py:48-84``)
  -> closed form in `integrate_ray_straight`

Parity pinning (see ``tests/golden/make_golden.py`` and ``tests/test_oracle.py``):
the oracle is checked against outputs of the *reference's own modules* executed
in the build container (with stub modules for the absent astropy/h5py/dask and a
NumPy shim of the TF1 ops used by ``tomography/integrate.py``), committed as
``tests/golden/*.npz``, against the installed SciPy where the algorithm is still
shipped (RGI; ``simpson`` for odd N; ``ndimage.convolve``), and -- for the even-N
``'avg'`` rule, which no installed library implements any more -- against the known
answers published in old SciPy's own ``simps`` docstring (1642.5 / 1644.5 / 40.5).
"""
from __future__ import annotations

import numpy as np

TECU = 1e13  # "TEC unit / km": inversion/forward_equation.py:12, gradient.py:14
SPEED_OF_LIGHT = 299792458.0  # inversion/iterative_newton.py:15


# --------------------------------------------------------------------------
# quadrature: scipy.integrate.simps(y, x, even='avg'), last axis
# --------------------------------------------------------------------------
def _basic_simps(y, start, stop, x):
    """Composite Simpson over triples starting at start, start+2, ... < stop.

    Non-uniform-x form; follows tomography/integrate.py:50-74 (the reference's
    port of SciPy's ``_basic_simps``), same operation order.
    """
    h = np.diff(x, axis=-1)
    h0 = h[..., start:stop:2]
    h1 = h[..., start + 1:stop + 1:2]
    hsum = h0 + h1
    hprod = h0 * h1
    h0divh1 = h0 / h1
    tmp = hsum / 6.0 * (y[..., start:stop:2] * (2 - 1.0 / h0divh1)
                        + y[..., start + 1:stop + 1:2] * hsum * hsum / hprod
                        + y[..., start + 2:stop + 2:2] * (2 - h0divh1))
    return np.sum(tmp, axis=-1)


def simps_avg(y, x):
    """``simps(y, x, axis=-1, even='avg')`` as the reference ran it.

    tomography/integrate.py:76-153; call sites inversion/forward_equation.py:28,
    inversion/iterative_newton.py:119,179.  ``x`` has the shape of ``y`` or (N,).
    """
    y = np.asarray(y, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1 and y.ndim > 1:
        x = np.broadcast_to(x, y.shape)
    N = y.shape[-1]
    if N % 2 == 0:
        val = 0.0
        result = 0.0
        # 'first': Simpson on the first N-2 intervals, trapezoid on the last
        last_dx = x[..., -1] - x[..., -2]
        val = val + 0.5 * last_dx * (y[..., -1] + y[..., -2])
        result = _basic_simps(y, 0, N - 3, x)
        # 'last': trapezoid on the first interval, Simpson on the last N-2
        first_dx = x[..., 1] - x[..., 0]
        val = val + 0.5 * first_dx * (y[..., 1] + y[..., 0])
        result = result + _basic_simps(y, 1, N - 2, x)
        val = val / 2.0
        result = result / 2.0
        return result + val
    return _basic_simps(y, 0, N - 2, x)


def simps_weights(x):
    """Weight vector w with ``simps_avg(y, x) == sum(w*y)`` (rule is linear in y).

    Used by the exact-transpose adjoint.
    Built by applying `simps_avg` to unit vectors so it is the same rule by
    construction.
    """
    x = np.asarray(x, dtype=np.float64)
    N = x.shape[-1]
    eye = np.eye(N)
    w = np.empty(x.shape, dtype=np.float64)
    flat_x = x.reshape(-1, N)
    flat_w = w.reshape(-1, N)
    for r in range(flat_x.shape[0]):
        flat_w[r] = simps_avg(eye, flat_x[r])
    return w


def simps_weights_fast(x):
    """Closed-form Simpson-avg weights, vectorised over leading axes.

    Same rule as `simps_weights` (asserted equal in tests/test_oracle.py) but
    O(N) per ray, for the adjoint oracle at larger sizes.
    """
    x = np.asarray(x, dtype=np.float64)
    N = x.shape[-1]
    h = np.diff(x, axis=-1)
    w = np.zeros_like(x)

    def add_triples(start, stop, scale):
        if stop <= start:      # N < 3 (or N = 2,3 shifted): no triple
            return
        h0 = h[..., start:stop:2]
        h1 = h[..., start + 1:stop + 1:2]
        hsum = h0 + h1
        c = hsum / 6.0
        w[..., start:stop:2] += scale * c * (2 - h1 / h0)
        w[..., start + 1:stop + 1:2] += scale * c * hsum * hsum / (h0 * h1)
        w[..., start + 2:stop + 2:2] += scale * c * (2 - h0 / h1)

    if N % 2 == 0:
        add_triples(0, N - 3, 0.5)
        add_triples(1, N - 2, 0.5)
        w[..., -1] += 0.25 * h[..., -1]
        w[..., -2] += 0.25 * h[..., -1]
        w[..., 0] += 0.25 * h[..., 0]
        w[..., 1] += 0.25 * h[..., 0]
    else:
        add_triples(0, N - 2, 1.0)
    return w


# --------------------------------------------------------------------------
# grid interpolation: TriCubic.interp == scipy RGI(method='linear')
# --------------------------------------------------------------------------
def find_indices(grid, x):
    """Per-axis (i, t) of SciPy's RGI: SURVEY Appendix A.1.

    ``i = clip(searchsorted(grid, x, 'right') - 1, 0, n-2)``,
    ``t = (x - g[i]) / (g[i+1] - g[i])``.  Reference restatement:
    tomography/interpolation.py:145-195.
    """
    grid = np.asarray(grid, dtype=np.float64)
    i = np.searchsorted(grid, x, side='right') - 1
    i = np.clip(i, 0, grid.size - 2)
    t = (x - grid[i]) / (grid[i + 1] - grid[i])
    return i, t


def rgi_linear(xvec, yvec, zvec, M, x, y, z, bounds_error=True):
    """Trilinear sample of ``M`` at points (x, y, z).

    geometry/tri_cubic.py:22,59,69-70 (``interp``, bounds_error=True raises
    ValueError) and :71-75 (``extrapolate``: bounds_error=False,
    fill_value=None keeps the clipped cell and lets t leave [0,1]).
    Corner order and weight product order follow SciPy's ``_evaluate_linear``:
    corners 000,001,...,111 (z fastest), weight ((wx*wy)*wz).
    """
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    z = np.asarray(z, dtype=np.float64)
    M = np.asarray(M, dtype=np.float64)
    if bounds_error:
        for d, (g, p) in enumerate(((xvec, x), (yvec, y), (zvec, z))):
            if not np.logical_and(np.all(g[0] <= p), np.all(p <= g[-1])):
                raise ValueError(
                    "One of the requested xi is out of bounds in dimension %d" % d)
    ix, tx = find_indices(xvec, x)
    iy, ty = find_indices(yvec, y)
    iz, tz = find_indices(zvec, z)
    out = np.zeros(x.shape, dtype=np.float64)
    for cx in (0, 1):
        wx = tx if cx else 1 - tx
        for cy in (0, 1):
            wy = ty if cy else 1 - ty
            for cz in (0, 1):
                wz = tz if cz else 1 - tz
                out = out + M[ix + cx, iy + cy, iz + cz] * ((wx * wy) * wz)
    return out


def bisection(array, value):
    """Scalar binary search of geometry/tri_cubic.py:105-132 (same branches)."""
    n = len(array)
    if value < array[0]:
        return -1
    elif value > array[n - 1]:
        return n
    jl = 0
    ju = n - 1
    while ju - jl > 1:
        jm = (ju + jl) >> 1
        if value >= array[jm]:
            jl = jm
        else:
            ju = jm
    if value == array[0]:
        return 0
    elif value == array[n - 1]:
        return n - 1
    return jl


def tci_inner(xvec, yvec, zvec, A, B):
    """``TriCubic.inner``: triple simps of A*B (geometry/tri_cubic.py:61-67)."""
    P = A * B
    return simps_avg(simps_avg(simps_avg(P, zvec), yvec), xvec)


# --------------------------------------------------------------------------
# ray generation: Fermat.integrate_ray (straight, type 'z') + cast_ray
# --------------------------------------------------------------------------
def integrate_ray_straight(origin, direction, tmax, N):
    """Closed form of ``Fermat.integrate_ray`` for the shipped straight case.

    inversion/fermat.py:150-174 with ``euler_ode`` :48-84 at n=1, grad n=0
    (type 'z'): xdot=px/pz, ydot=py/pz, zdot=1, sdot=1/pz; independent variable
    z on ``linspace(z0, tmax, N)``.
    """
    x0, y0, z0 = origin
    xd, yd, zd = direction
    sdot = np.sqrt(xd ** 2 + yd ** 2 + zd ** 2)
    px, py, pz = xd / sdot, yd / sdot, zd / sdot
    z = np.linspace(z0, tmax, N)
    dzs = z - z0
    x = x0 + (px / pz) * dzs
    y = y0 + (py / pz) * dzs
    s = dzs / pz
    return x, y, z, s


def integrate_ray_arclength(origin, direction, smax, N):
    """``Fermat(type='s').integrate_ray`` for straight rays (inversion/fermat.py:74-82, :163-166): the
    independent variable is the arc length, ``s = linspace(0, smax, N)``, position ``= origin + p s``."""
    o = np.asarray(origin, dtype=np.float64)
    d = np.asarray(direction, dtype=np.float64)
    p = d / np.sqrt(d[0] ** 2 + d[1] ** 2 + d[2] ** 2)
    s = np.linspace(0., smax, N)
    return o[0] + p[0] * s, o[1] + p[1] * s, o[2] + p[2] * s, s


def cast_ray(origins, directions, tmax, N):
    """Vectorised ``cast_ray`` (geometry/calc_rays.py:61-96): (Na,Nt,Nd,3) ->
    rays (Na,Nt,Nd,4,N), rows x,y,z,s.  Same formulas as `integrate_ray_straight`."""
    origins = np.asarray(origins, dtype=np.float64)
    directions = np.asarray(directions, dtype=np.float64)
    sdot = np.sqrt(directions[..., 0] ** 2 + directions[..., 1] ** 2
                   + directions[..., 2] ** 2)
    px = directions[..., 0] / sdot
    py = directions[..., 1] / sdot
    pz = directions[..., 2] / sdot
    z0 = origins[..., 2]
    # np.linspace(z0, tmax, N) broadcast over rays: i*step + z0, last := tmax
    step = (tmax - z0) / (N - 1)
    idx = np.arange(N, dtype=np.float64)
    z = idx * step[..., None] + z0[..., None]
    z[..., -1] = tmax
    dzs = z - z0[..., None]
    rays = np.empty(origins.shape[:-1] + (4, N), dtype=np.float64)
    rays[..., 0, :] = origins[..., 0, None] + (px / pz)[..., None] * dzs
    rays[..., 1, :] = origins[..., 1, None] + (py / pz)[..., None] * dzs
    rays[..., 2, :] = z
    rays[..., 3, :] = dzs / pz[..., None]
    return rays


def ne2n(ne, frequency):
    """Fermat.ne2n (inversion/fermat.py:36-46)."""
    M = np.array(ne, dtype=np.float64, copy=True)
    M *= -8.980 ** 2 / frequency ** 2
    M += 1.
    np.sqrt(M, out=M)
    return M


def optical_path(ray, xvec, yvec, zvec, n_field):
    """s row of one ray (4, Ns) for ``straight_line_approx=False`` as shipped: the ODE
    ``sdot = n/pz`` with straight geometry (inversion/fermat.py:48-84).  The reference integrates
    with LSODA; here every sample interval is split at its grid-plane crossings and integrated
    with 2-point Gauss-Legendre, exact for the (piecewise cubic) trilinear interpolant."""
    Ns = ray.shape[1]
    x, y, z = ray[0], ray[1], ray[2]
    pz = (z[-1] - z[0]) / (ray[3, -1] - ray[3, 0])
    gl = 0.5 / np.sqrt(3.)
    s = np.zeros(Ns)
    for i in range(Ns - 1):
        za, zb = z[i], z[i + 1]
        kx, ky = (x[i + 1] - x[i]) / (zb - za), (y[i + 1] - y[i]) / (zb - za)
        cuts = [za, zb]
        cuts += [g for g in zvec if za < g < zb]
        for k, g0, vec in ((kx, x[i], xvec), (ky, y[i], yvec)):
            if k != 0:
                zc = za + (np.asarray(vec) - g0) / k
                cuts += [c for c in zc if za < c < zb]
        cuts = np.unique(cuts)
        seg = 0.
        for c0, c1 in zip(cuts[:-1], cuts[1:]):
            h, zm = c1 - c0, 0.5 * (c0 + c1)
            zq = np.array([zm - gl * h, zm + gl * h])
            nq = rgi_linear(xvec, yvec, zvec, n_field, x[i] + kx * (zq - za), y[i] + ky * (zq - za), zq)
            seg += 0.5 * h * nq.sum()
        s[i + 1] = s[i] + seg
    return s / pz


def pointing_rotation(lon_rad, ha_rad, dec_rad):
    """R = [east; north; up] of the reference's Pointing frame
    (astro/frames/pointing_frame.py:151-166): lonrad = lon - HA, latrad = dec."""
    lonrad = lon_rad - ha_rad
    sinlat, coslat = np.sin(dec_rad), np.cos(dec_rad)
    sinlon, coslon = np.sin(lonrad), np.cos(lonrad)
    north = [-sinlat * coslon, -sinlat * sinlon, coslat]
    east = [-sinlon, coslon, 0.]
    up = [coslat * coslon, coslat * sinlon, sinlat]
    return np.array([east, north, up])


def cast_ray_frames(ants_itrs_m, p0_itrs_m, R, dirs_itrs, tmax, N):
    """calc_rays' per-time loop (geometry/calc_rays.py:125-139) with the Pointing transform
    written out (pointing_frame.py:168-183): origins = R_t (p - p0) in km, directions = R_t d."""
    ants = np.asarray(ants_itrs_m, dtype=np.float64)
    Na, Nt, Nd = ants.shape[0], R.shape[0], dirs_itrs.shape[1]
    origins = np.zeros((Na, Nt, Nd, 3))
    directions = np.zeros((Na, Nt, Nd, 3))
    diff = ants - np.asarray(p0_itrs_m, dtype=np.float64)
    for j in range(Nt):
        o = np.stack([R[j, r, 0] * diff[:, 0] + R[j, r, 1] * diff[:, 1] + R[j, r, 2] * diff[:, 2]
                      for r in range(3)], -1) / 1000.0
        d = np.stack([R[j, r, 0] * dirs_itrs[j, :, 0] + R[j, r, 1] * dirs_itrs[j, :, 1]
                      + R[j, r, 2] * dirs_itrs[j, :, 2] for r in range(3)], -1)
        origins[:, j, :, :] += o[:, None, :]
        directions[:, j, :, :] += d[None]
    return cast_ray(origins, directions, tmax, N)


# --------------------------------------------------------------------------
# TEC forward (generation A): inversion/forward_equation.py
# --------------------------------------------------------------------------
def ne_from_m(m, K_ne):
    """``ne = exp(m) * (K_ne/TECU)`` per voxel: inversion/forward_equation.py:41-43."""
    ne = np.exp(np.asarray(m, dtype=np.float64))
    ne *= K_ne / TECU
    return ne


def tec(rays, xvec, yvec, zvec, ne, bounds_error=True):
    """Per-ray ``simps(interp(ne; x,y,z), s)``: do_forward_equation,
    inversion/forward_equation.py:13-33 (batched like iterative_newton.py:108,119)."""
    nevec = rgi_linear(xvec, yvec, zvec, ne, rays[..., 0, :], rays[..., 1, :],
                       rays[..., 2, :], bounds_error=bounds_error)
    return simps_avg(nevec, rays[..., 3, :])


def forward_equation(rays, K_ne, xvec, yvec, zvec, m, i0):
    """dTEC (Na,Nt,Nd): inversion/forward_equation.py:36-51."""
    t = tec(rays, xvec, yvec, zvec, ne_from_m(m, K_ne))
    return t - t[i0, :, :]


# --------------------------------------------------------------------------
# phase forward (generation B): inversion/iterative_newton.py:86-127
# --------------------------------------------------------------------------
def _interp_batched(xvec, yvec, zvec, M, rays, reference_axis_scramble):
    """``tci.interp(rays[...,0,:], rays[...,1,:], rays[...,2,:])`` on 4-D inputs.

    geometry/tri_cubic.py:69-70 evaluates ``rgi(np.array([x,y,z]).T)`` and then
    ``np.reshape(..., np.shape(x))``.  For 1-D inputs (generation A, one ray at a
    time) that is the identity; for the (Na,Nt,Nd,Ns) inputs of generation B
    (iterative_newton.py:108,157,160) ``.T`` reverses ALL axes and the reshape
    does not undo it, so the reference returns the samples in a scrambled order
    (element [a,t,d,s] is taken from the C-order flattening of the
    (Ns,Nd,Nt,Na) array).  ``reference_axis_scramble=True`` reproduces that
    bit-for-bit (used only to pin the oracle against the golden vectors); the
    default is the evidently intended per-ray ordering.
    """
    v = rgi_linear(xvec, yvec, zvec, M, rays[..., 0, :], rays[..., 1, :], rays[..., 2, :])
    if reference_axis_scramble:
        v = np.reshape(np.transpose(v), v.shape)
    return v


def phase_forward_equation(mu, clock, const, xvec, yvec, zvec, rays, freqs,
                           K=1e11, i0=0, reference_axis_scramble=False):
    """Phase (Na,Nt,Nd,Nf): inversion/iterative_newton.py:86-127, same order."""
    Na, Nt, Nd, _, Ns = rays.shape
    freqs = np.asarray(freqs, dtype=np.float64)
    Nf = freqs.shape[0]
    ne = np.exp(mu).reshape(len(xvec), len(yvec), len(zvec)) * K
    g = np.einsum("i,j,k,l,i->ijkl", np.ones(Na), np.ones(Nt), np.ones(Nd),
                  np.ones(Nf), const)
    ne_rays = _interp_batched(xvec, yvec, zvec, ne, rays, reference_axis_scramble)
    for l in range(Nf):
        a_ = 2 * np.pi * freqs[l]
        dg = a_ * np.einsum("ij,k->ijk", clock, np.ones(Nd))
        n_p = 1.2404e-2 * freqs[l] ** 2
        n_rays = ne_rays / (-n_p)
        n_rays += 1
        np.sqrt(n_rays, out=n_rays)
        n_rays *= -1
        n_rays += 1
        phi_ion = simps_avg(n_rays, rays[:, :, :, 3, :])
        phi_ion -= phi_ion[i0, ...]
        phi_ion *= (a_ / SPEED_OF_LIGHT)
        dg -= phi_ion
        g[:, :, :, l] += dg
    return g


def prior_penalty_mu(mu, mu_prior, xvec, yvec, zvec, rays, freqs, K=1e11, i0=0,
                     reference_axis_scramble=False):
    """inversion/iterative_newton.py:138-184."""
    Na, Nt, Nd, _, Ns = rays.shape
    freqs = np.asarray(freqs, dtype=np.float64)
    Nf = freqs.shape[0]
    shape = (len(xvec), len(yvec), len(zvec))
    ne = np.exp(mu).reshape(shape) * K
    r = np.zeros([Na, Nt, Nd, Nf], dtype=float)
    ne_rays = _interp_batched(xvec, yvec, zvec, ne, rays, reference_axis_scramble)
    dmu_rays = _interp_batched(xvec, yvec, zvec, (mu_prior - mu).reshape(shape), rays,
                               reference_axis_scramble)
    for l in range(Nf):
        n_p = 1.2404e-2 * freqs[l] ** 2
        a_ = 2 * np.pi * freqs[l]
        b_ = a_ / (2 * n_p * SPEED_OF_LIGHT)
        n_rays = ne_rays / (-n_p)
        n_rays += 1
        np.sqrt(n_rays, out=n_rays)
        integrand = ne_rays / n_rays
        integrand *= dmu_rays
        ion = simps_avg(integrand, rays[:, :, :, 3, :])
        ion -= ion[i0, ...]
        ion *= b_
        r[:, :, :, l] -= ion
    return r


# --------------------------------------------------------------------------
# misfit and weighted residual
# --------------------------------------------------------------------------
def misfit(g, dobs, CdCt):
    """``S = sum((g-dobs)^2/(CdCt+1e-15))/2``: inversion/line_search.py:48-49,
    tests/test_inversion.py:34."""
    return np.sum((g - dobs) ** 2 / (CdCt + 1e-15)) / 2.0


def weighted_residual(g, dobs, CdCt):
    """``dd = (g-dobs)/(CdCt+1e-15)``: inversion/gradient.py:33-37."""
    return (g - dobs) / (CdCt + 1e-15)


# --------------------------------------------------------------------------
# adjoint A10: exact transpose of the dTEC forward
# --------------------------------------------------------------------------
def adjoint_ray_coefficients(dd, i0):
    """c[ray] = dd[ray] - [ray is (i0,t,d)] * sum_i dd[i,t,d]  (transpose of the
    ``tec - tec[i0]`` step, inversion/forward_equation.py:50)."""
    c = np.array(dd, dtype=np.float64, copy=True)
    c[i0, :, :] -= dd.sum(axis=0)
    return c


def backproject(rays, xvec, yvec, zvec, coef):
    """acc[v] = sum_ray coef[ray] * sum_s w_s(ray) * phi_v(x_s): transpose of
    `tec` w.r.t. the grid values (trilinear hats, Simpson-avg weights)."""
    nx, ny, nz = len(xvec), len(yvec), len(zvec)
    ix, tx = find_indices(xvec, rays[..., 0, :])
    iy, ty = find_indices(yvec, rays[..., 1, :])
    iz, tz = find_indices(zvec, rays[..., 2, :])
    w = simps_weights_fast(rays[..., 3, :]) * np.asarray(coef)[..., None]
    acc = np.zeros(nx * ny * nz, dtype=np.float64)
    for cx in (0, 1):
        wx = tx if cx else 1 - tx
        for cy in (0, 1):
            wy = ty if cy else 1 - ty
            for cz in (0, 1):
                wz = tz if cz else 1 - tz
                flat = ((ix + cx) * ny + (iy + cy)) * nz + (iz + cz)
                acc += np.bincount(flat.ravel(), weights=(w * ((wx * wy) * wz)).ravel(),
                                   minlength=acc.size)
    return acc.reshape(nx, ny, nz)


def gradient_exact(rays, g, dobs, i0, K_ne, xvec, yvec, zvec, m, CdCt):
    """dS/dm for S = misfit(forward_equation(m)): the gradient the reference's
    finite-difference protocol (tests/test_inversion.py:71-87) checks.
    ``ne[v] * backproject(adjoint coefficients of dd)`` (chain rule through
    ``ne = K exp(m)/TECU``)."""
    dd = weighted_residual(g, dobs, CdCt)
    c = adjoint_ray_coefficients(dd, i0)
    return ne_from_m(m, K_ne) * backproject(rays, xvec, yvec, zvec, c)


# --------------------------------------------------------------------------
# adjoint A8: chord-length "ray dirac" gradient (geometry/ray_dirac.py,
# geometry/slab_method.py, inversion/gradient.py:15-20), sparse restatement
# --------------------------------------------------------------------------
def slab_method_ray_box(r0, n, inv_n, x_min, y_min, z_min, x_max, y_max, z_max):
    """geometry/slab_method.py:19-58 (returns only the chord length)."""
    def axis(lo, hi, o, inv):
        with np.errstate(invalid='ignore'):
            t1 = (lo - o) * inv
            t2 = (hi - o) * inv
        if np.isnan(t1):
            t1 = 0.
        if np.isnan(t2):
            t2 = 0.
        return min(t1, t2), max(t1, t2)
    tmin_x, tmax_x = axis(x_min, x_max, r0[0], inv_n[0])
    tmin_y, tmax_y = axis(y_min, y_max, r0[1], inv_n[1])
    tmin_z, tmax_z = axis(z_min, z_max, r0[2], inv_n[2])
    tmax = max(tmin_x, tmin_y, tmin_z)   # entry (reference names are swapped)
    tmin = min(tmax_x, tmax_y, tmax_z)   # exit
    if tmax < tmin and tmax > 0:
        return float(np.linalg.norm(n * (tmax - tmin)))
    return 0.0


def ray_dirac_sparse(ray, xvec, yvec, zvec):
    """One ray's ``dirac_ray`` of geometry/ray_dirac.py:16-33 as {(xi,yi,zi): ds}.

    ``ray`` is (4, Ns).  The line is first->last point (ray_dirac.py:21); for
    every sample the +-1 cell neighbourhood of its bisection cell is tested
    against the cell-centred box (ray_dirac.py:24-31); the value is ASSIGNED
    (idempotent), so a dict reproduces the dense array exactly.
    """
    x, y, z = xvec, yvec, zvec
    dx, dy, dz = x[1] - x[0], y[1] - y[0], z[1] - z[0]
    r0 = np.array(ray[0:3, 0], dtype=np.float64)
    n = np.array(ray[0:3, -1] - ray[0:3, 0], dtype=np.float64)
    n /= np.linalg.norm(n)
    with np.errstate(divide='ignore'):
        inv_n = 1. / n
    out = {}
    Ns = ray.shape[1]
    for s in range(Ns):
        xi_c = bisection(x, ray[0, s])
        yi_c = bisection(y, ray[1, s])
        zi_c = bisection(z, ray[2, s])
        for xi in range(max(0, xi_c - 1), min(len(x), xi_c + 2)):
            for yi in range(max(0, yi_c - 1), min(len(y), yi_c + 2)):
                for zi in range(max(0, zi_c - 1), min(len(z), zi_c + 2)):
                    if (xi, yi, zi) in out:
                        continue  # assignment of the same value
                    out[(xi, yi, zi)] = slab_method_ray_box(
                        r0, n, inv_n, x[xi] - dx / 2., y[yi] - dy / 2., z[zi] - dz / 2.,
                        x[xi] + dx / 2., y[yi] + dy / 2., z[zi] + dz / 2.)
    return out


def gradient_chord(rays, g, dobs, i0, K_ne, xvec, yvec, zvec, m, CdCt,
                   bug_compat=False):
    """Adjoint A8: ``G[v] = sum_ray l(ray,v) * ne[v] * dd[ray]``
    (inversion/gradient.py:15-20 einsum "ijklm,klm,ij->klm" over
    geometry/ray_dirac.py), dd as gradient.py:33-37.  ``bug_compat`` applies the
    reference's ``gradient -= gradient[i0,...]`` (gradient.py:55), which indexes
    grid-x rather than the antenna axis."""
    dd = weighted_residual(g, dobs, CdCt)
    ne = ne_from_m(m, K_ne)
    acc = np.zeros_like(ne)
    Na, Nt, Nd = rays.shape[:3]
    for i in range(Na):
        for j in range(Nt):
            for k in range(Nd):
                for (xi, yi, zi), ds in ray_dirac_sparse(rays[i, j, k], xvec, yvec, zvec).items():
                    acc[xi, yi, zi] += ds * dd[i, j, k]
    grad = acc * ne
    if bug_compat:
        grad = grad - grad[i0, ...]
    return grad


# --------------------------------------------------------------------------
# adjoint B (row A9): Cm.G^t.dd with a Gaussian model covariance,
# inversion/gradient_and_adjoint.py:12-167
# --------------------------------------------------------------------------
def gaussian_adjoint_ray(ray, ne_ray, xvec, yvec, zvec, sigma_m, L_m, Nkernel):
    """Contribution of ONE ray with unit residual, as a sparse dict
    ``{(xi,yi,zi): value}`` (gradient_and_adjoint.py:32-53 builds the per-voxel sample
    range, :67-93 integrates it).

    A voxel belongs to the box of sample ``idx`` when it lies in the slices
    ``max(0,c-Nk) : min(n-1, c+Nk+1)`` around the sample's ``bisection`` cell ``c`` on
    every axis (:42-44; the upper bound ``n-1`` is exclusive, so the last node of an axis
    never receives anything).  ``idx_min``/``idx_max`` are the first and last sample whose
    box holds the voxel; the integrand ``sigma_m^2 exp(-r^2/(2 L_m^2)) ne(s)`` is
    integrated with ``simps`` over ALL samples idx_min..idx_max (:80-92)."""
    x, y, z, s = ray
    Ns = x.shape[0]
    nx, ny, nz = len(xvec), len(yvec), len(zvec)
    idx_min = np.full((nx, ny, nz), Ns, dtype=np.int64)
    idx_max = np.full((nx, ny, nz), -1, dtype=np.int64)
    for idx in range(Ns):
        xi, yi, zi = bisection(xvec, x[idx]), bisection(yvec, y[idx]), bisection(zvec, z[idx])
        box = (slice(max(0, xi - Nkernel), min(nx - 1, xi + Nkernel + 1)),
               slice(max(0, yi - Nkernel), min(ny - 1, yi + Nkernel + 1)),
               slice(max(0, zi - Nkernel), min(nz - 1, zi + Nkernel + 1)))
        idx_max[box] = np.maximum(idx_max[box], idx)
        idx_min[box] = np.minimum(idx_min[box], idx)
    out = {}
    for xi, yi, zi in zip(*np.nonzero(idx_max >= 0)):
        seg = slice(idx_min[xi, yi, zi], idx_max[xi, yi, zi] + 1)
        Cm = (xvec[xi] - x[seg]) ** 2 + (yvec[yi] - y[seg]) ** 2 + (zvec[zi] - z[seg]) ** 2
        Cm = np.exp(Cm / (-2. * L_m ** 2)) * sigma_m ** 2 * ne_ray[seg]
        out[(int(xi), int(yi), int(zi))] = float(simps_avg(Cm, s[seg]))
    return out


def gaussian_adjoint(rays, dd, i0, K_ne, xvec, yvec, zvec, m, sigma_m, Nkernel, size_cell,
                     bug_compat=False):
    """``sum_d do_adjoint(rays[:,:,d], dd[:,:,d], ...)`` (gradient_and_adjoint.py:12-103,
    162-163).  ``ne`` along the ray is ``K_ne*exp(interp(m))/1e13`` (:37: the LOG model is
    interpolated, then exponentiated -- unlike the forward, which interpolates ne).
    ``bug_compat`` applies ``grad -= grad[i0,:,:]`` (:102: antenna index used on grid-x)."""
    L_m = Nkernel * size_cell
    Na, Nt, Nd = rays.shape[:3]
    grad = np.zeros((len(xvec), len(yvec), len(zvec)))
    for i in range(Na):
        for j in range(Nt):
            for k in range(Nd):
                ray = rays[i, j, k]
                ne_ray = K_ne * np.exp(rgi_linear(xvec, yvec, zvec, m, ray[0], ray[1], ray[2])) / 1e13
                for v, c in gaussian_adjoint_ray(ray, ne_ray, xvec, yvec, zvec, sigma_m, L_m,
                                                 Nkernel).items():
                    grad[v] += c * dd[i, j, k]
    if bug_compat:
        grad = grad - grad[i0, :, :]
    return grad


def compute_adjoint(rays, g, dobs, i0, K_ne, xvec, yvec, zvec, m, m_prior, CdCt, sigma_m,
                    Nkernel, size_cell, bug_compat=True):
    """``compute_adjoint`` (gradient_and_adjoint.py:137-167):
    ``Cm.G^t.Cd^-1.(g - dobs) + (m - m_prior)``."""
    dd = weighted_residual(g, dobs, CdCt)
    grad = gaussian_adjoint(rays, dd, i0, K_ne, xvec, yvec, zvec, m, sigma_m, Nkernel, size_cell,
                            bug_compat=bug_compat)
    return grad + m - m_prior


# --------------------------------------------------------------------------
# line search: inversion/line_search.py
# --------------------------------------------------------------------------
def vertex(x1, x2, x3, y1, y2, y3):
    """Vertex of the parabola through three points (inversion/line_search.py:13-43),
    restated in Lagrange form (algebraically identical; checked against the
    reference's expression in tests/test_oracle.py via the golden file)."""
    denom = (x1 - x2) * (x1 - x3) * (x2 - x3)
    A = (x3 * (y2 - y1) + x2 * (y1 - y3) + x1 * (y3 - y2)) / denom
    B = (x3 * x3 * (y1 - y2) + x2 * x2 * (y3 - y1) + x1 * x1 * (y2 - y3)) / denom
    C = (x2 * x3 * (x2 - x3) * y1 + x3 * x1 * (x3 - x1) * y2 + x1 * x2 * (x1 - x2) * y3) / denom
    xv = -B / (2 * A)
    return xv, C - B * B / (4 * A)


def line_search(rays, K_ne, xvec, yvec, zvec, m, i0, gradient, g, dobs, CdCt):
    """inversion/line_search.py:45-100 (without the plotting)."""
    S0 = misfit(g, dobs, CdCt)
    ep_a, S_a = [], []
    S = S0
    dd = (g - dobs) / (CdCt + 1e-15)
    ep = 1e-3
    g_ = forward_equation(rays, K_ne, xvec, yvec, zvec, m - ep * gradient, i0)
    Gm = (g - g_) / ep
    numerator = 2. * np.sum(dd * Gm)
    denominator = np.sum(Gm * Gm / (CdCt + 1e-15))
    epsilon_n = np.abs(numerator / denominator)
    it = 0
    while S >= S0 or it < 3:
        epsilon_n /= 2.
        g2 = forward_equation(rays, K_ne, xvec, yvec, zvec, m - epsilon_n * gradient, i0)
        S = misfit(g2, dobs, CdCt)
        ep_a.append(epsilon_n)
        S_a.append(S)
        if not np.isnan(S):
            it += 1
        if len(ep_a) > 200:
            break
    epsilon_n, S_p = vertex(*ep_a[-3:], *S_a[-3:])
    g3 = forward_equation(rays, K_ne, xvec, yvec, zvec, m - epsilon_n * gradient, i0)
    S = misfit(g3, dobs, CdCt)
    return epsilon_n, S, (S / S0 - 1.)


# --------------------------------------------------------------------------
# synthetic ionosphere (benchmark inputs): ionosphere/iri.py:20-68,
# ionosphere/simulation.py:45-112, inversion/initial_model.py:38-41,75-84
# --------------------------------------------------------------------------
def a_priori_model_(h, zenith, thin_f=False):
    """Four Chapman layers D/E/F1/F2 vs solar zenith angle: ionosphere/iri.py:20-68."""
    def peak_density(n0, dn, tau, b, zenith):
        y = zenith / tau
        return n0 + dn * np.exp(-y ** 2) / (1. + y ** (2 * b))

    def peak_height(z0, dz, rho, chi0, zenith):
        return z0 + dz / (1. + np.exp(-(zenith - chi0) / rho))

    def layer_density(nm, zm, H, z):
        y = (z - zm) / H
        return nm * np.exp(1. / 2. * (1. - y - np.exp(-y)))
    y = zenith / 58.
    nm_d = 4e8 + 5.9e8 * np.exp(-y ** 2) if y < 1 else 4e8
    n_d = layer_density(nm_d, peak_height(81., 7., 7.46, 100., zenith), 8., h)
    n_e = layer_density(peak_density(1.6e9, 1.6e11, 87., 8.7, zenith), 110., 11., h)
    H_f1 = 20. if thin_f else 40.
    n_f1 = layer_density(peak_density(2.0e11, 9.1e10, 54., 13.6, zenith), 185., H_f1, h)
    H_f2 = 27.5 if thin_f else 55.
    n_f2 = layer_density(peak_density(7.7e10, 4.4e11, 111., 4.8, zenith),
                         peak_height(242., 75., 7.46, 96., zenith), H_f2, h)
    return np.atleast_1d(n_d + n_e + n_f1 + n_f2)


def turbulent_realization(xvec, yvec, zvec, sigma, corr, seed):
    """Matern-5/2 Gaussian random field: IonosphereSimulation.__init__ and
    .realization, ionosphere/simulation.py:45-112 (same spectrum, same FFT
    de-shift by sign flips, same rescale to sigma)."""
    from math import gamma
    nx, ny, nz = np.size(xvec), np.size(yvec), np.size(zvec)
    dx, dy, dz = xvec[1] - xvec[0], yvec[1] - yvec[0], zvec[1] - zvec[0]
    sx, sy, sz = 1. / (dx * nx), 1. / (dy * ny), 1. / (dz * nz)
    lvec = np.linspace(0, sx * nx / 2., nx)
    mvec = np.linspace(0, sy * ny / 2., ny)
    nvec = np.linspace(0, sz * nz / 2., nz)
    L, Mm, Nn = np.meshgrid(lvec, mvec, nvec, indexing='ij')
    s2 = L ** 2
    s2 += Mm ** 2
    s2 += Nn ** 2
    s2 = np.fft.ifftshift(s2)
    n = 3.
    nu = 5 / 2.
    S = sigma ** 2 * 2 ** n * np.pi ** (n / 2.) * gamma(nu + n / 2.) * (2 * nu) ** nu \
        / gamma(nu) / corr ** (2 * nu) * (2 * nu / corr ** 2 + 4 * np.pi ** 2 * s2) ** (-nu - n / 2.)
    S = np.sqrt(S)
    if seed is not None:
        np.random.seed(seed)
    Z = np.random.normal(size=S.shape) + 1j * np.random.normal(size=S.shape)
    Y = S * Z
    B = (np.fft.ifftn(Y, (nx, ny, nz), axes=(0, 1, 2))).real * (sx * nx) * (sy * ny) * (sz * nz)
    B[::2, :, :] *= -1
    B[:, ::2, :] *= -1
    B[:, :, ::2] *= -1
    B *= sigma / np.std(B)
    return B


# --------------------------------------------------------------------------
# true tricubic interpolation and bent rays (notebook specs only):
# notebooks/TricubicInterpolation.ipynb[cell 0]:138-299,1192-1257, notebooks/DeriveTricubic.ipynb[cell 0]:87-141,
# notebooks/FermatClass.ipynb[cell 0]:60-96.  No reference numbers exist for these; tests validate this
# restatement against exact tricubic polynomials and against scipy.integrate.odeint.
# --------------------------------------------------------------------------
def _diff_axis(f, g, axis):
    """d f / d axis at the nodes: 4th-order central (8(f[i+1]-f[i-1]) - (f[i+2]-f[i-2]))/12 over the local spacing
    (g[i+1]-g[i-1])/2 (DeriveTricubic.ipynb[cell 0]:87-107,124-141); 2nd-order central one node from a face and
    one-sided on the face (extension: the notebook asserts 2 <= i <= n-3)."""
    f = np.moveaxis(np.asarray(f, dtype=np.float64), axis, 0)
    g = np.asarray(g, dtype=np.float64)
    n = g.size
    d = np.empty_like(f)
    sh = (slice(None),) + (None,) * (f.ndim - 1)
    if n >= 5:
        h = 0.5 * (g[3:n - 1] - g[1:n - 3])
        d[2:n - 2] = (8.0 * (f[3:n - 1] - f[1:n - 3]) - (f[4:n] - f[0:n - 4])) / 12.0 / h[sh]
    for i in ([1, n - 2] if n >= 3 else []):
        if 1 <= i <= n - 2 and not (2 <= i <= n - 3):
            d[i] = (f[i + 1] - f[i - 1]) / (g[i + 1] - g[i - 1])
    d[0] = (f[1] - f[0]) / (g[1] - g[0])
    d[n - 1] = (f[n - 1] - f[n - 2]) / (g[n - 1] - g[n - 2])
    return np.moveaxis(d, 0, axis)


def tricubic_derivs(xvec, yvec, zvec, f):
    """The 8 grids (f, fx, fy, fz, fxy, fxz, fyz, fxyz) the tricubic interpolant is built from."""
    f = np.asarray(f, dtype=np.float64)
    fx, fy, fz = _diff_axis(f, xvec, 0), _diff_axis(f, yvec, 1), _diff_axis(f, zvec, 2)
    fxy, fxz, fyz = _diff_axis(fx, yvec, 1), _diff_axis(fx, zvec, 2), _diff_axis(fy, zvec, 2)
    return np.stack([f, fx, fy, fz, fxy, fxz, fyz, _diff_axis(fxy, zvec, 2)])


def _hermite(t):
    t2, t3 = t * t, t * t * t
    h = np.stack([2 * t3 - 3 * t2 + 1, -2 * t3 + 3 * t2, t3 - 2 * t2 + t, t3 - t2])
    dh = np.stack([6 * t2 - 6 * t, -6 * t2 + 6 * t, 3 * t2 - 4 * t + 1, 3 * t2 - 2 * t])
    return h, dh


def tricubic_interp(xvec, yvec, zvec, derivs, x, y, z, grad=False):
    """C1 tricubic (Lekien-Marsden) interpolant: the tensor product of cubic Hermite bases matching
    (f, fx, fy, fz, fxy, fxz, fyz, fxyz) at the 8 corners of the cell -- the polynomial the notebook's 64x64 matrix
    yields (TricubicInterpolation.ipynb[cell 0]:286-299,1192-1257), with derivatives scaled by the cell size."""
    x, y, z = (np.asarray(v, dtype=np.float64) for v in (x, y, z))
    shape = x.shape
    x, y, z = x.ravel(), y.ravel(), z.ravel()
    idx, ts, hs = [], [], []
    for g, p in ((xvec, x), (yvec, y), (zvec, z)):
        g = np.asarray(g, dtype=np.float64)
        i, t = find_indices(g, p)
        idx.append(i)
        hs.append(g[i + 1] - g[i])
        ts.append(t)
    (bu, du), (bv, dv), (bw, dw) = _hermite(ts[0]), _hermite(ts[1]), _hermite(ts[2])
    which = {(0, 0, 0): 0, (1, 0, 0): 1, (0, 1, 0): 2, (0, 0, 1): 3, (1, 1, 0): 4, (1, 0, 1): 5, (0, 1, 1): 6, (1, 1, 1): 7}
    f = np.zeros_like(x)
    fx, fy, fz = np.zeros_like(x), np.zeros_like(x), np.zeros_like(x)
    for i in (0, 1):
        for j in (0, 1):
            for k in (0, 1):
                for a in (0, 1):
                    for b in (0, 1):
                        for c in (0, 1):
                            val = derivs[which[(a, b, c)]][idx[0] + i, idx[1] + j, idx[2] + k]
                            val = val * (hs[0] if a else 1.) * (hs[1] if b else 1.) * (hs[2] if c else 1.)
                            Bx, By, Bz = bu[2 * a + i], bv[2 * b + j], bw[2 * c + k]
                            f += val * Bx * By * Bz
                            fx += val * du[2 * a + i] * By * Bz
                            fy += val * Bx * dv[2 * b + j] * Bz
                            fz += val * Bx * By * dw[2 * c + k]
    if grad:
        return f.reshape(shape), np.stack([fx / hs[0], fy / hs[1], fz / hs[2]], -1).reshape(shape + (3,))
    return f.reshape(shape)


def bent_ray_rhs(xvec, yvec, zvec, derivs_n, state, z):
    """FermatClass.ipynb[cell 0]:60-96, type 'z': d[px,py,pz,x,y,s]/dz."""
    px, py, pz, x, y, s = state
    n, g = tricubic_interp(xvec, yvec, zvec, derivs_n, np.array([x]), np.array([y]), np.array([z]), grad=True)
    n, g = float(n[0]), g[0]
    return np.array([g[0] * n / pz, g[1] * n / pz, g[2] * n / pz, px / pz, py / pz, n / pz])


def bent_ray_rk4(xvec, yvec, zvec, derivs_n, origin, direction, tmax, N, substeps=4):
    """Classical RK4 in z with ``substeps`` steps per sample interval; returns x, y, z, s at
    ``z = linspace(z0, tmax, N)`` (the layout of Fermat.integrate_ray)."""
    o = np.asarray(origin, dtype=np.float64)
    d = np.asarray(direction, dtype=np.float64)
    p = d / np.sqrt(d[0] ** 2 + d[1] ** 2 + d[2] ** 2)
    q = np.array([p[0], p[1], p[2], o[0], o[1], 0.0])
    zs = np.linspace(o[2], tmax, N)
    dzs = (tmax - o[2]) / (N - 1) if N > 1 else 0.0
    out = np.zeros((4, N))
    for i in range(N):
        zi = tmax if (i == N - 1 and N > 1) else o[2] + i * dzs
        out[:, i] = q[3], q[4], zi, q[5]
        if i == N - 1:
            break
        h = dzs / substeps
        for k in range(substeps):
            zz = zi + k * h
            k1 = bent_ray_rhs(xvec, yvec, zvec, derivs_n, q, zz)
            k2 = bent_ray_rhs(xvec, yvec, zvec, derivs_n, q + 0.5 * h * k1, zz + 0.5 * h)
            k3 = bent_ray_rhs(xvec, yvec, zvec, derivs_n, q + 0.5 * h * k2, zz + 0.5 * h)
            k4 = bent_ray_rhs(xvec, yvec, zvec, derivs_n, q + h * k3, zz + h)
            q = q + h / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)
    assert np.allclose(out[2], zs)
    return out
