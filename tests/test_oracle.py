"""The oracle against (a) golden vectors produced by the reference's own modules
(tests/golden/make_golden.py), (b) the installed SciPy where it still ships the
algorithm.  CPU only."""
import numpy as np
import pytest
import scipy.integrate
from scipy.interpolate import RegularGridInterpolator

from oracle import ionotomo_oracle as O

RTOL = 1e-12  # oracle vs reference-generated vectors (same algorithm, fp64)


def test_interp_matches_reference_tricubic(golden):
    g = golden("tricubic")
    x, y, z = np.meshgrid(g["xvec"], g["yvec"], g["zvec"], indexing="ij")
    M = x * y * z + x - y - 2 * z + x ** 2          # tests/test_tricubic.py:11
    res = O.rgi_linear(g["xvec"], g["yvec"], g["zvec"], M, g["pts"], g["pts"], g["pts"])
    np.testing.assert_allclose(res, g["res_batch"], rtol=RTOL, atol=1e-15)
    # batch == per-point, the reference's own protocol (tests/test_tricubic.py:21-27)
    np.testing.assert_array_equal(g["res_batch"][:50], g["res_scalar"])
    one = np.array([O.rgi_linear(g["xvec"], g["yvec"], g["zvec"], M, p, p, p) for p in g["pts"][:50]])
    np.testing.assert_array_equal(one, res[:50])
    res = O.rgi_linear(g["xvec"], g["yvec"], g["zvec"], M, g["px"], g["py"], g["pz"])
    np.testing.assert_allclose(res, g["res_rand"], rtol=RTOL, atol=1e-15)
    res = O.rgi_linear(g["xvec"], g["yvec"], g["zvec"], M, g["ex"], g["ey"], g["ez"],
                       bounds_error=False)
    np.testing.assert_allclose(res, g["res_extrap"], rtol=RTOL, atol=1e-15)
    assert bool(g["oob_raises"])
    with pytest.raises(ValueError):
        O.rgi_linear(g["xvec"], g["yvec"], g["zvec"], M, np.array([0.5, 1.2]),
                     np.array([0.5, 0.5]), np.array([0.5, 0.5]))
    with pytest.raises(ValueError):
        O.rgi_linear(g["xvec"], g["yvec"], g["zvec"], M, np.array([np.nan]),
                     np.array([0.5]), np.array([0.5]))


def test_bisection_matches_reference(golden):
    g = golden("tricubic")
    idx = np.array([O.bisection(g["xvec"], v) for v in g["bvals"]])
    np.testing.assert_array_equal(idx, g["bidx"])


def test_interp_matches_installed_scipy():
    rng = np.random.RandomState(0)
    xv = np.sort(rng.uniform(0, 1, 13)); yv = np.linspace(-1, 2, 9); zv = np.linspace(0, 5, 17)
    M = rng.normal(size=(13, 9, 17))
    p = np.stack([rng.uniform(xv[0], xv[-1], 300), rng.uniform(-1, 2, 300), rng.uniform(0, 5, 300)], -1)
    rgi = RegularGridInterpolator((xv, yv, zv), M, bounds_error=True)
    np.testing.assert_allclose(O.rgi_linear(xv, yv, zv, M, p[:, 0], p[:, 1], p[:, 2]), rgi(p),
                               rtol=1e-13, atol=1e-14)
    rge = RegularGridInterpolator((xv, yv, zv), M, bounds_error=False, fill_value=None)
    q = p * 1.7 - 0.4
    np.testing.assert_allclose(O.rgi_linear(xv, yv, zv, M, q[:, 0], q[:, 1], q[:, 2], bounds_error=False),
                               rge(q), rtol=1e-12, atol=1e-13)


def test_simps_odd_matches_installed_scipy():
    rng = np.random.RandomState(1)
    x = np.sort(rng.uniform(size=(6, 5, 31)), axis=-1)
    y = rng.normal(size=x.shape)
    np.testing.assert_allclose(O.simps_avg(y, x), scipy.integrate.simpson(y, x=x, axis=-1),
                               rtol=1e-12, atol=1e-14)


def test_simps_even_is_avg_rule():
    """even='avg' composed from installed-SciPy pieces (odd sub-ranges + trapezoids)."""
    rng = np.random.RandomState(2)
    for N in (2 + 2, 10, 30, 128):
        x = np.sort(rng.uniform(size=(7, N)), axis=-1)
        y = rng.normal(size=x.shape)
        first = scipy.integrate.simpson(y[:, :-1], x=x[:, :-1], axis=-1) \
            + 0.5 * (x[:, -1] - x[:, -2]) * (y[:, -1] + y[:, -2])
        last = scipy.integrate.simpson(y[:, 1:], x=x[:, 1:], axis=-1) \
            + 0.5 * (x[:, 1] - x[:, 0]) * (y[:, 1] + y[:, 0])
        np.testing.assert_allclose(O.simps_avg(y, x), 0.5 * (first + last), rtol=1e-12, atol=1e-14)
    # and it is NOT modern simpson for even N: the distinction matters
    x = np.linspace(0, 1, 30); y = np.exp(3 * x)
    assert abs(O.simps_avg(y, x) - scipy.integrate.simpson(y, x=x)) > 1e-8


def test_simps_even_known_answers_from_old_scipy_docstring():
    """Published known answers for an EVEN number of samples: the docstring example of
    ``scipy.integrate.simps`` (SciPy <= 1.10) -- ``x = np.arange(0, 10); y = x**3`` gives
    ``simps(y, x) == 1642.5`` with the default ``even='avg'`` and ``1644.5`` with ``even='first'``
    (exact integral 1640.25); ``simps(x, x) == 40.5``.  Modern ``simpson`` returns 1640.5 instead."""
    x = np.arange(0, 10.)
    y = np.power(x, 3)
    assert O.simps_avg(x, x) == 40.5
    assert O.simps_avg(y, x) == 1642.5
    first = O._basic_simps(y, 0, 10 - 3, x) + 0.5 * (x[-1] - x[-2]) * (y[-1] + y[-2])
    assert first == 1644.5
    assert scipy.integrate.simpson(y, x=x) == 1640.5          # the rule the reference did NOT run
    w = O.simps_weights_fast(x)
    np.testing.assert_allclose((w * y).sum(), 1642.5, rtol=1e-14)


def test_simps_weights():
    rng = np.random.RandomState(3)
    for N in (3, 4, 5, 9, 10, 31, 64):
        x = np.sort(rng.uniform(size=(3, N)), axis=-1)
        y = rng.normal(size=x.shape)
        w = O.simps_weights(x)
        np.testing.assert_allclose((w * y).sum(-1), O.simps_avg(y, x), rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(O.simps_weights_fast(x), w, rtol=1e-12, atol=1e-15)


@pytest.mark.parametrize("tag", ["odd", "even"])
def test_rays_match_reference_odeint(golden, tag):
    g = golden("forward_" + tag)
    Ns = int(g["Ns"])
    rays = O.cast_ray(g["origins"], g["directions"], float(g["tmax"]), Ns)
    # LSODA (rtol=atol~1.5e-8) vs closed form: SURVEY Appendix A.3 measured 7e-13 km
    np.testing.assert_allclose(rays, g["rays"], rtol=0, atol=1e-9)
    one = np.stack(O.integrate_ray_straight(g["single_origin"], g["single_direction"], float(g["tmax"]), Ns))
    np.testing.assert_allclose(one, g["single"], rtol=0, atol=1e-9)
    np.testing.assert_array_equal(one, rays[1, 0, 2])


@pytest.mark.parametrize("tag", ["odd", "even"])
def test_forward_matches_reference(golden, tag):
    g = golden("forward_" + tag)
    rays = g["rays"]
    dtec = O.forward_equation(rays, float(g["K_ne"]), g["xvec"], g["yvec"], g["zvec"], g["m"], int(g["i0"]))
    scale = np.abs(g["dtec"]).max()
    np.testing.assert_allclose(dtec, g["dtec"], rtol=0, atol=1e-11 * scale)
    t0 = O.tec(rays[0], g["xvec"], g["yvec"], g["zvec"], O.ne_from_m(g["m"], float(g["K_ne"])))
    np.testing.assert_allclose(t0, g["tec_a0"], rtol=RTOL)


@pytest.mark.parametrize("tag", ["odd", "even"])
def test_phase_forward_matches_reference(golden, tag):
    g = golden("forward_" + tag)
    ph = O.phase_forward_equation(g["mu"], g["clock"], g["const"], g["xvec"], g["yvec"], g["zvec"],
                                  g["rays"], g["freqs"], K=1e11, i0=int(g["i0"]),
                                  reference_axis_scramble=True)
    np.testing.assert_allclose(ph, g["phase"], rtol=1e-10, atol=1e-10 * np.abs(g["phase"]).max())
    pen = O.prior_penalty_mu(g["mu"], g["mu_prior"], g["xvec"], g["yvec"], g["zvec"], g["rays"],
                             g["freqs"], K=1e11, i0=int(g["i0"]), reference_axis_scramble=True)
    np.testing.assert_allclose(pen, g["penalty"], rtol=1e-10, atol=1e-10 * np.abs(g["penalty"]).max())


def test_chord_adjoint_matches_reference(golden):
    g = golden("chord")
    rays = g["rays"]
    Na, _, Nd = rays.shape[:3]
    dense = np.zeros_like(g["dirac"])
    for i in range(Na):
        for k in range(Nd):
            for (xi, yi, zi), ds in O.ray_dirac_sparse(rays[i, 0, k], g["xvec"], g["yvec"], g["zvec"]).items():
                dense[i, k, xi, yi, zi] = ds
    np.testing.assert_allclose(dense, g["dirac"], rtol=1e-12, atol=1e-12)
    # do_gradient(rays, dd, ne_tci,...) = einsum(dirac, ne, dd): with CdCt+1e-15 == 1 and
    # g-dobs == dd, K_ne e^m / TECU == ne
    dd = g["dd"][:, None, :]
    grad = O.gradient_chord(rays, dd, np.zeros_like(dd), 0, 1e13, g["xvec"], g["yvec"], g["zvec"],
                            np.log(g["ne"]), np.ones_like(dd) - 1e-15)
    np.testing.assert_allclose(grad, g["G"], rtol=1e-10, atol=1e-10 * np.abs(g["G"]).max())


def test_line_search_matches_reference(golden):
    g = golden("line_search")
    np.testing.assert_allclose(O.vertex(*g["vx"][:3], *g["vy"][:3]), g["v1"], rtol=1e-9)
    np.testing.assert_allclose(O.vertex(*g["vx"][3:], *g["vy"][3:]), g["v2"], rtol=1e-7)
    eps, S, red = O.line_search(g["rays"], float(g["K_ne"]), g["xvec"], g["yvec"], g["zvec"], g["m0"],
                                0, g["grad"], g["g"], g["dobs"], g["CdCt"])
    np.testing.assert_allclose(eps, float(g["eps"]), rtol=1e-6)
    np.testing.assert_allclose(S, float(g["S"]), rtol=1e-6)
    np.testing.assert_allclose(O.misfit(g["g"], g["dobs"], g["CdCt"]), float(g["S0"]), rtol=1e-12)


def test_synthetic_matches_reference(golden):
    g = golden("synthetic")
    dm = O.turbulent_realization(g["xvec"], g["yvec"], g["zvec"], np.log(2.), 20., 1234)
    np.testing.assert_allclose(dm, g["dm"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(O.a_priori_model_(g["h"], 45.), g["chap45"], rtol=1e-13)
    np.testing.assert_allclose(O.a_priori_model_(g["h"], 80., thin_f=True), g["chap80"], rtol=1e-13)


def test_exact_adjoint_dot_product_and_fd(golden):
    """A10 is the transpose of the dTEC forward: <G x, y> == <x, G^T y>, and it is the
    gradient the reference's FD protocol checks (tests/test_inversion.py:71-87)."""
    g = golden("forward_even")
    xv, yv, zv, rays = g["xvec"], g["yvec"], g["zvec"], g["rays"]
    i0 = int(g["i0"])
    rng = np.random.RandomState(5)
    xfield = rng.normal(size=g["ne"].shape)
    yv_ = rng.normal(size=rays.shape[:3])
    t = O.tec(rays, xv, yv, zv, xfield)
    Gx = t - t[i0]
    GTy = O.backproject(rays, xv, yv, zv, O.adjoint_ray_coefficients(yv_, i0))
    assert abs((Gx * yv_).sum() - (xfield * GTy).sum()) <= 1e-12 * abs((Gx * yv_).sum())
    K_ne, m = float(g["K_ne"]), g["m"]
    dobs = g["dtec"] + 0.01 * rng.normal(size=g["dtec"].shape)
    CdCt = np.full(dobs.shape, 0.01 ** 2)
    gm = O.forward_equation(rays, K_ne, xv, yv, zv, m, i0)
    grad = O.gradient_exact(rays, gm, dobs, i0, K_ne, xv, yv, zv, m, CdCt)
    S0 = O.misfit(gm, dobs, CdCt)
    idx = np.argsort(-np.abs(grad).ravel())[:5]
    for f in idx:
        v = np.unravel_index(f, m.shape)
        mp = m.copy(); mp[v] += 1e-6
        mm = m.copy(); mm[v] -= 1e-6
        fd = (O.misfit(O.forward_equation(rays, K_ne, xv, yv, zv, mp, i0), dobs, CdCt)
              - O.misfit(O.forward_equation(rays, K_ne, xv, yv, zv, mm, i0), dobs, CdCt)) / 2e-6
        assert abs(fd - grad[v]) <= 1e-5 * abs(grad[v]) + 1e-9 * abs(S0)


def test_optical_path_matches_reference_odeint(golden):
    """straight_line_approx=False as shipped: same geometry, s = int n dz / pz (LSODA in the reference)."""
    g = golden("optical_path")
    n_field = O.ne2n(g["ne"], float(g["frequency"]))
    np.testing.assert_allclose(n_field, g["n_field"], rtol=1e-15)
    straight = O.cast_ray(g["origins"], g["directions"], float(g["tmax"]), int(g["Ns"]))
    np.testing.assert_allclose(straight[..., :3, :], g["rays"][..., :3, :], rtol=0, atol=1e-9)
    for idx in np.ndindex(*g["rays"].shape[:3]):
        s = O.optical_path(straight[idx], g["xvec"], g["yvec"], g["zvec"], n_field)
        # LSODA runs at rtol = atol ~ 1.5e-8 on a path of ~800 km
        np.testing.assert_allclose(s, g["rays"][idx][3], rtol=0, atol=2e-4)
        assert np.all(s[1:] < straight[idx][3, 1:])      # n < 1: the optical path is shorter


def test_config1_matches_reference(golden):
    """BASELINE.json configs[0] (10 antennas x 20 directions x 1 time, 50x50x30 grid, Ns = 30), produced
    end to end by the reference's cast_ray (odeint per ray) and forward_equation."""
    g = golden("config1")
    o = np.broadcast_to(g["origins"][:, None, None, :], (10, 1, 20, 3))
    d = np.broadcast_to(g["directions"][None], (10, 1, 20, 3))
    rays = O.cast_ray(o, d, 1000., 30)
    np.testing.assert_allclose(rays, g["rays"], rtol=0, atol=1e-9)
    dtec = O.forward_equation(rays, float(g["K_ne"]), g["xvec"], g["yvec"], g["zvec"], g["m"], 0)
    np.testing.assert_allclose(dtec, g["dtec"], rtol=0, atol=1e-12 * np.abs(g["dtec"]).max() + 1e-13)


def test_synthetic_generators_match_reference(golden):
    """The benchmark's input recipes (torch versions in the package) against the reference's generators."""
    from ionotomo_b200.ionosphere import synthetic as S
    g = golden("synthetic")
    dm = S.matern52_field(g["xvec"], g["yvec"], g["zvec"], np.log(2.), 20., 1234).numpy()
    np.testing.assert_allclose(dm, g["dm"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(S.chapman_profile(g["h"], 45.), g["chap45"], rtol=1e-13)
    np.testing.assert_allclose(S.chapman_profile(g["h"], 80., thin_f=True), g["chap80"], rtol=1e-13)
    ants = S.lofar_stations_enu_km()
    assert ants.shape == (62, 3) and abs(ants.mean(0)).max() < 1e-6      # ENU about the centroid
    assert -48 < ants[:, 0].min() < -47 and 9 < ants[:, 0].max() < 10     # extents quoted in SURVEY.md (d)
    d = S.directions_in_fov(200, 4., 1234)
    assert np.allclose(np.linalg.norm(d, axis=1), 1.) and np.degrees(np.arccos(d[:, 2])).max() <= 2.0 + 1e-9
    dt = S.track_directions(d, 100)
    assert np.array_equal(dt[50], d) and np.allclose(np.linalg.norm(dt, axis=-1), 1.)
    # 8 s of Earth rotation = 2 arcmin on the sky at most
    assert np.degrees(np.arccos(np.clip((dt[51] * dt[50]).sum(-1), -1, 1))).max() <= 8 * 15 / 3600. + 1e-9
    w = S.make_workload(Na=10, Nt=3, Nd=20, nx=50, ny=50, nz=30, device="cpu")
    assert tuple(w["origins"].shape) == (10, 3, 20, 3) and w["Ns"] == 30
    # every ray of the case stays inside the tight box, with the 20-cell padding
    ends_x = w["origins"][..., 0] + w["directions"][..., 0] / w["directions"][..., 2] * (1000. - w["origins"][..., 2])
    assert float(ends_x.min()) - w["xvec"][0] >= 19.9 * w["dx_km"] and w["xvec"][-1] - float(ends_x.max()) >= 19.9 * w["dx_km"]


def test_frames_helpers():
    from ionotomo_b200.geometry import frames
    lon, dec = np.radians(6.87), np.radians(54.)
    ha = np.radians(np.array([-20., 0., 35.]))
    R = frames.pointing_rotation(lon, ha, dec)
    for j in range(3):
        np.testing.assert_array_equal(R[j], O.pointing_rotation(lon, ha[j], dec))
        np.testing.assert_allclose(R[j] @ R[j].T, np.eye(3), atol=1e-15)
    # GMST advances by one sidereal day per 0.99727 solar days
    g0, g1 = frames.gmst_rad(2457700.5), frames.gmst_rad(2457700.5 + 0.9972695663)
    assert abs(((g1 - g0 + np.pi) % (2 * np.pi)) - np.pi) < 1e-6
    d = frames.icrs_to_itrs_simple(np.array([1.0, 2.0]), np.array([0.3, -0.2]), np.array([2457700.5, 2457700.6]))
    assert d.shape == (2, 2, 3) and np.allclose(np.linalg.norm(d, axis=-1), 1.)


def test_gaussian_adjoint_matches_reference(golden):
    """Adjoint B against the reference's own do_adjoint / compute_adjoint (gradient_and_adjoint.py:12-167)."""
    g = golden("adjoint_gauss")
    xv, yv, zv, i0 = g["xvec"], g["yvec"], g["zvec"], int(g["i0"])
    K, sig, Nk, cell = float(g["K_ne"]), float(g["sigma_m"]), int(g["Nkernel"]), float(g["size_cell"])
    dd = O.weighted_residual(g["g"], g["dobs"], g["CdCt"])
    s0 = O.gaussian_adjoint(g["rays"][:, :, :1], dd[:, :, :1], i0, K, xv, yv, zv, g["m"], sig, Nk, cell,
                            bug_compat=True)
    np.testing.assert_allclose(s0, g["slice0"], rtol=RTOL, atol=RTOL * np.abs(g["slice0"]).max())
    adj = O.compute_adjoint(g["rays"], g["g"], g["dobs"], i0, K, xv, yv, zv, g["m"], g["m_prior"], g["CdCt"],
                            sig, Nk, cell)
    np.testing.assert_allclose(adj, g["adj"], rtol=RTOL, atol=RTOL * np.abs(g["adj"]).max())
    wide = O.compute_adjoint(g["rays"][:2, :1], g["g"][:2, :1], g["dobs"][:2, :1], 0, K, xv, yv, zv, g["m"],
                             g["m_prior"], g["CdCt"][:2, :1], 1.3, 5, 7.)
    np.testing.assert_allclose(wide, g["adj_wide"], rtol=RTOL, atol=RTOL * np.abs(g["adj_wide"]).max())
    # the exclusive upper slice bound: the last node of every axis receives nothing
    raw = O.gaussian_adjoint(g["rays"], dd, i0, K, xv, yv, zv, g["m"], sig, Nk, cell)
    assert np.all(raw[-1] == 0) and np.all(raw[:, -1] == 0) and np.all(raw[:, :, -1] == 0)
    assert np.abs(raw).max() > 0


def test_simps_nonuniform_even_n_against_old_scipy(golden):
    """tests/golden/simps_even.npz comes from the old SciPy routine itself (tests/golden/old_scipy_simps.py, a
    transcription of scipy/integrate/quadrature.py of SciPy 0.19-1.5) on NON-uniform abscissae: the oracle's
    even='avg' rule and its per-sample weights must reproduce it for even and odd N."""
    g = golden("simps_even")
    assert float(g["doc_avg"]) == 1642.5 and float(g["doc_first"]) == 1644.5      # the old docstring's answers
    for N in (2, 3, 4, 5, 6, 9, 10, 30, 31, 64, 128, 129, 256):
        x, y, ref = g["x%d" % N], g["y%d" % N], g["avg%d" % N]
        np.testing.assert_allclose(O.simps_avg(y, x), ref, rtol=1e-13, atol=1e-13 * np.abs(y).max() * np.ptp(x))
        w = np.stack([O.simps_weights(xi) for xi in x])
        np.testing.assert_allclose((w * y).sum(1), ref, rtol=1e-12, atol=1e-12 * np.abs(y).max() * np.ptp(x))
        if N % 2 == 0 and N > 2:
            # 'avg' is the mean of the two one-sided rules
            np.testing.assert_allclose(0.5 * (g["first%d" % N] + g["last%d" % N]), ref, rtol=1e-13)


def test_fermat_arclength_golden(golden):
    """Fermat(type='s') of the reference (odeint on the straight ODE) against the closed form."""
    g = golden("fermat_s")
    o, d = g["origins"], g["directions"]
    for idx in np.ndindex(*o.shape[:3]):
        x, y, z, s = O.integrate_ray_arclength(o[idx], d[idx], float(g["tmax"]), int(g["Ns"]))
        np.testing.assert_allclose(np.stack([x, y, z, s]), g["rays"][idx], rtol=0, atol=1e-9)


def test_tricubic_oracle_reproduces_tricubic_polynomials_and_is_c1():
    """The 4th-order differences are exact for cubics, so the Hermite data are exact and the interpolant must
    reproduce any polynomial of degree <= 3 per variable (uniform axes, away from the faces); across a cell face
    value and gradient are continuous."""
    xv, yv, zv = np.linspace(-3, 4, 12), np.linspace(0, 5, 11), np.linspace(1, 9, 13)
    X, Y, Z = np.meshgrid(xv, yv, zv, indexing="ij")

    def poly(x, y, z):
        return (1 + 0.3 * x - 0.2 * x ** 2 + 0.05 * x ** 3) * (2 - 0.1 * y + 0.03 * y ** 3) * \
            (0.5 + 0.2 * z - 0.01 * z ** 2 + 0.002 * z ** 3)
    D = O.tricubic_derivs(xv, yv, zv, poly(X, Y, Z))
    rng = np.random.RandomState(0)
    px, py, pz = rng.uniform(xv[2], xv[-3], 300), rng.uniform(yv[2], yv[-3], 300), rng.uniform(zv[2], zv[-3], 300)
    f, g = O.tricubic_interp(xv, yv, zv, D, px, py, pz, grad=True)
    np.testing.assert_allclose(f, poly(px, py, pz), rtol=1e-13)
    eps = 1e-6
    np.testing.assert_allclose(g[:, 1], (poly(px, py + eps, pz) - poly(px, py - eps, pz)) / (2 * eps), rtol=1e-7, atol=1e-7)
    # C1 across a face, for a non-polynomial field
    D2 = O.tricubic_derivs(xv, yv, zv, np.sin(X) * np.cos(0.7 * Y) + 0.1 * Z ** 2)
    x0 = xv[5]
    fl, gl = O.tricubic_interp(xv, yv, zv, D2, np.array([x0 - 1e-9]), np.array([2.2]), np.array([4.4]), grad=True)
    fr, gr = O.tricubic_interp(xv, yv, zv, D2, np.array([x0 + 1e-9]), np.array([2.2]), np.array([4.4]), grad=True)
    assert abs(fl - fr) < 1e-8 and np.abs(gl - gr).max() < 1e-7


def test_bent_ray_oracle_against_odeint():
    """RK4 restatement of the notebook's ray equations (FermatClass.ipynb[cell 0]:60-96) against SciPy's odeint (what
    the notebook calls), and sanity: n = 1 gives the straight ray of Fermat.integrate_ray."""
    from scipy.integrate import odeint
    xv, yv, zv = np.linspace(-3, 4, 12), np.linspace(0, 5, 11), np.linspace(1, 9, 13)
    X, Y, Z = np.meshgrid(xv, yv, zv, indexing="ij")
    ne = 1e12 * np.exp(-((Z - 5) / 2.) ** 2) * (1 + 0.3 * np.sin(X))
    Dn = O.tricubic_derivs(xv, yv, zv, O.ne2n(ne, 30e6))
    o, d = np.array([0.2, 2.4, 1.5]), np.array([0.05, -0.02, 1.0])
    r = O.bent_ray_rk4(xv, yv, zv, Dn, o, d, 8.5, 12, substeps=8)
    p = d / np.linalg.norm(d)
    Yo = odeint(lambda q, z: O.bent_ray_rhs(xv, yv, zv, Dn, q, z), [p[0], p[1], p[2], o[0], o[1], 0.0],
                np.linspace(o[2], 8.5, 12), rtol=1e-11, atol=1e-12)
    assert np.abs(r[0] - Yo[:, 3]).max() < 1e-6 and np.abs(r[1] - Yo[:, 4]).max() < 1e-6
    assert np.abs(r[3] - Yo[:, 5]).max() < 1e-5
    assert np.abs(r[0] - (o[0] + p[0] / p[2] * (r[2] - o[2]))).max() > 0.05        # the ray really bends
    D1 = O.tricubic_derivs(xv, yv, zv, np.ones_like(ne))
    s = O.bent_ray_rk4(xv, yv, zv, D1, o, d, 8.5, 12, substeps=1)
    x, y, z, sl = O.integrate_ray_straight(o, d, 8.5, 12)
    np.testing.assert_allclose(s, np.stack([x, y, z, sl]), rtol=0, atol=1e-12)
