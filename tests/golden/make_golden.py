#!/usr/bin/env python
"""Generate tests/golden/*.npz by executing the REFERENCE's own modules.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference package cannot be imported as a whole here (astropy, h5py, dask,
tensorflow, matplotlib are absent and ``scipy.integrate.simps`` was removed from
SciPy >= 1.14), so this script

* registers an empty ``ionotomo`` namespace whose ``__path__`` is the
  reference's source directory, so that the hot-path leaf modules
  (``geometry/tri_cubic.py``, ``inversion/fermat.py``, ``geometry/calc_rays.py``,
  ``inversion/forward_equation.py``, ``inversion/iterative_newton.py``,
  ``geometry/ray_dirac.py``, ``geometry/slab_method.py``, ``inversion/gradient.py``,
  ``inversion/line_search.py``, ``inversion/gradient_and_adjoint.py``,
  ``ionosphere/simulation.py``, ``ionosphere/iri.py``)
  are executed UNMODIFIED from where they lie;
* satisfies their imports of absent third-party packages with inert stub
  modules (none of the stubs is reached by the functions called below);
* provides ``scipy.integrate.simps`` built ONLY from the installed SciPy:
  ``simpson`` for an odd number of samples (identical to the old ``simps``
  there) and, for an even number, the documented ``even='avg'`` definition
  (tomography/integrate.py:91-99) composed from ``simpson`` on the two odd-length
  sub-ranges plus the end trapezoids.

Nothing from ``oracle/`` or ``ionotomo_b200/`` is imported, so the fixtures are
independent of the code they check.  All inputs are seeded.
"""
import importlib.abc
import importlib.machinery
import os
import sys
import types

import numpy as np
import scipy.integrate

REF_SRC = "/root/reference/src"
OUT = os.path.dirname(os.path.abspath(__file__))

STUB_ROOTS = ("astropy", "h5py", "dask", "pylab", "matplotlib", "tensorflow",
              "pyiri2016", "keras", "gpflow", "mayavi")
STUB_EXACT = ("ionotomo.astro.frames.pointing_frame", "ionotomo.astro.real_data",
              "ionotomo.inversion.solution", "ionotomo.ionosphere.covariance",
              "ionotomo.plotting.plot_tools")


class _Anything(object):
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        return _Anything()


class _StubModule(types.ModuleType):
    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] in STUB_ROOTS or fullname in STUB_EXACT:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _StubModule(spec.name)

    def exec_module(self, module):
        pass


# scipy.integrate.simps is gone from the installed SciPy: the reference's modules get the routine they were
# written against, transcribed from the SciPy releases of their time (old_scipy_simps.py)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from old_scipy_simps import simps as _simps  # noqa: E402


def install():
    sys.meta_path.insert(0, _StubFinder())
    scipy.integrate.simps = _simps
    if not hasattr(np, "bool"):
        np.bool = bool  # inversion/gradient_and_adjoint.py:21 (NumPy < 1.24 name)
    pkg = types.ModuleType("ionotomo")
    pkg.__path__ = [os.path.join(REF_SRC, "ionotomo")]
    sys.modules["ionotomo"] = pkg


def small_problem(seed, Na, Nt, Nd, Ns, nx, ny, nz):
    """Seeded toy set-up: near-vertical rays through a smooth positive field."""
    rng = np.random.RandomState(seed)
    xvec = np.linspace(-60., 60., nx)
    yvec = np.linspace(-55., 65., ny)
    zvec = np.linspace(-10., 1010., nz)
    X, Y, Z = np.meshgrid(xvec, yvec, zvec, indexing='ij')
    ne = 1e11 * np.exp(-((Z - 300.) / 150.) ** 2) * (1. + 0.3 * np.sin(X / 20.) * np.cos(Y / 25.)) + 1e9
    ne *= np.exp(0.2 * rng.normal(size=ne.shape))
    ants = np.stack([rng.uniform(-20, 20, Na), rng.uniform(-20, 20, Na), rng.uniform(-0.3, 0.1, Na)], -1)
    dirs = np.stack([rng.uniform(-0.02, 0.02, (Nt, Nd)), rng.uniform(-0.02, 0.02, (Nt, Nd)),
                     np.ones((Nt, Nd))], -1)
    origins = np.zeros((Na, Nt, Nd, 3))
    directions = np.zeros((Na, Nt, Nd, 3))
    origins += ants[:, None, None, :]
    directions += dirs[None]
    return xvec, yvec, zvec, ne, origins, directions, rng


def golden_gauss():
    """9. adjoint B (gradient_and_adjoint.py:12-167): Gaussian model covariance along the rays."""
    from ionotomo.geometry.tri_cubic import TriCubic
    from ionotomo.inversion.fermat import Fermat
    from ionotomo.geometry.calc_rays import cast_ray
    from ionotomo.inversion.forward_equation import forward_equation
    from ionotomo.inversion.gradient_and_adjoint import do_adjoint, compute_adjoint
    Ns = 14
    xvec, yvec, zvec, ne, origins, directions, rng = small_problem(21, 3, 2, 2, Ns, 10, 9, 12)
    origins[0, ..., 0] = 38.          # boxes clipped at the upper x edge
    origins[2, ..., 0] = -38.         # and at the lower one
    origins[2, ..., 1] = 43.          # upper y edge
    ne_tci = TriCubic(xvec, yvec, zvec, ne)
    fermat = Fermat(ne_tci=ne_tci, frequency=120e6, type='z', straight_line_approx=True)
    rays = cast_ray((origins, directions), fermat, 1000., Ns)
    K_ne = np.median(ne)
    m_tci = ne_tci.copy()
    m_tci.M = np.log(m_tci.M / K_ne)
    i0 = 1
    g = forward_equation(rays, K_ne, m_tci, i0)
    dobs = g + 0.05 * rng.normal(size=g.shape)
    CdCt = (0.01 + 0.01 * rng.uniform(size=g.shape)) ** 2
    m_prior = m_tci.M + 0.1 * rng.normal(size=m_tci.M.shape)
    sigma_m, Nkernel, size_cell = 0.7, 2, 20.
    dd = (g - dobs) / (CdCt + 1e-15)
    slice0 = do_adjoint(rays[:, :, 0], dd[:, :, 0], K_ne, m_tci, sigma_m, Nkernel, size_cell, i0)
    adj = compute_adjoint(rays, g.copy(), dobs, i0, K_ne, m_tci, m_prior, CdCt, sigma_m, Nkernel, size_cell)
    # a second kernel size whose boxes reach both ends of every axis
    adj_wide = compute_adjoint(rays[:2, :1], g[:2, :1].copy(), dobs[:2, :1], 0, K_ne, m_tci, m_prior,
                               CdCt[:2, :1], 1.3, 5, 7.)
    np.savez(os.path.join(OUT, "adjoint_gauss.npz"), xvec=xvec, yvec=yvec, zvec=zvec, m=m_tci.M, K_ne=K_ne,
             rays=rays, g=g, dobs=dobs, CdCt=CdCt, m_prior=m_prior, i0=i0, sigma_m=sigma_m, Nkernel=Nkernel,
             size_cell=size_cell, slice0=slice0, adj=adj, adj_wide=adj_wide)


def main():
    install()
    if "--only-gauss" in sys.argv:
        golden_gauss()
        return
    from ionotomo.geometry.tri_cubic import TriCubic, bisection
    from ionotomo.inversion.fermat import Fermat
    from ionotomo.geometry.calc_rays import cast_ray
    from ionotomo.inversion.forward_equation import forward_equation, do_forward_equation
    import ionotomo.inversion.iterative_newton as newton
    from ionotomo.geometry.ray_dirac import get_ray_dirac
    from ionotomo.inversion.gradient import do_gradient
    import ionotomo.inversion.line_search as ls
    from ionotomo.ionosphere.simulation import IonosphereSimulation
    from ionotomo.ionosphere.iri import a_priori_model_

    # 1. TriCubic.interp / extrapolate / bisection on the field of tests/test_tricubic.py:6-12
    xvec = np.linspace(-0.1, 1.1, 100)
    yvec = np.linspace(-0.1, 1.1, 100)
    zvec = np.linspace(-0.1, 1.11, 100)
    x, y, z = np.meshgrid(xvec, yvec, zvec, indexing='ij')
    M = x * y * z + x - y - 2 * z + x ** 2
    tci = TriCubic(xvec, yvec, zvec, M)
    pts = np.array([1. / i for i in range(1, 1000)])
    res_batch = tci.interp(pts, pts, pts)
    res_scalar = np.array([tci.interp(p, p, p) for p in pts[:50]])
    rng = np.random.RandomState(7)
    px, py, pz = (rng.uniform(-0.1, 1.1, 500), rng.uniform(-0.1, 1.1, 500),
                  rng.uniform(-0.1, 1.11, 500))
    # exact nodes and the domain corners
    px[:5] = [xvec[0], xvec[-1], xvec[17], xvec[50], xvec[99]]
    py[:5] = [yvec[0], yvec[-1], yvec[3], yvec[98], yvec[0]]
    pz[:5] = [zvec[0], zvec[-1], zvec[64], zvec[1], zvec[99]]
    res_rand = tci.interp(px, py, pz)
    ex = np.array([1.1, 2., -0.5, 1.3, 0.5])
    ey = np.array([1.1, 2., 0.2, -0.3, 0.5])
    ez = np.array([1.1, 2., 1.5, 0.4, -1.0])
    res_extrap = tci.extrapolate(ex, ey, ez)
    oob_raises = False
    try:
        tci.interp(np.array([0.5, 1.2]), np.array([0.5, 0.5]), np.array([0.5, 0.5]))
    except ValueError:
        oob_raises = True
    bvals = np.concatenate([rng.uniform(-0.2, 1.2, 40), xvec[[0, 1, 50, 98, 99]]])
    bidx = np.array([bisection(xvec, v) for v in bvals])
    np.savez(os.path.join(OUT, "tricubic.npz"), xvec=xvec, yvec=yvec, zvec=zvec, pts=pts,
             res_batch=res_batch, res_scalar=res_scalar, px=px, py=py, pz=pz, res_rand=res_rand,
             ex=ex, ey=ey, ez=ez, res_extrap=res_extrap, oob_raises=oob_raises,
             bvals=bvals, bidx=bidx)

    # 2..5 on small seeded problems, odd and even Ns
    for tag, Ns in (("odd", 9), ("even", 10)):
        xvec, yvec, zvec, ne, origins, directions, rng = small_problem(11, 3, 2, 4, Ns, 12, 11, 10)
        ne_tci = TriCubic(xvec, yvec, zvec, ne)
        fermat = Fermat(ne_tci=ne_tci, frequency=120e6, type='z', straight_line_approx=True)
        rays = cast_ray((origins, directions), fermat, 1000., Ns)      # scipy odeint per ray
        single = np.stack(fermat.integrate_ray(origins[1, 0, 2], directions[1, 0, 2], 1000., N=Ns))
        K_ne = np.median(ne)
        m_tci = ne_tci.copy()
        m_tci.M = np.log(m_tci.M / K_ne)
        i0 = 1
        dtec = forward_equation(rays, K_ne, m_tci, i0)
        ne_scaled = ne_tci.copy()
        ne_scaled.M = np.exp(m_tci.M) * (K_ne / 1e13)
        tec_a0 = do_forward_equation(rays[0], ne_scaled)
        # phase variant (iterative_newton.py:86-127), K=1e11
        freqs = np.array([120e6, 150e6])
        mu = np.log(ne / 1e11).flatten()
        clock = 1e-9 * rng.normal(size=(3, 2))
        const = 0.1 * rng.normal(size=3)
        tci_b = ne_tci.copy()
        phase = newton.forward_equation((mu.copy(), clock, const), tci_b, rays, freqs, K=1e11, i0=i0)
        mu_prior = mu + 0.05 * rng.normal(size=mu.shape)
        tci_c = ne_tci.copy()
        penalty = newton.prior_penalty_mu((mu.copy(), clock, const),
                                          (mu_prior, clock, const), tci_c, rays, freqs,
                                          K=1e11, i0=i0)
        np.savez(os.path.join(OUT, "forward_%s.npz" % tag), xvec=xvec, yvec=yvec, zvec=zvec,
                 ne=ne, origins=origins, directions=directions, rays=rays, single=single,
                 single_origin=origins[1, 0, 2], single_direction=directions[1, 0, 2],
                 K_ne=K_ne, m=m_tci.M, i0=i0, dtec=dtec, tec_a0=tec_a0, freqs=freqs, mu=mu,
                 clock=clock, const=const, phase=phase, mu_prior=mu_prior, penalty=penalty,
                 tmax=1000., Ns=Ns)

    # 6. chord-length adjoint A8 (ray_dirac.py + slab_method.py + gradient.py:15-20), toy size
    xvec, yvec, zvec, ne, origins, directions, rng = small_problem(5, 2, 1, 3, 7, 7, 6, 8)
    zvec = np.linspace(-100., 1100., 8)
    ne_tci = TriCubic(xvec, yvec, zvec, ne)
    fermat = Fermat(ne_tci=ne_tci, frequency=120e6, type='z', straight_line_approx=True)
    rays = cast_ray((origins, directions), fermat, 1000., 7)
    dirac, _ = get_ray_dirac(rays[:, 0], ne_tci)          # (Na, Nd, nx, ny, nz)
    dd = rng.normal(size=(2, 3))
    G = do_gradient(rays[:, 0], dd, ne_tci, 1., 3, 5., 0)
    np.savez(os.path.join(OUT, "chord.npz"), xvec=xvec, yvec=yvec, zvec=zvec, ne=ne, rays=rays,
             dirac=dirac, dd=dd, G=G)

    # 7. line search + vertex (line_search.py:13-100); the final dask call is the same
    #    function as the serial one (tests/test_forward_equation.py:26-27 asserts equality)
    ls.forward_equation_dask = ls.forward_equation
    TriCubic.get_shaped_array = lambda self: self.M      # older accessor used at line_search.py:46
    xvec, yvec, zvec, ne, origins, directions, rng = small_problem(3, 4, 1, 5, 11, 12, 11, 10)
    ne_tci = TriCubic(xvec, yvec, zvec, ne)
    fermat = Fermat(ne_tci=ne_tci, frequency=120e6, type='z', straight_line_approx=True)
    rays = cast_ray((origins, directions), fermat, 1000., 11)
    K_ne = np.mean(ne)
    m_true = np.log(ne / K_ne)
    m_tci_true = TriCubic(xvec, yvec, zvec, m_true)
    dobs = forward_equation(rays, K_ne, m_tci_true, 0)
    m0 = m_true + 0.1 * rng.normal(size=m_true.shape)
    m_tci = TriCubic(xvec, yvec, zvec, m0)
    g = forward_equation(rays, K_ne, m_tci, 0)
    CdCt = (0.01 * np.ones(dobs.shape)) ** 2
    grad = rng.normal(size=m0.shape) * 1e-3   # any descent-ish direction exercises the code path
    # use a true descent direction: finite differences are too slow; take the sign from a probe
    S0 = np.sum((g - dobs) ** 2 / (CdCt + 1e-15)) / 2.
    g_probe = forward_equation(rays, K_ne, TriCubic(xvec, yvec, zvec, m0 - 1e-3 * grad), 0)
    if np.sum((g_probe - dobs) ** 2 / (CdCt + 1e-15)) / 2. > S0:
        grad = -grad
    import io
    import contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        eps, S, red = ls.line_search(rays, K_ne, m_tci, 0, grad, g.copy(), dobs, CdCt)
    vx = np.array([0.3, 0.11, 0.7, 1e-3, 5e-4, 2.5e-4])
    vy = np.array([1.7, 1.1, 2.9, 10.2, 9.7, 9.9])
    v1 = ls.vertex(*vx[:3], *vy[:3])
    v2 = ls.vertex(*vx[3:], *vy[3:])
    np.savez(os.path.join(OUT, "line_search.npz"), xvec=xvec, yvec=yvec, zvec=zvec, rays=rays,
             K_ne=K_ne, m0=m0, dobs=dobs, g=g, CdCt=CdCt, grad=grad, eps=eps, S=S, red=red,
             S0=S0, vx=vx, vy=vy, v1=np.array(v1), v2=np.array(v2))

    # 7b. the shipped "curved" mode: Fermat(straight_line_approx=False) -> odeint with sdot = n/pz
    xvec, yvec, zvec, ne, origins, directions, rng = small_problem(13, 2, 1, 3, 12, 12, 11, 10)
    ne_tci = TriCubic(xvec, yvec, zvec, ne * 20.)
    fermat_c = Fermat(ne_tci=ne_tci, frequency=40e6, type='z', straight_line_approx=False)
    rays_c = cast_ray((origins, directions), fermat_c, 600., 12)       # tmax inside the grid: LSODA probes ahead
    np.savez(os.path.join(OUT, "optical_path.npz"), xvec=xvec, yvec=yvec, zvec=zvec, ne=ne * 20., frequency=40e6,
             origins=origins, directions=directions, rays=rays_c, tmax=600., Ns=12, n_field=fermat_c.n_tci.M)

    # 7c. BASELINE.json configs[0]: 10 antennas x 20 directions x 1 time, 50x50x30 grid, Ns = nz = 30
    xvec, yvec, zvec, ne, origins, directions, rng = small_problem(1, 10, 1, 20, 30, 50, 50, 30)
    ne_tci = TriCubic(xvec, yvec, zvec, ne)
    fermat = Fermat(ne_tci=ne_tci, frequency=120e6, type='z', straight_line_approx=True)
    rays = cast_ray((origins, directions), fermat, 1000., ne_tci.nz)
    K_ne = np.median(ne)
    m_tci = ne_tci.copy()
    m_tci.M = np.log(m_tci.M / K_ne)
    dtec = forward_equation(rays, K_ne, m_tci, 0)
    np.savez_compressed(os.path.join(OUT, "config1.npz"), xvec=xvec, yvec=yvec, zvec=zvec, m=m_tci.M, K_ne=K_ne,
                        origins=origins[:, 0, 0, :], directions=directions[0], rays=rays, dtec=dtec)

    # 8. synthetic-input recipe (ionosphere/simulation.py:45-112, ionosphere/iri.py:20-68)
    xvec = np.linspace(-50., 50., 16)
    yvec = np.linspace(-40., 40., 12)
    zvec = np.linspace(-100., 1100., 20)
    sim = IonosphereSimulation(xvec, yvec, zvec, np.log(2.), 20., type='m52')
    dm = sim.realization(seed=1234)
    h = np.linspace(0., 1000., 101)
    chap45 = a_priori_model_(h, 45.)
    chap80 = a_priori_model_(h, 80., thin_f=True)
    np.savez(os.path.join(OUT, "synthetic.npz"), xvec=xvec, yvec=yvec, zvec=zvec, dm=dm, h=h,
             chap45=chap45, chap80=chap80)
    golden_gauss()
    golden_simps_even()
    golden_fermat_s()
    print("golden vectors written to", OUT)


def golden_simps_even():
    """Old SciPy's simps on NON-uniform abscissae, even and odd N, all three `even` modes: pins the oracle's and the
    kernels' Simpson weights beyond the uniform-x numbers of the old docstring."""
    rng = np.random.RandomState(77)
    out = {}
    for N in (2, 3, 4, 5, 6, 9, 10, 30, 31, 64, 128, 129, 256):
        x = np.cumsum(rng.uniform(0.3, 2.5, size=(5, N)), axis=1)
        y = rng.normal(size=(5, N)) + np.sin(x / 7.)
        out["x%d" % N], out["y%d" % N] = x, y
        out["avg%d" % N] = _simps(y, x, axis=1, even='avg')
        if N % 2 == 0:
            out["first%d" % N] = _simps(y, x, axis=1, even='first')
            out["last%d" % N] = _simps(y, x, axis=1, even='last')
    x10 = np.arange(0, 10)
    out["doc_avg"], out["doc_first"] = _simps(np.power(x10, 3), x10), _simps(np.power(x10, 3), x10, even='first')
    np.savez(os.path.join(OUT, "simps_even.npz"), **out)


def golden_fermat_s():
    """Fermat(type='s'): arc length as the independent variable (inversion/fermat.py:74-82,163-166), straight rays."""
    from ionotomo.geometry.tri_cubic import TriCubic
    from ionotomo.inversion.fermat import Fermat
    xvec, yvec, zvec, ne, origins, directions, rng = small_problem(21, 3, 2, 4, 15, 10, 9, 11)
    ne_tci = TriCubic(xvec, yvec, zvec, ne)
    fermat = Fermat(ne_tci=ne_tci, frequency=120e6, type='s', straight_line_approx=True)
    rays = np.zeros(origins.shape[:3] + (4, 15))
    for i in range(origins.shape[0]):
        for j in range(origins.shape[1]):
            for k in range(origins.shape[2]):
                rays[i, j, k] = np.stack(fermat.integrate_ray(origins[i, j, k], directions[i, j, k], 950., N=15))
    np.savez(os.path.join(OUT, "fermat_s.npz"), origins=origins, directions=directions, rays=rays, tmax=950., Ns=15)


if __name__ == "__main__":
    install()
    if len(sys.argv) > 1 and sys.argv[1] == "--extra-only":     # the two fixtures added in round 2
        golden_simps_even()
        golden_fermat_s()
    else:
        main()
