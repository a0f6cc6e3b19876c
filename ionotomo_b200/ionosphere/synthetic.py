"""Synthetic inputs for tests and benchmarks: LOFAR-like station layout,
directions in a 4-degree field of view tracked over time, a Chapman-layer electron density
with Matern-5/2 turbulence on a box that tightly encloses the rays.

Recipes follow the reference's own generators: ``ionosphere/iri.py:20-68``
(``a_priori_model_``: D/E/F1/F2 Chapman layers vs solar zenith angle),
``ionosphere/simulation.py:45-112`` (``IonosphereSimulation``: Matern-5/2 spectrum, FFT,
sign-flip de-shift, rescale to sigma), ``inversion/initial_model.py:13-36,75-84`` (domain =
ray extent + 20 cells of padding; ``ne * exp(dm)``, sigma = ln 2, corr = 20 km) and
``astro/real_data.py:514-557`` (4 degree field of view, 8 s time steps).  One-off set-up
code: NumPy for the seeded random draws, torch for the heavy array math (on the GPU when
there is one; cuFFT is used as a library here, off the hot path).
"""
import math
import os

import numpy as np
import torch

_DATA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data")
LOFAR_LAT_DEG = 52.914764


def lofar_stations_enu_km():
    """(62, 3) ENU km about the centroid of the reference's ``lofar.hba.antenna.cfg``
    (derived table, see tools/make_lofar_enu.py)."""
    rows = [l.split() for l in open(os.path.join(_DATA, "lofar_hba_enu_km.txt")) if not l.startswith("#")]
    return np.array([[float(r[0]), float(r[1]), float(r[2])] for r in rows])


def chapman_profile(h, zenith=45., thin_f=False):
    """``a_priori_model_`` (ionosphere/iri.py:20-68) on a torch/NumPy array of heights (km)."""
    xp = torch if isinstance(h, torch.Tensor) else np

    def peak_density(n0, dn, tau, b):
        y = zenith / tau
        return n0 + dn * math.exp(-y ** 2) / (1. + y ** (2 * b))

    def peak_height(z0, dz, rho, chi0):
        return z0 + dz / (1. + math.exp(-(zenith - chi0) / rho))

    def layer(nm, zm, H):
        y = (h - zm) / H
        return nm * xp.exp(0.5 * (1. - y - xp.exp(-y)))
    y = zenith / 58.
    nm_d = 4e8 + 5.9e8 * math.exp(-y ** 2) if y < 1 else 4e8
    n = layer(nm_d, peak_height(81., 7., 7.46, 100.), 8.)
    n = n + layer(peak_density(1.6e9, 1.6e11, 87., 8.7), 110., 11.)
    n = n + layer(peak_density(2.0e11, 9.1e10, 54., 13.6), 185., 20. if thin_f else 40.)
    n = n + layer(peak_density(7.7e10, 4.4e11, 111., 4.8), peak_height(242., 75., 7.46, 96.),
                  27.5 if thin_f else 55.)
    return n


def matern52_field(xvec, yvec, zvec, sigma, corr, seed, device=None):
    """``IonosphereSimulation(...,'m52').realization(seed)`` (ionosphere/simulation.py:45-112).
    The Gaussian draws come from ``np.random.seed(seed)`` exactly as in the reference; the
    spectrum shaping and inverse FFT run in torch on ``device``."""
    nx, ny, nz = len(xvec), len(yvec), len(zvec)
    dx, dy, dz = xvec[1] - xvec[0], yvec[1] - yvec[0], zvec[1] - zvec[0]
    sx, sy, sz = 1. / (dx * nx), 1. / (dy * ny), 1. / (dz * nz)
    dev = torch.device(device) if device is not None else torch.device("cpu")
    l = torch.linspace(0, sx * nx / 2., nx, dtype=torch.float64, device=dev)
    m = torch.linspace(0, sy * ny / 2., ny, dtype=torch.float64, device=dev)
    n = torch.linspace(0, sz * nz / 2., nz, dtype=torch.float64, device=dev)
    s2 = l[:, None, None] ** 2 + m[None, :, None] ** 2 + n[None, None, :] ** 2
    s2 = torch.fft.ifftshift(s2)
    nd, nu = 3., 2.5
    S = sigma ** 2 * 2 ** nd * math.pi ** (nd / 2.) * math.gamma(nu + nd / 2.) * (2 * nu) ** nu \
        / math.gamma(nu) / corr ** (2 * nu) * (2 * nu / corr ** 2 + 4 * math.pi ** 2 * s2) ** (-nu - nd / 2.)
    S = torch.sqrt(S)
    np.random.seed(seed)
    zr = np.random.normal(size=(nx, ny, nz))
    zi = np.random.normal(size=(nx, ny, nz))
    Z = torch.complex(torch.as_tensor(zr, device=dev), torch.as_tensor(zi, device=dev))
    B = torch.fft.ifftn(S * Z).real * ((sx * nx) * (sy * ny) * (sz * nz))
    B[::2, :, :] *= -1
    B[:, ::2, :] *= -1
    B[:, :, ::2] *= -1
    B *= sigma / torch.std(B, unbiased=False)
    return B


def directions_in_fov(Nd, fov_deg=4., seed=1234):
    """Nd unit vectors uniform over a disc of diameter ``fov_deg`` about the frame's z axis."""
    rng = np.random.RandomState(seed)
    r = np.radians(fov_deg / 2.) * np.sqrt(rng.uniform(size=Nd))
    phi = 2 * np.pi * rng.uniform(size=Nd)
    return np.stack([np.sin(r) * np.cos(phi), np.sin(r) * np.sin(phi), np.cos(r)], -1)


def track_directions(dirs, Nt, dt_s=8., lat_deg=LOFAR_LAT_DEG):
    """Earth rotation without astropy: rotate the directions about the celestial pole (ENU
    components (0, cos lat, sin lat)) by 15 arcsec/s, fixed at the middle time step
    (the reference fixes its frame at ``times[Nt>>1]``, inversion_pipeline.py / tests)."""
    lat = np.radians(lat_deg)
    k = np.array([0., np.cos(lat), np.sin(lat)])
    out = np.empty((Nt,) + dirs.shape)
    for j in range(Nt):
        ang = np.radians(15. / 3600.) * dt_s * (j - (Nt >> 1))
        c, s = np.cos(ang), np.sin(ang)
        out[j] = dirs * c + np.cross(k, dirs) * s + np.outer(dirs @ k, k) * (1 - c)
    return out


def tight_axes(ants, dirs_t, nx, ny, nz, tmax=1000., pad_cells=20, zlim=(-100., 1100.)):
    """Grid axes enclosing every ray with ``pad_cells`` cells of padding horizontally
    (inversion/initial_model.py:13-36) -- the 'tight box' of SURVEY.md §8d."""
    d = dirs_t.reshape(-1, 3)
    ends_x = (ants[:, 0][:, None] + d[None, :, 0] / d[None, :, 2] * (tmax - ants[:, 2][:, None]))
    ends_y = (ants[:, 1][:, None] + d[None, :, 1] / d[None, :, 2] * (tmax - ants[:, 2][:, None]))
    axes = []
    for lo, hi, n in ((min(ants[:, 0].min(), ends_x.min()), max(ants[:, 0].max(), ends_x.max()), nx),
                      (min(ants[:, 1].min(), ends_y.min()), max(ants[:, 1].max(), ends_y.max()), ny)):
        dcell = (hi - lo) / (n - 1 - 2 * pad_cells)
        axes.append(np.linspace(lo - pad_cells * dcell, hi + pad_cells * dcell, n))
    axes.append(np.linspace(zlim[0], zlim[1], nz))
    return axes


def make_workload(Na=62, Nt=100, Nd=200, nx=256, ny=256, nz=128, seed=1234, device="cuda", tmax=1000.,
                  isotropic_spacing=None, t_slice=None, d_slice=None):
    """The LOFAR-like benchmark case (BASELINE.json configs[1..2]) as device tensors.

    ``t_slice=(t0, t1)`` / ``d_slice=(d0, d1)`` keep only that block of time steps / directions
    (ray sharding across GPUs: the grid and the full-problem geometry are identical on every
    rank; both axes keep the reference antenna local)."""
    ants = lofar_stations_enu_km()[:Na]
    dirs_t = track_directions(directions_in_fov(Nd, 4., seed), Nt)
    if isotropic_spacing is None:
        xvec, yvec, zvec = tight_axes(ants, dirs_t, nx, ny, nz, tmax)
    else:
        sp = float(isotropic_spacing)
        xvec = (np.arange(nx) - (nx - 1) / 2.) * sp + ants[:, 0].mean()
        yvec = (np.arange(ny) - (ny - 1) / 2.) * sp + ants[:, 1].mean()
        zvec = np.linspace(-100., 1100., nz)
    dev = torch.device(device)
    dm = matern52_field(xvec, yvec, zvec, math.log(2.), 20., seed, device=dev)
    z = torch.as_tensor(zvec, device=dev)
    ne_prior = chapman_profile(z, 45.)[None, None, :].expand(nx, ny, nz).contiguous()
    ne_true = ne_prior * torch.exp(dm)
    K_ne = float(ne_true.mean())
    t0, t1 = (0, Nt) if t_slice is None else t_slice
    d0, d1 = (0, Nd) if d_slice is None else d_slice
    Nd_total, Nd = Nd, d1 - d0
    o = torch.as_tensor(ants, device=dev)[:, None, None, :].expand(Na, t1 - t0, Nd, 3).contiguous()
    d = torch.as_tensor(dirs_t[t0:t1, d0:d1], device=dev)[None].expand(Na, t1 - t0, Nd, 3).contiguous()
    return dict(xvec=xvec, yvec=yvec, zvec=zvec, K_ne=K_ne, m_true=torch.log(ne_true / K_ne),
                m_prior=torch.log(ne_prior / K_ne), origins=o, directions=d, tmax=tmax, Ns=nz,
                Na=Na, Nt=t1 - t0, Nd=Nd, Nt_total=Nt, Nd_total=Nd_total, seed=seed,
                dx_km=float(xvec[1] - xvec[0]), dy_km=float(yvec[1] - yvec[0]), dz_km=float(zvec[1] - zvec[0]))
