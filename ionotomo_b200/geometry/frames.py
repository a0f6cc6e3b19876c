"""Frame math for ray generation without astropy: the rotation of the reference's ``Pointing``
frame (``astro/frames/pointing_frame.py:140-190``) and a frame-aware ``calc_rays`` that takes
ITRS antenna positions and ITRS direction vectors and applies the per-time rotation on the GPU,
one thread per ray (``iono_cast_rays_frames_f64``).

What is *not* reproduced here is astropy's ICRS -> ITRS chain (precession, nutation, polar
motion): callers either pass ITRS direction vectors computed with astropy, or use
``icrs_to_itrs_simple`` (Earth rotation angle only; fine for synthetic data, a few arcsec off
for real pointings).
"""
import numpy as np
import torch

from .. import _lib


def pointing_rotation(lon_rad, ha_rad, dec_rad):
    """``R = [east; north; up]`` with ``lonrad = lon - HA`` and ``latrad = dec``
    (pointing_frame.py:151-166).  Scalars or arrays of equal shape -> ``(..., 3, 3)``."""
    lonrad = np.asarray(lon_rad, dtype=np.float64) - np.asarray(ha_rad, dtype=np.float64)
    dec = np.broadcast_to(np.asarray(dec_rad, dtype=np.float64), lonrad.shape)
    sinlat, coslat = np.sin(dec), np.cos(dec)
    sinlon, coslon = np.sin(lonrad), np.cos(lonrad)
    zero = np.zeros_like(sinlon)
    east = np.stack([-sinlon, coslon, zero], -1)
    north = np.stack([-sinlat * coslon, -sinlat * sinlon, coslat], -1)
    up = np.stack([coslat * coslon, coslat * sinlon, sinlat], -1)
    return np.stack([east, north, up], -2)


def gmst_rad(jd_ut1):
    """Greenwich mean sidereal time (IAU 1982 polynomial), radians."""
    T = (np.asarray(jd_ut1, dtype=np.float64) - 2451545.0) / 36525.0
    sec = 67310.54841 + (876600.0 * 3600.0 + 8640184.812866) * T + 0.093104 * T ** 2 - 6.2e-6 * T ** 3
    return np.mod(sec, 86400.0) * (2 * np.pi / 86400.0)


def icrs_to_itrs_simple(ra_rad, dec_rad, jd_ut1):
    """Unit vectors ``(Nt, Nd, 3)`` in the Earth-fixed frame from (ra, dec) by the Earth rotation
    angle alone (no precession/nutation/polar motion)."""
    g = np.atleast_1d(gmst_rad(jd_ut1))[:, None]
    ra = np.asarray(ra_rad, dtype=np.float64)[None, :]
    dec = np.asarray(dec_rad, dtype=np.float64)[None, :]
    lon = ra - g
    return np.stack([np.cos(dec) * np.cos(lon), np.cos(dec) * np.sin(lon), np.sin(dec) * np.ones_like(lon)], -1)


def calc_rays_itrs(antennas_itrs_m, dirs_itrs, R, array_center_itrs_m, tmax, N):
    """Rays ``(Na, Nt, Nd, 4, N)`` from ITRS antenna positions ``(Na, 3)`` in metres, ITRS unit
    direction vectors ``(Nt, Nd, 3)`` and per-time pointing rotations ``R (Nt, 3, 3)``: the loop of
    ``calc_rays`` (geometry/calc_rays.py:125-143) as one GPU launch.  CUDA tensors in -> CUDA
    tensor out, NumPy in -> NumPy out."""
    want_numpy = not isinstance(antennas_itrs_m, torch.Tensor)
    a = _lib.to_device(antennas_itrs_m)
    d = _lib.to_device(dirs_itrs, a.device)
    Rd = _lib.to_device(R, a.device)
    p0 = _lib.to_device(array_center_itrs_m, a.device).reshape(3)
    Na, Nt, Nd = a.shape[0], d.shape[0], d.shape[1]
    assert tuple(a.shape) == (Na, 3) and tuple(d.shape) == (Nt, Nd, 3) and tuple(Rd.shape) == (Nt, 3, 3)
    rays = torch.empty((Na, Nt, Nd, 4, int(N)), dtype=torch.float64, device=a.device)
    _lib.call("iono_cast_rays_frames_f64", _lib.ptr(a), _lib.ptr(p0), _lib.ptr(Rd), _lib.ptr(d), Na, Nt, Nd,
              float(tmax), int(N), _lib.ptr(rays), _lib.stream_ptr())
    return rays.cpu().numpy() if want_numpy else rays
