"""Seeded synthetic problems shared by the tests (NumPy only)."""
import numpy as np


def small_problem(seed, Na, Nt, Nd, Ns, nx, ny, nz, tmax=1000., uniform=True):
    """Near-vertical rays through a smooth positive density field; everything in bounds."""
    rng = np.random.RandomState(seed)
    if uniform:
        xvec = np.linspace(-60., 60., nx)
        yvec = np.linspace(-55., 65., ny)
        zvec = np.linspace(-10., 1010., nz)
    else:
        xvec = np.sort(rng.uniform(-60., 60., nx)); xvec[0], xvec[-1] = -60., 60.
        yvec = np.sort(rng.uniform(-55., 65., ny)); yvec[0], yvec[-1] = -55., 65.
        zvec = np.cumsum(np.linspace(1., 3., nz)); zvec = -10. + (zvec - zvec[0]) * 1020. / (zvec[-1] - zvec[0])
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
    K_ne = float(np.median(ne))
    m = np.log(ne / K_ne)
    return dict(xvec=xvec, yvec=yvec, zvec=zvec, ne=ne, m=m, K_ne=K_ne, origins=origins,
                directions=directions, tmax=tmax, Ns=Ns, rng=rng)
