"""Ray generation: ``cast_ray`` / ``calc_rays`` of ``geometry/calc_rays.py:61-145``.

Output layout is the reference's: ``rays[Na, Nt, Nd, 4, N]`` (rows x, y, z, s),
float64, C-order -- on the device if the inputs are CUDA tensors.
"""
import numpy as np
import torch

from .. import _lib
from ..inversion.fermat import Fermat


def cast_ray(batch, fermat, tmax, N):
    """``batch = (origins, directions)``, each ``(Na, Nt, Nd, 3)`` in the model frame
    (calc_rays.py:61-96).  One GPU launch instead of one ``odeint`` per ray."""
    origins, directions = batch
    return fermat.cast(origins, directions, tmax, N)


def _is_array(a):
    return isinstance(a, (np.ndarray, torch.Tensor, list, tuple))


def calc_rays(antennas, patches, times, array_center, fixtime, phase, ne_tci, frequency,
              straight_line_approx, tmax, N=None):
    """Same signature as the reference (calc_rays.py:109-145).

    Two kinds of inputs are accepted:

    * arrays already in the model ("pointing") frame -- ``antennas`` ``(Na, 3)`` in km and
      ``patches`` ``(Nd, 3)`` or ``(Nt, Nd, 3)`` direction vectors; ``times`` only gives
      ``Nt`` (``array_center``, ``fixtime``, ``phase`` are unused);
    * astropy coordinate objects, transformed per time step with the caller's
      ``Pointing`` frame exactly as the reference does (needs astropy + the
      reference's frame class importable as ``ionotomo.astro.frames.pointing_frame``).
    """
    if N is None:
        N = ne_tci.nz
    Nt = len(times)
    if _is_array(antennas) and _is_array(patches):
        on_device = isinstance(antennas, torch.Tensor) and antennas.is_cuda
        ants = _lib.to_device(antennas) if on_device else torch.as_tensor(_lib.host_f64(antennas))
        dirs = torch.as_tensor(_lib.host_f64(patches)) if not isinstance(patches, torch.Tensor) else patches
        dirs = dirs.to(ants.device, torch.float64)
        if dirs.dim() == 2:
            dirs = dirs.unsqueeze(0).expand(Nt, -1, -1)
        Na, Nd = ants.shape[0], dirs.shape[1]
        origins = ants[:, None, None, :].expand(Na, Nt, Nd, 3).contiguous()
        directions = dirs[None].expand(Na, Nt, Nd, 3).contiguous()
        want_numpy = not on_device
    else:
        import astropy.units as au
        from ionotomo.astro.frames.pointing_frame import Pointing
        Na, Nd = len(antennas), len(patches)
        origins = np.zeros([Na, Nt, Nd, 3], dtype=np.double)
        directions = np.zeros([Na, Nt, Nd, 3], dtype=np.double)
        for j in range(Nt):
            pointing = Pointing(location=array_center.earth_location, obstime=times[j],
                                fixtime=fixtime, phase=phase)
            d = patches.transform_to(pointing).cartesian.xyz.value.transpose()
            o = antennas.transform_to(pointing).cartesian.xyz.to(au.km).value.transpose()
            origins[:, j, :, :] += np.expand_dims(o, 1)
            directions[:, j, :, :] += d
        want_numpy = True
    fermat = Fermat(ne_tci=ne_tci, frequency=frequency, type='z', straight_line_approx=straight_line_approx)
    rays = cast_ray((_lib.to_device(origins), _lib.to_device(directions)), fermat, tmax, N)
    return rays.cpu().numpy() if want_numpy else rays
