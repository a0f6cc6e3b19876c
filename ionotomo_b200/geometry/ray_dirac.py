"""``get_ray_dirac(rays, tci)`` of ``geometry/ray_dirac.py:5-34``: per ray the chord length of the straight line
first -> last sample through every cell-centred voxel box within +-1 cell of a sample.

The reference allocates a dense ``(N1, N2, nx, ny, nz)`` array (and a ``(3, N1, N2, nx, ny, nz)`` array of segment
midpoints), which limits it to toy sizes; its only consumer (``do_gradient``, ``inversion/gradient.py:15-20``)
contracts it with the residuals at once, which is what ``iono_chord_adjoint_f64`` /
``compute_gradient_chord`` do without materialising it.  This callable exists for parity at those toy sizes: it runs
the same kernel once per ray with a one-hot residual.  The midpoints (unused by every caller in the reference) are
not produced: the second return value is ``None``."""
import numpy as np
import torch

from .. import _lib


def get_ray_dirac(rays, tci):
    want_numpy = not isinstance(rays, torch.Tensor)
    rays_d = _lib.to_device(rays)
    N1, N2, four, Ns = rays_d.shape
    assert four == 4
    shape = (tci.nx, tci.ny, tci.nz)
    dirac = torch.zeros((N1, N2) + shape, dtype=torch.float64, device=rays_d.device)
    one = torch.ones((1, 1, 1), dtype=torch.float64, device=rays_d.device)
    grid = tci.grid()
    for i in range(N1):
        for j in range(N2):
            ray = rays_d[i, j].reshape(1, 1, 1, 4, Ns).contiguous()
            _lib.call("iono_chord_adjoint_f64", grid.handle, _lib.ptr(ray), 1, 1, 1, Ns, _lib.ptr(one), 1,
                      _lib.ptr(dirac[i, j]), _lib.stream_ptr())
    return (dirac.cpu().numpy() if want_numpy else dirac), None
