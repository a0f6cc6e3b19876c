"""ionotomo_b200 -- B200-native ray-integral forward model and adjoint for IonoTomo.

Keeps the reference's call surface for the hot path (``TriCubic``, ``calc_rays`` /
``cast_ray``, ``Fermat``, ``forward_equation``, ``compute_gradient``, ``line_search``);
underneath, Python holds PyTorch tensors as buffers and calls the C ABI of
``libionob200.so`` (hand-written sm_100a CUDA) through ctypes.  No CPU fallback.
"""
from .geometry.tri_cubic import TriCubic, bisection
from .geometry.calc_rays import calc_rays, cast_ray
from .inversion.fermat import Fermat
from .inversion.forward_equation import ForwardProjector, forward_equation, forward_equation_dask
from .inversion.gradient import BackProjector, compute_gradient, compute_gradient_dask, misfit
from .inversion.line_search import line_search, vertex

__all__ = ["TriCubic", "bisection", "calc_rays", "cast_ray", "Fermat", "forward_equation",
           "forward_equation_dask", "ForwardProjector", "BackProjector", "compute_gradient", "compute_gradient_dask", "misfit",
           "line_search", "vertex"]
