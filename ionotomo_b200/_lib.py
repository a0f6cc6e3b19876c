"""ctypes binding of libionob200.so (include/ionob200.h).

There is deliberately no CPU fallback: if the library is missing or a call fails,
the Python shims raise.
"""
import ctypes
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("IONO_LIB") or os.path.join(_HERE, "libionob200.so")   # IONO_LIB: kernel-variant builds (tools/)

IONO_OK, IONO_EBADARG, IONO_EOOB, IONO_ECUDA = 0, 1, 2, 3
ORDER_NATURAL, ORDER_TIME, ORDER_ANTENNA = 0, 1, 2
ORDERS = {"natural": ORDER_NATURAL, "time": ORDER_TIME, "antenna": ORDER_ANTENNA}

_vp = ctypes.c_void_p
_i = ctypes.c_int
_i64 = ctypes.c_int64
_d = ctypes.c_double

# name -> (restype, argtypes); every symbol include/ionob200.h declares
SIGNATURES = {
    "iono_version": (_i, []),
    "iono_last_error": (ctypes.c_char_p, []),
    "iono_grid_create": (_i, [_vp, _vp, _vp, _i, _i, _i, ctypes.POINTER(_vp)]),
    "iono_grid_destroy": (_i, [_vp]),
    "iono_grid_is_uniform": (_i, [_vp]),
    "iono_ne_from_m_f64": (_i, [_vp, _i64, _d, _vp, _vp]),
    "iono_mul_f64": (_i, [_vp, _vp, _i64, _vp, _vp]),
    "iono_cast_rays_straight_f64": (_i, [_vp, _vp, _i64, _d, _i, _vp, _vp]),
    "iono_cast_rays_arclength_f64": (_i, [_vp, _vp, _i64, _d, _i, _vp, _vp]),
    "iono_cast_rays_frames_f64": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _d, _i, _vp, _vp]),
    "iono_ne_to_refractive_index_f64": (_i, [_vp, _i64, _d, _vp, _vp]),
    "iono_optical_path_f64": (_i, [_vp, _vp, _vp, _i64, _i, _vp, _vp]),
    "iono_tricubic_derivs_f64": (_i, [_vp, _vp, _vp, _vp]),
    "iono_tricubic_interp_f64": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    "iono_bent_rays_f64": (_i, [_vp, _vp, _vp, _vp, _i64, _d, _i, _i, _vp, _vp, _vp]),
    "iono_tci_interp_f64": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _i, _vp, _vp, _vp]),
    "iono_tec_forward_f64": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "iono_quads_from_ne_f64": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "iono_ne_quads_from_m_f64": (_i, [_vp, _i, _i, _i, _d, _vp, _vp, _vp]),
    "iono_tec_forward_quads_f64": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "iono_residual_scratch_elems": (_i64, [_i, _i, _i]),
    "iono_residual_f64": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "iono_dtec_f64": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "iono_adjoint_coef_f64": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "iono_tec_adjoint_f64": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _i, _i, _vp, _vp, _vp]),
    "iono_misfit_scratch_elems": (_i64, []),
    "iono_misfit_f64": (_i, [_vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    "iono_convolve3d_nearest_f64": (_i, [_vp, _i, _i, _i, _vp, _i, _vp, _vp]),
    "iono_copy2d_h2d": (_i, [_vp, _i64, _vp, _i64, _i64, _i64, _vp]),
    "iono_chord_adjoint_f64": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _i, _vp, _vp]),
    "iono_gaussian_adjoint_f64": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _d, _d, _i, _i, _vp, _vp]),
    "iono_phase_integrals_f64": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _i, _i, _vp, _vp, _vp]),
    "iono_simps_rows_f64": (_i, [_vp, _vp, _vp, _i64, _i, _i, _d, _vp, _i, _vp]),
    "iono_phase_assemble_f64": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp]),
    "iono_backprojector_create": (_i, [_vp, _vp, _i, _i, _i, _i, ctypes.POINTER(_vp), _vp, _vp]),
    "iono_backprojector_apply_f64": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "iono_backprojector_apply_chunks_f64": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp]),
    "iono_backprojector_apply_permuted_f64": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp]),
    "iono_backprojector_apply_gradient_f64": (_i, [_vp, _vp, _vp, _d, _vp, _i, _i, _vp]),
    "iono_backprojector_apply_compact_f64": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp]),
    "iono_backprojector_ne_rows_f64": (_i, [_vp, _vp, _d, _vp, _vp]),
    "iono_backprojector_n_rows": (ctypes.c_longlong, [_vp]),
    "iono_backprojector_row_voxels": (_i, [_vp, _vp, _vp]),
    "iono_backprojector_chunk_voxels": (ctypes.c_longlong, [_vp, _i]),
    "iono_backprojector_nnz": (ctypes.c_longlong, [_vp]),
    "iono_backprojector_bytes": (ctypes.c_longlong, [_vp]),
    "iono_backprojector_destroy": (_i, [_vp]),
    "iono_multi_dot_scratch_elems": (_i64, []),
    "iono_multi_dot_f64": (_i, [_vp, _i64, _i, _vp, _vp, _i64, _vp, _vp, _vp]),
    "iono_multi_dot3_f64": (_i, [_vp, _i64, _i, _i, _i, _i, _vp, _i64, _vp, _vp, _vp]),
    "iono_lincomb_f64": (_i, [_vp, _i64, _i, _vp, _vp, _i64, _vp, _vp]),
    "iono_zero_f64": (_i, [_vp, _i64, _vp]),
    "iono_gather_f64": (_i, [_vp, _vp, _i64, _vp, _vp]),
    "iono_scatter_set_f64": (_i, [_vp, _vp, _i64, _vp, _vp]),
    "iono_scatter_axpy_f64": (_i, [_vp, _vp, _vp, _vp, _i64, _vp, _vp]),
    "iono_peer_alloc": (_i, [_i64, ctypes.POINTER(_vp), _vp]),
    "iono_peer_open": (_i, [_vp, ctypes.POINTER(_vp)]),
    "iono_peer_close": (_i, [_vp]),
    "iono_peer_free": (_i, [_vp]),
    "iono_peer_flag_bytes": (_i64, []),
    "iono_peer_reduce_expand_f64": (_i, [_vp, _vp, _vp, _i, _i, _i64, _vp, _i64, _vp, _d, _vp, _vp, _vp]),
    "iono_forwardprojector_create": (_i, [_vp, _vp, _i, _i, _i, _i, ctypes.POINTER(_vp), _vp, _vp]),
    "iono_forwardprojector_apply_f64": (_i, [_vp, _vp, _vp, _vp]),
    "iono_forwardprojector_apply_quads_f64": (_i, [_vp, _vp, _vp, _vp]),
    "iono_forwardprojector_quads_from_m_f64": (_i, [_vp, _vp, _d, _vp, _vp]),
    "iono_forwardprojector_n_records": (ctypes.c_longlong, [_vp]),
    "iono_forwardprojector_bytes": (ctypes.c_longlong, [_vp]),
    "iono_forwardprojector_factored": (ctypes.c_int, [_vp]),
    "iono_forwardprojector_n_voxels": (ctypes.c_longlong, [_vp]),
    "iono_forwardprojector_voxels": (ctypes.c_int, [_vp, _vp, _vp]),
    "iono_forwardprojector_adjoint_f64": (ctypes.c_int, [_vp, _vp, _vp, _vp]),
    "iono_forwardprojector_finish_gradient_f64": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_double, _vp, _vp]),
    "iono_forwardprojector_finish_compact_f64": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "iono_forwardprojector_destroy": (_i, [_vp]),
}

# kernels launched per C call (for bench.py's ``gpu_launches``; memsets are not counted)
KERNEL_LAUNCHES = {
    "iono_ne_from_m_f64": 1, "iono_mul_f64": 1, "iono_cast_rays_straight_f64": 1,
    "iono_cast_rays_frames_f64": 1, "iono_cast_rays_arclength_f64": 1, "iono_ne_to_refractive_index_f64": 1, "iono_optical_path_f64": 1,
    "iono_tci_interp_f64": 1, "iono_tricubic_derivs_f64": 7, "iono_tricubic_interp_f64": 1, "iono_bent_rays_f64": 1, "iono_tec_forward_f64": 1, "iono_dtec_f64": 1, "iono_adjoint_coef_f64": 1,
    "iono_tec_adjoint_f64": 1, "iono_misfit_f64": 2, "iono_convolve3d_nearest_f64": 1,
    "iono_phase_integrals_f64": 1, "iono_simps_rows_f64": 1, "iono_phase_assemble_f64": 1, "iono_chord_adjoint_f64": 1,
    "iono_gaussian_adjoint_f64": 1,
    "iono_backprojector_apply_f64": 4, "iono_backprojector_apply_chunks_f64": 3,
    "iono_backprojector_apply_permuted_f64": 3, "iono_backprojector_apply_gradient_f64": 3,
    "iono_backprojector_apply_compact_f64": 3, "iono_backprojector_ne_rows_f64": 1, "iono_forwardprojector_quads_from_m_f64": 1,
    "iono_forwardprojector_create": 5, "iono_forwardprojector_apply_f64": 2, "iono_forwardprojector_apply_quads_f64": 1,
    "iono_forwardprojector_adjoint_f64": 1, "iono_forwardprojector_finish_gradient_f64": 1,
    "iono_forwardprojector_finish_compact_f64": 1,
    "iono_peer_reduce_expand_f64": 1, "iono_multi_dot_f64": 2, "iono_multi_dot3_f64": 4, "iono_lincomb_f64": 1, "iono_gather_f64": 1, "iono_scatter_set_f64": 1,
    "iono_scatter_axpy_f64": 1,
    "iono_quads_from_ne_f64": 1, "iono_ne_quads_from_m_f64": 1, "iono_tec_forward_quads_f64": 1, "iono_residual_f64": 1,
}
launch_count = 0

_lib = None


class IonoError(RuntimeError):
    pass


def load():
    """Load the shared library (once) and declare every prototype."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise IonoError(
            "libionob200.so is not built (%s). Run `python -m ionotomo_b200.build` or "
            "`python -c 'import __graft_entry__ as g; g.build()'`; there is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status, what):
    if status != IONO_OK:
        msg = load().iono_last_error().decode("utf-8", "replace")
        raise IonoError("%s failed with status %d: %s" % (what, status, msg))


def call(name, *args):
    """Invoke a status-returning entry point, raise on error, count its kernel launches."""
    global launch_count
    status = getattr(load(), name)(*args)
    check(status, name)
    launch_count += KERNEL_LAUNCHES.get(name, 0)


def require_cuda():
    if not torch.cuda.is_available():
        raise IonoError("ionotomo_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a contiguous fp64 CUDA tensor."""
    assert t.is_cuda and t.dtype == torch.float64 and t.is_contiguous(), \
        "expected a contiguous float64 CUDA tensor"
    return ctypes.c_void_p(t.data_ptr())


def to_device(a, device=None):
    """numpy / torch (any device) -> contiguous float64 CUDA tensor (no copy if already so)."""
    require_cuda()
    if isinstance(a, torch.Tensor):
        t = a
    else:
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64))
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return t.to(device=device, dtype=torch.float64, non_blocking=True).contiguous()


def host_f64(a):
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(a, dtype=np.float64)
