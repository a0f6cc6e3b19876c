"""NVLink peer memory between the ranks of one node (one process per GPU).

Buffers are allocated by the library (``iono_peer_alloc``: cudaMalloc + CUDA IPC handle), the 64-byte
handles are exchanged through ``torch.distributed`` (plumbing), and every rank maps the others'
buffers (``iono_peer_open``).  ``PeerReducer`` drives ``iono_peer_reduce_expand_f64``: the cross-rank
sum of the compact adjoint accumulators, the chain-rule scaling and the expansion to the grid as one
kernel per rank -- the replacement of the reference's ``da.sum`` over dask workers
(``inversion/gradient.py:52-54``).
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib


class _ForeignCuda(object):
    """``__cuda_array_interface__`` view of library-owned device memory (for torch.as_tensor)."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


class PeerBuffer(object):
    """``nbytes`` of zeroed device memory on every rank, each rank's block mapped into all others."""

    def __init__(self, nbytes, group=None):
        lib = _lib.load()
        _lib.require_cuda()
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.nbytes = int(nbytes)
        p = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        _lib.call("iono_peer_alloc", self.nbytes, ctypes.byref(p), handle)
        self.local = p.value
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle.raw), group=group)
        self.ptrs = []
        err = None
        for r, h in enumerate(handles):
            if r == self.rank:
                self.ptrs.append(self.local)
            else:
                q = ctypes.c_void_p()
                try:
                    _lib.call("iono_peer_open", ctypes.c_char_p(h), ctypes.byref(q))
                except _lib.IonoError as exc:      # keep going: every rank must reach the same collectives
                    err = exc
                self.ptrs.append(q.value)
        if err is not None:
            raise err
        self.table = (ctypes.c_void_p * self.world)(*self.ptrs)

    def tensor(self, dtype=torch.float64, offset_bytes=0, n=None):
        """Torch view of the LOCAL block."""
        size = torch.empty(0, dtype=dtype).element_size()
        typestr = {torch.float64: "<f8", torch.int64: "<i8", torch.uint8: "|u1"}[dtype]
        if n is None:
            n = (self.nbytes - offset_bytes) // size
        return torch.as_tensor(_ForeignCuda(self.local + offset_bytes, n, typestr), device="cuda")

    def close(self):
        lib = _lib.load()
        if getattr(self, "ptrs", None):
            torch.cuda.synchronize()
            if dist.is_initialized():
                dist.barrier(group=self.group)          # nobody still reads a block that is about to go
            for r, q in enumerate(self.ptrs):
                if r != self.rank:
                    lib.iono_peer_close(ctypes.c_void_p(q))
            if dist.is_initialized():
                dist.barrier(group=self.group)
            lib.iono_peer_free(ctypes.c_void_p(self.local))
            self.ptrs = None


class PeerReducer(object):
    """Compact accumulators, result vectors and flags of all ranks + the fused reduce/expand kernel."""

    def __init__(self, length, group=None):
        lib = _lib.load()
        self.L = int(length) + (int(length) & 1)
        self.acc = PeerBuffer(self.L * 8, group)
        self.res = PeerBuffer(self.L * 8, group)
        self.flags = PeerBuffer(int(lib.iono_peer_flag_bytes()), group)
        self.rank, self.world = self.acc.rank, self.acc.world
        self.acc_t = self.acc.tensor()
        self.res_t = self.res.tensor()
        torch.cuda.synchronize()
        dist.barrier(group=group)

    def reduce_expand(self, union_voxels, n_union, m, k, grad, misfit_out=None):
        _lib.call("iono_peer_reduce_expand_f64", self.acc.table, self.res.table, self.flags.table, self.world,
                  self.rank, self.L, ctypes.c_void_p(union_voxels.data_ptr()), int(n_union), _lib.ptr(m), float(k),
                  _lib.ptr(grad), _lib.ptr(misfit_out) if misfit_out is not None else None, _lib.stream_ptr())

    def phase_times_us(self):
        """Phases of the LAST reduce/expand call on this rank (CTA 0's time stamps): wait for all accumulators,
        reduce, push and expand my slice, wait for the first foreign slice, expand the others as they arrive --
        microseconds."""
        torch.cuda.synchronize()
        t = self.flags.tensor(torch.int64)[2 * 16 + 2:2 * 16 + 7].cpu().numpy().astype("float64")
        names = ("wait_accumulators", "reduce_push_expand_own_slice", "wait_first_foreign_slice", "expand_arriving_slices")
        return {n: float(t[i + 1] - t[i]) / 1e3 for i, n in enumerate(names)}

    def close(self):
        for b in (self.acc, self.res, self.flags):
            b.close()


class LocalExpander(object):
    """The reduce/expand kernel on ONE rank's vectors (N = 1: copy, scale, expand): the tail of the step when the
    cross-rank sum of ``acc_t`` is done by ``torch.distributed.all_reduce`` instead of the peer kernel."""

    def __init__(self, length, device):
        lib = _lib.load()
        self.L = int(length) + (int(length) & 1)
        self.acc_t = torch.zeros(self.L, dtype=torch.float64, device=device)
        self.res_t = torch.zeros(self.L, dtype=torch.float64, device=device)
        self.flags_t = torch.zeros(int(lib.iono_peer_flag_bytes()) // 8, dtype=torch.int64, device=device)
        self.acc_tab = (ctypes.c_void_p * 1)(self.acc_t.data_ptr())
        self.res_tab = (ctypes.c_void_p * 1)(self.res_t.data_ptr())
        self.flags_tab = (ctypes.c_void_p * 1)(self.flags_t.data_ptr())

    def reduce_expand(self, union_voxels, n_union, m, k, grad, misfit_out=None):
        _lib.call("iono_peer_reduce_expand_f64", self.acc_tab, self.res_tab, self.flags_tab, 1, 0, self.L,
                  ctypes.c_void_p(union_voxels.data_ptr()), int(n_union), _lib.ptr(m), float(k), _lib.ptr(grad),
                  _lib.ptr(misfit_out) if misfit_out is not None else None, _lib.stream_ptr())
