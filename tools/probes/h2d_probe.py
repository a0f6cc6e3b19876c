#!/usr/bin/env python
"""Probe: host->device bandwidth for the shapes misfit_and_gradient uses (pinned memory)."""
import ctypes, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from ionotomo_b200 import _lib
Na, Nt, Nd, Ns = 62, 100, 200, 128
row = Nd * 4 * Ns * 8
h = torch.empty((Na, Nt, Nd, 4, Ns), dtype=torch.float64, pin_memory=True)
h.fill_(1.0)
d = torch.empty_like(h, device="cuda")
def t(fn, n=3):
    fn(); torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.time() - t0) / n
gb = h.numel() * 8 / 1e9
print("contiguous copy_      : %.1f GB/s" % (gb / t(lambda: d.copy_(h, non_blocking=True))))
for tb in (1, 2, 5, 10, 25):
    buf = torch.empty((Na, tb, Nd, 4, Ns), dtype=torch.float64, device="cuda")
    def blocks():
        for t0 in range(0, Nt, tb):
            _lib.call("iono_copy2d_h2d", ctypes.c_void_p(buf.data_ptr()), tb * row, ctypes.c_void_p(h.data_ptr() + t0 * row),
                      Nt * row, tb * row, Na, _lib.stream_ptr())
    print("2-D blocks of %2d times: %.1f GB/s" % (tb, gb / t(blocks)))
# per-antenna contiguous slabs (1-D copies)
buf = torch.empty((Nt, Nd, 4, Ns), dtype=torch.float64, device="cuda")
def slabs():
    for a in range(Na):
        buf.copy_(h[a], non_blocking=True)
print("per-antenna 1-D copies : %.1f GB/s" % (gb / t(slabs)))
