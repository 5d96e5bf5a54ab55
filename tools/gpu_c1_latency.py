"""Latency breakdown of a C1-sized call (dim 2, one control, complex128, MIDPOINT) on the GPU box.

Prints per-call host time of the host-pointer entry (pageable and page-locked buffers), of the device entry with a
synchronize per call, the back-to-back device-entry rate, the kernel's own event time, and the floor set by the CUDA
runtime for the same three operations (H2D of the amplitude bytes, an empty-ish kernel, D2H of 64 bytes + synchronize).
"""
import ctypes
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from parament_b200 import Parament  # noqa: E402
from parament_b200._lib import lib, library_path  # noqa: E402
from parament_b200 import constants as K  # noqa: E402
from workloads import make_workload  # noqa: E402


def timeit(f, reps):
    for _ in range(20):
        f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e6


def main():
    reps = 2000
    for pts in (100, 1000, 10_000, 100_000):
        w = make_workload("C1", pts=pts)
        ctx = Parament(precision="fp64")
        ctx.set_hamiltonian(w.H0, w.H1, use_magnus=w.use_magnus, quadrature_mode=w.quadrature)
        h = ctx._handle
        carr = np.ascontiguousarray(w.carr.astype(np.complex128))
        out = np.empty((2, 2), np.complex128)
        A = carr.shape[0] if carr.ndim == 2 else 1
        cp = carr.ctypes.data_as(ctypes.c_void_p)
        op = out.ctypes.data_as(ctypes.c_void_p)
        rawlib = ctypes.CDLL(str(library_path()))   # second handle of the same library: plain pointer argtypes, no ndarray checks
        raw = rawlib.Parament_equiprop_fp64
        raw.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double, ctypes.c_uint, ctypes.c_uint, ctypes.c_void_p]
        t_page = timeit(lambda: raw(h, cp, w.dt, pts, A, op), reps)
        ms_kernel = lib.Parament_lastStat(h, K.STAT_DEVICE_MS) * 1e3
        launches = lib.Parament_lastStat(h, K.STAT_LAUNCHES)
        # page-locked
        tc = torch.from_numpy(carr.copy()).pin_memory()
        to = torch.empty(4, dtype=torch.complex128).pin_memory()
        t_pin = timeit(lambda: raw(h, tc.data_ptr(), w.dt, pts, A, to.data_ptr()), reps)
        # device entry
        dc = tc.cuda()
        do = torch.empty(4, dtype=torch.complex128, device="cuda")
        dfn = rawlib.Parament_equipropDevice_fp64
        dfn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double, ctypes.c_uint, ctypes.c_uint, ctypes.c_uint, ctypes.c_void_p, ctypes.c_void_p]
        st = torch.cuda.current_stream().cuda_stream

        def dev_sync():
            dfn(h, dc.data_ptr(), w.dt, pts, A, 1, do.data_ptr(), st)
            torch.cuda.current_stream().synchronize()

        t_dev_sync = timeit(dev_sync, reps)
        t_dev_async = timeit(lambda: dfn(h, dc.data_ptr(), w.dt, pts, A, 1, do.data_ptr(), st), reps)
        # runtime floor
        hc = torch.from_numpy(carr.copy())
        ho = torch.empty(4, dtype=torch.complex128)

        def floor_pageable():
            dc.copy_(hc, non_blocking=True)
            do.add_(1.0)
            ho.copy_(do)

        def floor_pinned():
            dc.copy_(tc, non_blocking=True)
            do.add_(1.0)
            to.copy_(do, non_blocking=True)
            torch.cuda.current_stream().synchronize()

        t_floor_page = timeit(floor_pageable, reps)
        t_floor_pin = timeit(floor_pinned, reps)
        print(f"pts {pts:7d} in {carr.nbytes/1024:7.1f} KB | host pageable {t_page:6.1f} us  pinned {t_pin:6.1f} us | device entry +sync {t_dev_sync:6.1f} us  "
              f"back-to-back {t_dev_async:6.1f} us  kernel events {ms_kernel:6.1f} us launches {launches:.0f} | torch floor pageable {t_floor_page:6.1f} pinned {t_floor_pin:6.1f} us",
              flush=True)
        ctx.destroy()


if __name__ == "__main__":
    main()
