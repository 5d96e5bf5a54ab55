"""Host-side mirror of the reference's Python interface for the equiprop path.

Same class name, method names, argument meaning and exceptions as
/root/reference/src/python/pyparament/parament/parament.py:48-314 (`Parament.set_hamiltonian`,
`.equiprop`, `.destroy`, context manager, error-code -> exception mapping at :287-308), written
against include/parament.h.  The unchanged reference wrapper also works on this library
(INTEGRATION.md); this mirror exists so that the package is usable without the reference tree and to
expose the additive entry points: ensembles, time slices, device-resident operands, statistics.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import constants as K
from ._lib import lib

_EXC = {
    K.PARAMENT_STATUS_HOST_ALLOC_FAILED: MemoryError,
    K.PARAMENT_STATUS_DEVICE_ALLOC_FAILED: MemoryError,
    K.PARAMENT_STATUS_CUBLAS_INIT_FAILED: RuntimeError,
    K.PARAMENT_STATUS_INVALID_VALUE: ValueError,
    K.PARAMENT_STATUS_CUBLAS_FAILED: RuntimeError,
    K.PARAMENT_STATUS_SELECT_SMALLER_DT: RuntimeError,
    K.PARAMENT_STATUS_NO_HAMILTONIAN: RuntimeError,
    K.PARAMENT_STATUS_INVALID_QUADRATURE_SELECTION: ValueError,
    K.PARAMENT_FAIL: RuntimeError,
}


class Parament:
    """A Parament context (reference parament.py:48-103).  precision: 'fp32' (complex64) | 'fp64' (complex128)."""

    def __init__(self, precision="fp32", device=None):
        if precision not in ("fp32", "fp64"):
            raise ValueError("precision must be either 'fp32' or 'fp64'")
        self._use_doubles = precision == "fp64"
        self._sfx = "_fp64" if self._use_doubles else ""
        self._ctype = np.complex128 if self._use_doubles else np.complex64
        self._lib = lib
        self._handle = ctypes.c_void_p()
        self._check_error(self._fn("Parament_create")(ctypes.byref(self._handle)))
        self.dim = -1
        self.amps = -1
        self._quadrature, self._magnus = "none", False
        if device is not None:
            self._check_error(lib.Parament_setDevice(self._handle, int(device)))

    def _fn(self, name):
        return getattr(self._lib, name + self._sfx)

    def _alive(self):
        if self._handle is None:
            raise RuntimeError("Attempting to use a context that has been destroyed")

    # ---- lifecycle -------------------------------------------------------------------------------
    def destroy(self):
        self._alive()
        self._check_error(self._fn("Parament_destroy")(self._handle))
        self._handle = None

    def __del__(self):
        if getattr(self, "_handle", None) is not None and self._handle.value is not None:
            try:
                self.destroy()
            except Exception:
                pass

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc_val, exc_tb):
        self.destroy()

    # ---- reference interface -----------------------------------------------------------------------
    def set_hamiltonian(self, H0, *H1, use_magnus=False, quadrature_mode="none"):
        """Load drift and control Hamiltonians (reference parament.py:125-213)."""
        self._alive()
        if quadrature_mode not in K.QUADRATURE:
            raise ValueError("unknown quadrature mode selected")
        if len(H1) == 0:
            raise ValueError("provide at least 1 control amplitude")
        H0 = np.atleast_2d(np.asarray(H0))
        H1 = np.atleast_2d(np.asarray(H1))
        if H1.ndim == 4 and H1.shape[0] == 1:   # a single 3-D stack passed with *
            H1 = H1[0]
        amps = H1.shape[0] if H1.ndim > 2 else 1
        dim = H0.shape[0]
        self.dim, self.amps = dim, amps
        self._quadrature, self._magnus = quadrature_mode, bool(use_magnus)
        self._check_error(self._fn("Parament_setHamiltonian")(
            self._handle,
            np.ascontiguousarray(np.ravel(H0, order="C").astype(self._ctype)),
            np.ascontiguousarray(np.ravel(H1, order="C").astype(self._ctype)),
            dim, amps, bool(use_magnus), K.QUADRATURE[quadrature_mode]))

    def _carr(self, carr):
        amps = len(carr)
        if amps > self.amps:
            raise ValueError(f"Got {amps} amplitude arrays, but there are only {self.amps} Hamiltonians.")
        pts = np.shape(carr[0])[0]
        if any(np.shape(c) != (pts,) for c in carr):
            raise ValueError("All amplitude arrays must have the same length.")
        return np.ascontiguousarray(np.asarray(carr).astype(self._ctype).ravel(order="C")), pts, amps

    def equiprop(self, dt, *carr):
        """Propagator U = U_{N-1}...U_0 of the pulse `carr` (reference parament.py:216-276)."""
        self._alive()
        if self.amps < 0:
            raise RuntimeError("No hamiltonian set")
        flat, pts, amps = self._carr(carr)
        out = np.zeros(self.dim ** 2, dtype=self._ctype)
        self._check_error(self._fn("Parament_equiprop")(self._handle, flat, float(dt), pts, amps, out))
        return out.reshape(self.dim, self.dim)

    # ---- additive --------------------------------------------------------------------------------
    def equiprop_batch(self, dt, carr):
        """carr: (batch, amps, pts) -> (batch, dim, dim).  Independent pulses, one call."""
        self._alive()
        if self.amps < 0:
            raise RuntimeError("No hamiltonian set")
        carr = np.asarray(carr)
        if carr.ndim != 3:
            raise ValueError("carr must have shape (batch, amps, pts)")
        batch, amps, pts = carr.shape
        if amps > self.amps:
            raise ValueError(f"Got {amps} amplitude arrays, but there are only {self.amps} Hamiltonians.")
        flat = np.ascontiguousarray(carr.astype(self._ctype).ravel(order="C"))
        out = np.zeros(batch * self.dim ** 2, dtype=self._ctype)
        self._check_error(self._fn("Parament_equipropBatch")(self._handle, flat, float(dt), pts, amps, batch, out))
        return out.reshape(batch, self.dim, self.dim)

    def equiprop_slice(self, dt, carr, step_lo, step_hi):
        """Partial propagator of effective steps [step_lo, step_hi); carr: (amps, pts)."""
        self._alive()
        carr = np.atleast_2d(np.asarray(carr))
        amps, pts = carr.shape
        flat = np.ascontiguousarray(carr.astype(self._ctype).ravel(order="C"))
        out = np.zeros(self.dim ** 2, dtype=self._ctype)
        self._check_error(self._fn("Parament_equipropSlice")(self._handle, flat, float(dt), pts, amps,
                                                             int(step_lo), int(step_hi), out))
        return out.reshape(self.dim, self.dim)

    def steps_of(self, pts):
        """Effective steps of a pulse with `pts` points under the quadrature set last (parament.cpp:820-831)."""
        if self._quadrature == "simpson" or self._magnus:
            return max((pts - 1) // 2, 0)
        return max(pts - 1, 0) if self._quadrature == "midpoint" else pts

    def combine(self, parts):
        """parts[count-1] @ ... @ parts[0] on the device; parts: (count, dim, dim), earliest slice first."""
        self._alive()
        parts = np.ascontiguousarray(np.asarray(parts).astype(self._ctype))
        out = np.zeros(self.dim ** 2, dtype=self._ctype)
        self._check_error(self._fn("Parament_combine")(self._handle, parts.ravel(), parts.shape[0], out))
        return out.reshape(self.dim, self.dim)

    def combine_device(self, parts_ptr, count, out_ptr, stream=None):
        """Ordered product of `count` device-resident partial propagators (raw device pointers, context precision)."""
        self._alive()
        self._check_error(lib.Parament_combineDevice(self._handle, ctypes.c_void_p(parts_ptr), int(count), ctypes.c_void_p(out_ptr),
                                                     ctypes.c_void_p(stream) if stream else None))

    def equiprop_device(self, dt, carr_ptr, pts, amps, out_ptr, batch=1, stream=None):
        """Device-resident operands (raw device pointers, e.g. torch.Tensor.data_ptr())."""
        self._alive()
        self._check_error(self._fn("Parament_equipropDevice")(
            self._handle, ctypes.c_void_p(carr_ptr), float(dt), pts, amps, batch, ctypes.c_void_p(out_ptr),
            ctypes.c_void_p(stream) if stream else None))

    def set_iteration_cycles(self, cycles=None):
        """cycles=None restores the table-driven choice (reference parament.cpp:772-786)."""
        self._alive()
        if cycles is None:
            self._check_error(self._fn("Parament_automaticIterationCycles")(self._handle))
        else:
            self._check_error(self._fn("Parament_setIterationCyclesManually")(self._handle, int(cycles)))

    def set_devices(self, devices=0):
        """Single-process multi-GPU (include/parament.h, Parament_setDevices / Parament_setDeviceList): an int n uses n
        devices starting at the context's own (0 = all visible, 1 = back to one device); a sequence names the devices,
        the first one hosting the context.  Later equiprop / equiprop_batch calls with host arrays are shared."""
        self._alive()
        if isinstance(devices, (int, np.integer)):
            self._check_error(lib.Parament_setDevices(self._handle, int(devices)))
        else:
            arr = (ctypes.c_int * len(devices))(*[int(d) for d in devices])
            self._check_error(lib.Parament_setDeviceList(self._handle, arr, len(devices)))

    def stat(self, key):
        self._alive()
        return lib.Parament_lastStat(self._handle, int(key))

    def stats(self):
        names = ["device_ms", "launches", "degree_used", "degree_reference", "steps", "family", "h2d_bytes", "d2h_bytes", "hnorm",
                 "horner", "products_per_step"]
        return {n: self.stat(i) for i, n in enumerate(names)}

    # ---- errors ----------------------------------------------------------------------------------
    def _get_error_message(self, code=None):
        if code is None:
            code = lib.Parament_getLastError(self._handle)
        return lib.Parament_errorMessage(code).decode()

    def _check_error(self, error_code):
        if error_code == K.PARAMENT_STATUS_SUCCESS:
            return
        if error_code not in _EXC:
            raise AssertionError("Unknown error code ")
        raise _EXC[error_code](f"Error code {error_code}: {lib.Parament_errorMessage(error_code).decode()}")


def expm(m):
    """exp(m) through the propagator (reference debug_functions.py:22-31): H0 = i m, one step, dt = 1."""
    with Parament() as ctx:
        ctx.set_hamiltonian(1j * np.asarray(m), np.asarray(m), use_magnus=False, quadrature_mode="none")
        return ctx.equiprop(1.0, np.zeros(1))


def device_info():
    lib.device_info()
