"""Single-process multi-GPU mode of the C-ABI (include/parament.h: Parament_setDevices / Parament_setDeviceList /
$PARAMENT_NUM_GPUS; SURVEY.md 8e "one host process drives all devices").

The sharing logic -- time slices per device, one host thread per device, peer copy of the partial propagators, ordered
combine on the first device; pulse ranges for ensembles -- is exercised on ONE GPU by listing device 0 several times;
the tests at the end use every visible device when the box has more than one.
"""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from oracle.equiprop_oracle import equiprop_oracle, rel_frobenius
from parament_b200 import constants as K
from workloads import make_workload

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def pb():
    import parament_b200
    return parament_b200


def _single(pb, w, carr=None):
    with pb.Parament(w.precision) as ctx:
        ctx.set_hamiltonian(w.H0, *w.H1, use_magnus=w.use_magnus, quadrature_mode=w.quadrature)
        return ctx.equiprop(w.dt, *(w.carr if carr is None else carr))


@pytest.mark.parametrize("name,pts,devices,tol", [
    ("C2", 200001, [0, 0, 0], 2e-6),        # register-resident family, SIMPSON: slices share their end points
    ("C2", 131073, [0, 0], 2e-6),
    ("C1", 400000, [0, 0, 0], 1e-12),       # dim 2 complex128, MIDPOINT: one point of overlap per slice
    ("C3", 1500, [0, 0, 0, 0], 1e-12),      # shared-memory-resident family
    ("C4", 200, [0, 0, 0], 1e-12),          # batched GEMM pipeline
])
def test_time_axis_is_shared_between_listed_devices(pb, name, pts, devices, tol):
    w = make_workload(name, pts=pts)
    ref = _single(pb, w)
    with pb.Parament(w.precision) as ctx:
        ctx.set_devices(devices)
        assert ctx.stat(K.STAT_DEVICES_CONFIGURED) == len(devices)
        ctx.set_hamiltonian(w.H0, *w.H1, use_magnus=w.use_magnus, quadrature_mode=w.quadrature)
        U = ctx.equiprop(w.dt, *w.carr)
        assert ctx.stat(K.STAT_DEVICES_USED) == len(devices)
        assert ctx.stat(K.STAT_STEPS) == w.steps
        assert ctx.stat(K.STAT_LAUNCHES) >= len(devices) + 1       # one fused chain launch per device + the one-launch combine
        assert ctx.stat(K.STAT_DEVICE_MS) > 0
        U2 = ctx.equiprop(w.dt, *w.carr)            # scratch is reused; same slices, same result
    assert np.array_equal(U, U2)
    assert rel_frobenius(U, ref) < tol


def test_shared_call_matches_the_oracle(pb):
    w = make_workload("C2", pts=140001)
    with pb.Parament(w.precision) as ctx:
        ctx.set_hamiltonian(w.H0, *w.H1, use_magnus=w.use_magnus, quadrature_mode=w.quadrature)
        ctx.set_devices([0, 0, 0, 0])              # helpers created AFTER setHamiltonian get the Hamiltonian replayed
        U = ctx.equiprop(w.dt, *w.carr)
        assert ctx.stat(K.STAT_DEVICES_USED) == 4
    Uo = equiprop_oracle(w.H0, w.H1, w.carr, w.dt, w.quadrature, w.use_magnus, w.precision, workers=8)
    assert rel_frobenius(U, Uo) < 1e-5


def test_magnus_and_manual_degree_reach_the_helpers(pb):
    w = make_workload("C2", pts=100001)
    with pb.Parament("fp64") as one, pb.Parament("fp64") as many:
        many.set_devices([0, 0, 0])
        for ctx in (one, many):
            ctx.set_hamiltonian(w.H0, *w.H1, use_magnus=True, quadrature_mode="simpson")
        a, b = one.equiprop(w.dt, *w.carr), many.equiprop(w.dt, *w.carr)
        assert many.stat(K.STAT_DEVICES_USED) == 3 and rel_frobenius(b, a) < 1e-12
        for ctx in (one, many):
            ctx.set_iteration_cycles(4)             # a visibly truncated series: the helpers must truncate alike
        a4, b4 = one.equiprop(w.dt, *w.carr), many.equiprop(w.dt, *w.carr)
        assert many.stat(K.STAT_DEGREE_USED) == 4
        assert rel_frobenius(b4, a4) < 1e-12 and rel_frobenius(a4, a) > 1e-9
        for ctx in (one, many):
            ctx.set_iteration_cycles(None)
        assert rel_frobenius(many.equiprop(w.dt, *w.carr), a) < 1e-12


def test_ensemble_is_shared_by_pulse_ranges(pb):
    w = make_workload("C5", batch=700)
    with pb.Parament(w.precision) as ctx:
        ctx.set_hamiltonian(w.H0, *w.H1, quadrature_mode=w.quadrature)
        ref = ctx.equiprop_batch(w.dt, w.carr)
        ctx.set_devices([0, 0, 0])
        U = ctx.equiprop_batch(w.dt, w.carr)
        assert ctx.stat(K.STAT_DEVICES_USED) == 3
        assert ctx.stat(K.STAT_D2H) == 700 * 64 * 8
    assert U.shape == (700, 8, 8)
    assert np.abs(U - ref).max() < 2e-6


def test_small_calls_stay_on_one_device(pb):
    w = make_workload("C2", pts=2001)
    with pb.Parament(w.precision) as ctx:
        ctx.set_devices([0, 0])
        ctx.set_hamiltonian(w.H0, *w.H1, quadrature_mode=w.quadrature)
        U = ctx.equiprop(w.dt, *w.carr)
        assert ctx.stat(K.STAT_DEVICES_USED) == 1 and ctx.stat(K.STAT_DEVICES_CONFIGURED) == 2
        # zero effective steps -> identity, as on one device
        assert np.array_equal(ctx.equiprop(w.dt, np.zeros(1), np.zeros(1)), np.eye(16, dtype=np.complex64))
    assert rel_frobenius(U, _single(pb, w)) < 1e-6


def test_errors_of_a_shared_call(pb):
    w = make_workload("C2", pts=200001)
    with pb.Parament(w.precision) as ctx:
        ctx.set_devices([0, 0])
        with pytest.raises(RuntimeError, match="No hamiltonian set"):
            ctx._check_error(ctx._fn("Parament_equiprop")(ctx._handle, np.zeros(8, np.complex64), 0.1, 4, 2, np.zeros(256, np.complex64)))
        ctx.set_hamiltonian(w.H0, *w.H1, quadrature_mode=w.quadrature)
        with pytest.raises(RuntimeError, match="Timestep too large"):       # Hnorm * h far beyond the table: every device refuses
            ctx.equiprop(100.0, *w.carr)
        assert ctx._lib.Parament_getLastError(ctx._handle) == K.PARAMENT_STATUS_SELECT_SMALLER_DT
        U = ctx.equiprop(w.dt, *w.carr)                               # the context stays usable
        assert ctx.stat(K.STAT_DEVICES_USED) == 2
        with pytest.raises(ValueError):
            ctx.set_devices([0, 99])
        with pytest.raises(ValueError, match="Invalid quadrature selection"):
            ctx.set_hamiltonian(w.H0, *w.H1, use_magnus=True, quadrature_mode="midpoint")
        with pytest.raises(RuntimeError, match="No hamiltonian set"):
            ctx.equiprop(w.dt, *w.carr)
        ctx.set_devices(1)
        assert ctx.stat(K.STAT_DEVICES_CONFIGURED) == 1
    assert rel_frobenius(U, _single(pb, w)) < 2e-6


def test_all_visible_devices(pb, gpu_count):
    """Every GPU of the box (the real thing when there is more than one; on a one-GPU box this is the plain path)."""
    w = make_workload("C2", pts=400001)
    ref = _single(pb, w)
    with pb.Parament(w.precision) as ctx:
        ctx.set_devices(0)
        assert ctx.stat(K.STAT_DEVICES_CONFIGURED) == gpu_count
        ctx.set_hamiltonian(w.H0, *w.H1, quadrature_mode=w.quadrature)
        U = ctx.equiprop(w.dt, *w.carr)
        assert ctx.stat(K.STAT_DEVICES_USED) == min(gpu_count, 8)
    assert rel_frobenius(U, ref) < 2e-6
    if gpu_count > 1:
        w3 = make_workload("C3", pts=4000)
        ref3 = _single(pb, w3)
        with pb.Parament("fp64") as ctx:
            ctx.set_devices(list(range(gpu_count - 1, -1, -1)))     # the context moves to the last device
            ctx.set_hamiltonian(w3.H0, *w3.H1, quadrature_mode=w3.quadrature)
            U3 = ctx.equiprop(w3.dt, *w3.carr)
            assert ctx.stat(K.STAT_DEVICES_USED) == gpu_count
        assert rel_frobenius(U3, ref3) < 1e-12


def test_unchanged_reference_wrapper_with_num_gpus_env(tmp_path, gpu_count):
    """$PARAMENT_NUM_GPUS: the reference's own wrapper, unmodified, on all devices (0 = all visible)."""
    wrapper = os.path.join(ROOT, "oracle", "_ref", "pyparament")
    if not os.path.isdir(os.path.join(wrapper, "parament")):
        pytest.skip("reference wrapper not staged (oracle/build_ref.sh)")
    code = textwrap.dedent(f"""
        import sys, numpy as np
        np.float = float
        sys.path.insert(0, {wrapper!r}); sys.path.insert(0, {ROOT!r})
        import parament
        from workloads import make_workload
        w = make_workload("C3", pts=3000)
        ctx = parament.Parament(precision="fp64")
        ctx.set_hamiltonian(w.H0, *w.H1, use_magnus=False, quadrature_mode="none")
        U = ctx.equiprop(w.dt, *w.carr)
        used = ctx._lib.Parament_lastStat
        import ctypes
        used.restype = ctypes.c_double
        print("DEVICES", int(used(ctx._handle, 11)), int(used(ctx._handle, 12)))
        np.save({str(tmp_path / 'u.npy')!r}, np.asarray(U))
        ctx.destroy()
    """)
    env = dict(os.environ, PARAMENT_LIB_DIR=os.path.join(ROOT, "parament_b200", "lib"), PARAMENT_NUM_GPUS="0")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert f"DEVICES {gpu_count} {gpu_count}" in r.stdout
    w = make_workload("C3", pts=3000)
    import parament_b200 as pb
    assert rel_frobenius(np.load(tmp_path / "u.npy"), _single(pb, w)) < 1e-12
