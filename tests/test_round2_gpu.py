"""GPU tests added in round 2 (run with -m gpu on the B200 box), all through the C-ABI:
  * the device-resident entry points (Parament_equipropDevice, Parament_combineDevice, Parament_equipropSliceToDevice)
    against the float64 oracle, not against the host entry of the same library;
  * the spectral series norm of dim > 16 (fewer matrix products per step): parity, reported statistics, clamp to the
    reference's bound, A/B against PARAMENT_NORM=reference;
  * backward propagation (dt < 0), handles used after destroy, the caller's current device left alone.
"""
import ctypes
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from oracle.equiprop_oracle import equiprop_oracle, rel_frobenius
from workloads import make_workload, rand_herm

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = {"fp32": 1e-5, "fp64": 1e-12}


@pytest.fixture(scope="module")
def pb():
    import parament_b200
    return parament_b200


@pytest.mark.parametrize("name,pts", [("C2", 50001), ("C1", 4001), ("C3", 300), ("C4", 40), ("C5", 1000)])
def test_device_entry_vs_oracle(pb, name, pts):
    torch = pytest.importorskip("torch")
    w = make_workload(name, pts=pts, batch=3 if name == "C5" else None)
    tdt = torch.complex64 if w.precision == "fp32" else torch.complex128
    carr = torch.from_numpy(np.ascontiguousarray(w.carr)).cuda()
    out = torch.zeros(w.batch, w.dim, w.dim, dtype=tdt, device="cuda")
    with pb.Parament(w.precision) as ctx:
        ctx.set_hamiltonian(w.H0, *w.H1, use_magnus=w.use_magnus, quadrature_mode=w.quadrature)
        ctx.equiprop_device(w.dt, carr.data_ptr(), w.pts, w.amps, out.data_ptr(), batch=w.batch)
        torch.cuda.synchronize()
        assert ctx.stat(1) >= 1
    got = out.cpu().numpy()
    pulses = w.carr.reshape(w.batch, w.amps, w.pts)
    for b in range(w.batch):
        Uo = equiprop_oracle(w.H0, w.H1, pulses[b], w.dt, w.quadrature, w.use_magnus, w.precision)
        assert rel_frobenius(got[b], Uo) < TOL[w.precision], (name, b)


@pytest.mark.parametrize("name,pts,count", [("C2", 8001, 8), ("C2", 8001, 3), ("C3", 96, 8), ("C4", 16, 4), ("C1", 801, 5)])
def test_combine_device_vs_numpy(pb, name, pts, count):
    """Parament_combineDevice on device-resident partials = the ordered product parts[count-1] ... parts[0]."""
    torch = pytest.importorskip("torch")
    w = make_workload(name, pts=pts)
    rng = np.random.default_rng(7)
    n = w.dim
    parts = []
    for g in range(count):   # unitary-ish partial propagators: oracle propagators of short random pulses
        c = rng.uniform(-1, 1, (w.amps, 9)).astype(w.ctype)
        parts.append(equiprop_oracle(w.H0, w.H1, c, w.dt, "none", False, w.precision).astype(w.ctype))
    parts = np.stack(parts)
    exp = np.eye(n, dtype=np.complex128)
    for g in range(count):
        exp = parts[g].astype(np.complex128) @ exp
    tdt = torch.complex64 if w.precision == "fp32" else torch.complex128
    dparts = torch.from_numpy(parts).cuda()
    dout = torch.zeros(n, n, dtype=tdt, device="cuda")
    with pb.Parament(w.precision) as ctx:
        ctx.set_hamiltonian(w.H0, *w.H1, use_magnus=w.use_magnus, quadrature_mode=w.quadrature)
        ctx.combine_device(dparts.data_ptr(), count, dout.data_ptr())
        torch.cuda.synchronize()
        launches = ctx.stat(1)
        host = ctx.combine(parts)
    assert rel_frobenius(dout.cpu().numpy(), exp) < (2e-7 if w.precision == "fp32" else 1e-14)
    assert rel_frobenius(host, exp) < (2e-7 if w.precision == "fp32" else 1e-14)
    if n <= 16:
        assert launches == 1          # one launch for the register-resident family (round 1: five)


@pytest.mark.parametrize("name,pts", [("C2", 20001), ("C3", 200), ("C1", 2001)])
def test_slice_to_device_vs_oracle(pb, name, pts):
    """Host amplitudes in, partial propagator left on the GPU; slices compose to the whole pulse."""
    torch = pytest.importorskip("torch")
    w = make_workload(name, pts=pts)
    tdt = torch.complex64 if w.precision == "fp32" else torch.complex128
    sfx = "_fp64" if w.precision == "fp64" else ""
    N = w.steps
    bounds = [0, N // 3, N // 3, (2 * N) // 3 + 1, N]       # includes an empty slice
    parts = torch.zeros(len(bounds) - 1, w.dim, w.dim, dtype=tdt, device="cuda")
    flat = np.ascontiguousarray(w.carr.reshape(-1))
    with pb.Parament(w.precision) as ctx:
        ctx.set_hamiltonian(w.H0, *w.H1, use_magnus=w.use_magnus, quadrature_mode=w.quadrature)
        fn = getattr(pb._lib.lib, "Parament_equipropSliceToDevice" + sfx)
        for g in range(len(bounds) - 1):
            assert fn(ctx._handle, flat, float(w.dt), w.pts, w.amps, bounds[g], bounds[g + 1], ctypes.c_void_p(parts[g].data_ptr())) == 0
        out = torch.zeros(w.dim, w.dim, dtype=tdt, device="cuda")
        ctx.combine_device(parts.data_ptr(), len(bounds) - 1, out.data_ptr())
        torch.cuda.synchronize()
    assert np.allclose(parts[1].cpu().numpy(), np.eye(w.dim))               # empty slice -> identity, on the device
    Uo = equiprop_oracle(w.H0, w.H1, w.carr, w.dt, w.quadrature, w.use_magnus, w.precision)
    assert rel_frobenius(out.cpu().numpy(), Uo) < TOL[w.precision]


@pytest.mark.parametrize("dim,prec,quad", [(4, "fp64", "midpoint"), (16, "fp32", "simpson"), (32, "fp64", "none"), (64, "fp64", "simpson"),
                                           (96, "fp64", "none")])
def test_backward_propagation(pb, dim, prec, quad):
    """dt < 0: exp(+i |dt| H) per step, in every kernel family (J_k(-x) = (-1)^k J_k(x); the degree follows |dt|)."""
    rng = np.random.default_rng(11)
    ct = np.complex64 if prec == "fp32" else np.complex128
    H0 = (0.5 * rand_herm(rng, dim)).astype(ct)
    H1 = np.stack([(0.25 * rand_herm(rng, dim)).astype(ct) for _ in range(2)])
    carr = rng.uniform(-1, 1, (2, 41)).astype(ct)
    dt = -0.3
    with pb.Parament(prec) as ctx:
        ctx.set_hamiltonian(H0, *H1, quadrature_mode=quad)
        U = ctx.equiprop(dt, *carr)
        Uf = ctx.equiprop(-dt, *carr)
        assert ctx.stat(2) >= 3
    assert rel_frobenius(U, equiprop_oracle(H0, H1, carr, dt, quad, False, prec)) < TOL[prec]
    assert rel_frobenius(Uf, equiprop_oracle(H0, H1, carr, -dt, quad, False, prec)) < TOL[prec]


def test_spectral_norm_saves_a_product(pb):
    """dim > 16: the series is built for the spectral bound (stat 14) instead of the reference's row-sum bound (stat 8): at the
    BASELINE workloads that lowers the degree from 12 (four products) to 8 (three), at unchanged accuracy."""
    for name, pts in (("C3", 400), ("C4", 24)):
        w = make_workload(name, pts=pts)
        with pb.Parament("fp64") as ctx:
            ctx.set_hamiltonian(w.H0, *w.H1, quadrature_mode="none")
            U = ctx.equiprop(w.dt, *w.carr)
            hn, hs, m_used, m_ref, products = ctx.stat(8), ctx.stat(14), ctx.stat(2), ctx.stat(3), ctx.stat(10)
        assert abs(hn - 1.0) < 1e-12 and 0.05 < hs < 0.5 * hn
        # s(H0) + sum_k max_t |c_k(t)| s(H_k) with the amplitude maxima of THIS pulse, times the 5 % margin
        sig = np.linalg.norm(w.H0, 2) + sum(np.abs(c).max() * np.linalg.norm(h, 2) for c, h in zip(w.carr, w.H1))
        assert 1.04 * sig <= hs <= 1.0501 * sig
        worst = max(np.linalg.norm(w.H0 + np.tensordot(w.carr[:, j], w.H1, axes=1), 2) for j in range(0, pts, max(1, pts // 16)))
        assert worst <= hs                                   # a true bound of the sampled step Hamiltonians
        # dim 256, Hermitian matrices and real amplitudes: 12 of the 32 tiles of Y Y are mirrored, not computed (3 + 20/32 + 1 products)
        assert m_ref == 11 and m_used == 8 and products == (4.0 if name == "C3" else 3.625)
        assert rel_frobenius(U, equiprop_oracle(w.H0, w.H1, w.carr, w.dt, "none", False, "fp64")) < 1e-13


def test_spectral_norm_never_exceeds_reference_bound(pb):
    """Amplitudes above 1 leave the reference's series domain; the spectral bound is then clamped to Hnorm,
    so the degree is never lower than what the reference semantics give, and small amplitudes lower it further."""
    w = make_workload("C3", pts=50)
    with pb.Parament("fp64") as ctx:
        ctx.set_hamiltonian(w.H0, *w.H1, quadrature_mode="none")
        ctx.equiprop(w.dt, *(40.0 * w.carr))
        assert ctx.stat(14) == ctx.stat(8) and ctx.stat(2) == 12
        ctx.equiprop(w.dt, *(1e-3 * w.carr))
        hs_small = ctx.stat(14)
        U = ctx.equiprop(w.dt, *w.carr)
        assert hs_small < ctx.stat(14) < ctx.stat(8)
    assert rel_frobenius(U, equiprop_oracle(w.H0, w.H1, w.carr, w.dt, "none", False, "fp64")) < 1e-13


def test_reference_norm_switch(tmp_path):
    """PARAMENT_NORM=reference (read at Parament_create) builds the series for Hnorm at every dimension: degree 12 again."""
    code = textwrap.dedent(f"""
        import sys, numpy as np
        sys.path.insert(0, {ROOT!r})
        import parament_b200 as pb
        from workloads import make_workload
        w = make_workload("C3", pts=200)
        with pb.Parament("fp64") as ctx:
            ctx.set_hamiltonian(w.H0, *w.H1, quadrature_mode="none")
            U = ctx.equiprop(w.dt, *w.carr)
            assert ctx.stat(2) == 12 and ctx.stat(14) == ctx.stat(8), (ctx.stat(2), ctx.stat(14))
        np.save({str(tmp_path)!r} + "/u.npy", U)
    """)
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, PARAMENT_NORM="reference"), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    w = make_workload("C3", pts=200)
    assert rel_frobenius(np.load(tmp_path / "u.npy"), equiprop_oracle(w.H0, w.H1, w.carr, w.dt, "none", False, "fp64")) < 1e-13


def test_handle_used_after_destroy_is_rejected(pb):
    lib = pb._lib.lib
    h = ctypes.c_void_p()
    assert lib.Parament_create(ctypes.byref(h)) == 0
    assert lib.Parament_destroy(h) == 0
    out = np.zeros(4, dtype=np.complex64)
    assert lib.Parament_equiprop(h, np.zeros(1, dtype=np.complex64), 0.1, 1, 1, out) == 50     # not a crash, not stale memory
    assert lib.Parament_peekAtLastError(h) == 50
    assert lib.Parament_destroy(h) == 0                                                        # idempotent


def test_callers_current_device_is_left_alone(pb, gpu_count):
    torch = pytest.importorskip("torch")
    torch.cuda.set_device(0)
    other = 1 if gpu_count > 1 else 0
    w = make_workload("C2", pts=2001)
    with pb.Parament("fp32", device=other) as ctx:
        ctx.set_hamiltonian(w.H0, *w.H1, quadrature_mode="simpson")
        U = ctx.equiprop(w.dt, *w.carr)
        assert torch.cuda.current_device() == 0
    assert rel_frobenius(U, equiprop_oracle(w.H0, w.H1, w.carr, w.dt, "simpson", False, "fp32")) < 1e-5


# ---------------------------------------------------------------------------------------------------------
# complex64, dim <= 8: FP32 arithmetic as 3xTF32 on the warp-level tensor path (k1_tf32.cu)
# ---------------------------------------------------------------------------------------------------------
def _run_sub(code, env, tmp_path):
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    return r.stdout


TF32_CASES = textwrap.dedent("""
    import sys, numpy as np
    sys.path.insert(0, ROOT)
    import parament_b200 as pb
    from workloads import make_workload, rand_herm
    from oracle.equiprop_oracle import equiprop_oracle, rel_frobenius
    rng = np.random.default_rng(3)
    worst = 0.0
    # dims 2..8, every quadrature, Magnus (more terms than stay in registers), complex amplitudes, ragged lengths, chunked pulses
    for n, A, quad, mag, cplx_amp, pts in [(8, 2, "none", False, False, 1000), (8, 2, "simpson", False, True, 2001), (5, 1, "midpoint", False, False, 333),
                                            (2, 1, "none", False, False, 4000), (7, 3, "simpson", True, False, 801), (8, 4, "none", False, True, 64),
                                            (3, 2, "simpson", True, True, 5), (8, 2, "none", False, False, 1), (6, 2, "midpoint", False, False, 2)]:
        H0 = (0.5 * rand_herm(rng, n)).astype(np.complex64)
        H1 = np.stack([(0.5 / A * rand_herm(rng, n)).astype(np.complex64) for _ in range(A)])
        carr = rng.uniform(-1, 1, (A, pts)) + (1j * rng.uniform(-1, 1, (A, pts)) if cplx_amp else 0)
        carr = carr.astype(np.complex64)
        dt = 0.2 if quad in ("none", "midpoint") else 0.1
        with pb.Parament("fp32") as ctx:
            ctx.set_hamiltonian(H0, *H1, use_magnus=mag, quadrature_mode=quad)
            U = ctx.equiprop(dt, *carr)
            assert ctx.stat(15) == EXPECT_MATH, (n, quad, ctx.stat(15), ctx.stat(9))
        err = rel_frobenius(U, equiprop_oracle(H0, H1, carr, dt, quad, mag, "fp32"))
        assert err < 1e-5, (n, A, quad, mag, pts, err)
        worst = max(worst, err)
    # the C5 ensemble at full size: host-pointer batch call (copy groups) and the first 16 golden pulses + 4 oracle pulses
    w = make_workload("C5")
    with pb.Parament("fp32") as ctx:
        ctx.set_hamiltonian(w.H0, *w.H1)
        U = ctx.equiprop_batch(w.dt, w.carr)
        assert ctx.stat(15) == EXPECT_MATH
        one = ctx.equiprop(w.dt, *w.carr[4321])
    for b in (0, 15, 4321, 9999):
        err = rel_frobenius(U[b], equiprop_oracle(w.H0, w.H1, w.carr[b], w.dt, "none", False, "fp32"))
        assert err < 1e-5, (b, err)
        worst = max(worst, err)
    assert rel_frobenius(U[4321], one) < 2e-6
    print("worst", worst)
""")


@pytest.mark.parametrize("mode,expect", [("tf32", 1), ("f64", 0)])
def test_complex64_small_dim_arithmetic_paths(tmp_path, mode, expect):
    """Both arithmetic paths of complex64 contexts with dim <= 8 meet the 1e-5 tolerance on the same cases; the TF32 path
    must really be the one that ran when selected (Parament_lastStat key 15)."""
    code = TF32_CASES.replace("ROOT", repr(ROOT)).replace("EXPECT_MATH", str(expect))
    out = _run_sub(code, {"PARAMENT_C64_MATH": mode}, tmp_path)
    assert "worst" in out


def test_tf32_path_is_selected_by_step_count(pb):
    """Automatic choice by accumulated phase N h rho <= 128: C5's 1e3 steps run on the TF32 path, long pulses in FP64; dim 16 and
    complex128 contexts always in FP64."""
    w = make_workload("C5", pts=1000, batch=8)
    wl = make_workload("C5", pts=20000, batch=1)
    with pb.Parament("fp32") as ctx:
        ctx.set_hamiltonian(w.H0, *w.H1)
        ctx.equiprop_batch(w.dt, w.carr)
        assert ctx.stat(15) == 1
        U = ctx.equiprop(wl.dt, *wl.carr)
        assert ctx.stat(15) == 0
    assert rel_frobenius(U, equiprop_oracle(wl.H0, wl.H1, wl.carr, wl.dt, "none", False, "fp32")) < 1e-5
    w2 = make_workload("C2", pts=2001)
    with pb.Parament("fp32") as ctx:
        ctx.set_hamiltonian(w2.H0, *w2.H1, quadrature_mode="simpson")
        ctx.equiprop(w2.dt, *w2.carr)
        assert ctx.stat(15) == 2          # dim 9..16: mixed FP64 + TF32 (test_mixed_precision_dim16)
    w1 = make_workload("C1", pts=1001)
    with pb.Parament("fp64") as ctx:
        ctx.set_hamiltonian(w1.H0, *w1.H1, quadrature_mode="midpoint")
        ctx.equiprop(w1.dt, *w1.carr)
        assert ctx.stat(15) == 0


def test_mixed_precision_dim16(pb, monkeypatch):
    """complex64, dim 9..16, degree-8 form: X^2, the low-order terms and the running product in FP64, the two small series
    products as 3xTF32 (Parament_lastStat 15 == 2).  Must meet the tolerance wherever it is selected, agree with the all-FP64
    kernel, cover complex amplitudes / Magnus / ragged dims, and give way to FP64 beyond the accumulated-phase bound."""
    rng = np.random.default_rng(21)
    for n, A, quad, mag, cplx_amp, pts in [(16, 2, "simpson", False, False, 20001), (16, 3, "none", False, True, 3000), (11, 2, "simpson", True, False, 801),
                                            (9, 1, "midpoint", False, True, 77), (16, 2, "none", False, False, 1)]:
        H0 = (0.5 * rand_herm(rng, n)).astype(np.complex64)
        H1 = np.stack([(0.5 / A * rand_herm(rng, n)).astype(np.complex64) for _ in range(A)])
        carr = (rng.uniform(-1, 1, (A, pts)) + (1j * rng.uniform(-1, 1, (A, pts)) if cplx_amp else 0)).astype(np.complex64)
        dt = 0.2 if quad in ("none", "midpoint") else 0.1
        res = {}
        for mode in ("1", "0"):
            monkeypatch.setenv("PARAMENT_K1_MIXED", mode)
            with pb.Parament("fp32") as ctx:
                ctx.set_hamiltonian(H0, *H1, use_magnus=mag, quadrature_mode=quad)
                res[mode] = ctx.equiprop(dt, *carr)
                assert (ctx.stat(15), ctx.stat(9)) == (2, 3) if mode == "1" else ctx.stat(15) == 0
        Uo = equiprop_oracle(H0, H1, carr, dt, quad, mag, "fp32")
        assert rel_frobenius(res["1"], Uo) < 2e-6 and rel_frobenius(res["0"], Uo) < 2e-6, (n, quad, mag)
        assert rel_frobenius(res["1"], res["0"]) < 1e-6
    monkeypatch.delenv("PARAMENT_K1_MIXED")
    w = make_workload("C2", pts=4001)
    with pb.Parament("fp32") as ctx:
        ctx.set_hamiltonian(w.H0, *w.H1, quadrature_mode="simpson")
        ctx.equiprop(w.dt, *w.carr)
        assert ctx.stat(15) == 2
        with pytest.raises(RuntimeError, match="Timestep too large"):
            ctx.equiprop(4000.0 * w.dt, *w.carr[:, :5])    # x = 800: far beyond the tables -> error 70 before any selection
    with pb.Parament("fp32") as ctx:                        # a slice of a very long pulse: the phase of the WHOLE pulse decides
        ctx.set_hamiltonian(w.H0, *w.H1, quadrature_mode="simpson")
        long = np.ascontiguousarray(np.tile(w.carr, (1, 4000))[:, :16000001])
        ctx.equiprop_slice(w.dt, long, 0, 2000)
        assert ctx.stat(15) == 0


def test_pageable_staging_matches_page_locked_input(pb):
    """Large pageable caller buffers are staged by the library's copy threads (context.hpp Stager), page-locked ones are sent
    directly: same kernels, same copy groups, bit-identical propagators -- for a long single pulse (strided group copies), an
    ensemble, and a dim-64 pulse (one plain copy)."""
    torch = pytest.importorskip("torch")
    for name, pts, batch in (("C2", 700001, None), ("C5", 1000, 6000), ("C3", 40000, None)):
        w = make_workload(name, pts=pts, batch=batch)
        carr = np.ascontiguousarray(w.carr.reshape(w.batch, w.amps, w.pts))
        assert carr.nbytes > (2 << 20)
        pinned = torch.from_numpy(carr.copy()).pin_memory().numpy()
        with pb.Parament(w.precision) as ctx:
            ctx.set_hamiltonian(w.H0, *w.H1, use_magnus=w.use_magnus, quadrature_mode=w.quadrature)
            a = ctx.equiprop_batch(w.dt, carr)
            h2d = ctx.stat(6)
            b = ctx.equiprop_batch(w.dt, pinned)
            a2 = ctx.equiprop_batch(w.dt, carr)
        assert h2d == carr.nbytes
        assert np.array_equal(a, b) and np.array_equal(a, a2)
        Uo = equiprop_oracle(w.H0, w.H1, carr[0], w.dt, w.quadrature, w.use_magnus, w.precision, workers=8)
        assert rel_frobenius(a[0], Uo) < TOL[w.precision]


# ---- packed small systems (dim <= 4: two or four systems per 8 x 8 tensor-pipe tile, k1_warp.cu) -------------------------
def _small_system(rng, dim, A, prec, complex_amps, pts, batch):
    ct = np.complex64 if prec == "fp32" else np.complex128
    H0 = (0.7 * rand_herm(rng, dim)).astype(ct)
    H1 = np.stack([(0.4 * rand_herm(rng, dim)).astype(ct) for _ in range(A)])
    carr = rng.uniform(-1, 1, (batch, A, pts))
    if complex_amps:
        carr = carr + 1j * rng.uniform(-1, 1, (batch, A, pts))
    return H0, H1, carr.astype(ct)


@pytest.mark.parametrize("pts", [1, 2, 3, 4, 5, 7, 9, 64, 1001, 20000])
@pytest.mark.parametrize("dim,A,prec,quad,mag,complex_amps", [
    (2, 1, "fp64", "none", False, False), (2, 2, "fp64", "simpson", True, False), (2, 1, "fp32", "midpoint", False, True),
    (1, 1, "fp64", "none", False, False), (3, 2, "fp64", "midpoint", False, True), (3, 1, "fp32", "none", False, False),
    (4, 2, "fp64", "simpson", False, False), (4, 3, "fp32", "simpson", True, True), (4, 1, "fp64", "none", False, False)])
def test_packed_small_systems_single_pulse(pb, dim, A, prec, quad, mag, complex_amps, pts):
    """Every length around the pack factor (blocks that run out of steps early), all quadratures, Magnus, complex amplitudes."""
    if quad == "simpson" and pts < 3:
        pts = 3
    if quad == "simpson" and pts % 2 == 0:
        pts += 1
    if quad == "midpoint" and pts < 2:
        pts = 2
    rng = np.random.default_rng(1000 * dim + pts)
    H0, H1, carr = _small_system(rng, dim, A, prec, complex_amps, pts, 1)
    dt = 0.05
    with pb.Parament(prec) as ctx:
        ctx.set_hamiltonian(H0, *H1, use_magnus=mag, quadrature_mode=quad)
        U = ctx.equiprop(dt, *carr[0])
        assert ctx.stat(5) == 1
    assert rel_frobenius(U, equiprop_oracle(H0, H1, carr[0], dt, quad, mag, prec)) < TOL[prec]


@pytest.mark.parametrize("dim,prec,batch,pts", [(2, "fp64", 5, 33), (2, "fp32", 9000, 17), (4, "fp64", 7000, 10), (3, "fp32", 300, 201),
                                                (4, "fp32", 2, 5000), (2, "fp64", 40000, 3)])
def test_packed_small_systems_ensembles(pb, dim, prec, batch, pts):
    """Ensembles in both plans (a warp owns a pulse or a part of it; few long pulses spread over CTAs)."""
    rng = np.random.default_rng(77 + dim + batch)
    H0, H1, carr = _small_system(rng, dim, 2, prec, False, pts, batch)
    dt = 0.04
    with pb.Parament(prec) as ctx:
        ctx.set_hamiltonian(H0, *H1, quadrature_mode="none")
        U = ctx.equiprop_batch(dt, carr)
    for b in sorted(set([0, 1, batch // 2, batch - 1])):
        assert rel_frobenius(U[b], equiprop_oracle(H0, H1, carr[b], dt, "none", False, prec)) < TOL[prec], b
    if batch > 100:   # all pulses, against the first one's error scale: nothing is mixed up between blocks or pulses
        Us = np.stack([equiprop_oracle(H0, H1, carr[b], dt, "none", False, prec) for b in range(0, batch, max(1, batch // 64))])
        got = U[::max(1, batch // 64)]
        assert max(rel_frobenius(g, u) for g, u in zip(got, Us)) < TOL[prec]


def test_packing_switch_gives_the_same_propagator(pb, monkeypatch):
    """PARAMENT_K1_PACK=0 (one system per tile, as for dim 5..8) and the packed kernel agree to rounding."""
    rng = np.random.default_rng(5)
    H0, H1, carr = _small_system(rng, 2, 2, "fp64", False, 4097, 1)
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("PARAMENT_K1_PACK", mode)
        with pb.Parament("fp64") as ctx:
            ctx.set_hamiltonian(H0, *H1, quadrature_mode="midpoint")
            res[mode] = ctx.equiprop(0.02, *carr[0])
    assert rel_frobenius(res["1"], res["0"]) < 1e-13
    assert rel_frobenius(res["1"], equiprop_oracle(H0, H1, carr[0], 0.02, "midpoint", False, "fp64")) < 1e-12


@pytest.mark.parametrize("dim,prec,quad", [(2, "fp64", "simpson"), (4, "fp64", "none"), (2, "fp32", "midpoint")])
def test_packed_small_systems_long_pulse_host_pipeline(pb, dim, prec, quad):
    """A pulse long enough for the copy / compute groups of the host-pointer path (several launches of partial products,
    each unpacked to an ordinary padded partial before the ordered reduction)."""
    pts = 600_001 if prec == "fp64" else 1_200_001
    rng = np.random.default_rng(dim)
    H0, H1, carr = _small_system(rng, dim, 2, prec, False, pts, 1)
    dt = 0.002
    with pb.Parament(prec) as ctx:
        ctx.set_hamiltonian(H0, *H1, quadrature_mode=quad)
        U = ctx.equiprop(dt, *carr[0])
        assert ctx.stat(1) >= 3          # copy groups: more than one chain launch + the reduction
    assert rel_frobenius(U, equiprop_oracle(H0, H1, carr[0], dt, quad, False, prec)) < TOL[prec]


# ---- Hermitian shortcut of the dim 9..16 degree-8 kernels (right-operand layout of X without shuffles) -------------------
@pytest.mark.parametrize("prec,dim", [("fp32", 16), ("fp32", 11)])
def test_non_hermitian_inputs_take_the_general_path(pb, prec, dim):
    """The API accepts any matrices (the reference never checks): a non-Hermitian control Hamiltonian must not use the shortcut."""
    rng = np.random.default_rng(3)
    ct = np.complex64
    H0 = (0.5 * rand_herm(rng, dim)).astype(ct)
    G = rng.normal(size=(dim, dim)) + 1j * rng.normal(size=(dim, dim))
    H1 = np.stack([(0.3 * rand_herm(rng, dim) + 0.02 * G / np.linalg.norm(G, 2)).astype(ct), (0.2 * rand_herm(rng, dim)).astype(ct)])
    assert np.abs(H1[0] - H1[0].conj().T).max() > 1e-4
    carr = rng.uniform(-1, 1, (2, 3001)).astype(ct)
    with pb.Parament(prec) as ctx:
        ctx.set_hamiltonian(H0, *H1, quadrature_mode="simpson")
        U = ctx.equiprop(0.02, *carr)
    assert rel_frobenius(U, equiprop_oracle(H0, H1, carr, 0.02, "simpson", False, prec)) < TOL[prec]


def test_hermitian_shortcut_switch_gives_the_same_propagator(pb, monkeypatch):
    w = make_workload("C2", pts=40001)
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("PARAMENT_K1_HERM", mode)
        with pb.Parament(w.precision) as ctx:
            ctx.set_hamiltonian(w.H0, *w.H1, use_magnus=w.use_magnus, quadrature_mode=w.quadrature)
            res[mode] = ctx.equiprop(w.dt, *w.carr)
    assert rel_frobenius(res["1"], res["0"]) < 2e-7          # complex64 output rounding
    assert rel_frobenius(res["1"], equiprop_oracle(w.H0, w.H1, w.carr, w.dt, w.quadrature, w.use_magnus, w.precision)) < TOL["fp32"]


@pytest.mark.parametrize("A", [3, 5])      # 5: more control terms than the kernel keeps in registers
@pytest.mark.parametrize("dim,quad,complex_amps,hermitian", [(8, "none", True, True), (6, "simpson", True, True), (8, "midpoint", False, True),
                                                             (8, "none", False, False), (7, "simpson", True, False), (5, "midpoint", True, True)])
def test_tf32_kernel_hermitian_and_general_variants(pb, dim, quad, complex_amps, hermitian, A):
    """complex64, dim 5..8, short pulse (3xTF32 kernel): Hermitian tables (one register table per matrix, X^T = conj(X) for real
    coefficients, conjugated tables for complex ones) and the general variant for non-Hermitian inputs."""
    rng = np.random.default_rng(100 + dim)
    ct = np.complex64
    H0 = (0.5 * rand_herm(rng, dim)).astype(ct)
    H1 = np.stack([(0.9 / A * rand_herm(rng, dim)).astype(ct) for _ in range(A)])
    if not hermitian:
        G = rng.normal(size=(dim, dim)) + 1j * rng.normal(size=(dim, dim))
        H1[A - 1] = (H1[A - 1] + 0.03 * G / np.linalg.norm(G, 2)).astype(ct)
    pts = 401
    carr = rng.uniform(-1, 1, (A, pts))
    if complex_amps:
        carr = carr + 1j * rng.uniform(-1, 1, (A, pts))
        carr[1].imag = 0.0          # a mix of real and complex coefficients within a step
    carr = carr.astype(ct)
    dt = 0.05
    with pb.Parament("fp32") as ctx:
        ctx.set_hamiltonian(H0, *H1, quadrature_mode=quad)
        U = ctx.equiprop(dt, *carr)
        assert ctx.stat(15) == 1
    assert rel_frobenius(U, equiprop_oracle(H0, H1, carr, dt, quad, False, "fp32")) < TOL["fp32"]


# ---- dim > 64: Hermitian Y -> only the upper-triangular tiles of Y Y are computed, the rest mirrored --------------------------
@pytest.mark.parametrize("dim,quad,complex_amps,hermitian,series,dt", [
    (128, "none", False, True, None, 0.02), (100, "simpson", False, True, None, 0.02), (256, "midpoint", False, True, None, 0.02),
    (128, "none", True, True, None, 0.02), (128, "none", False, False, None, 0.02), (192, "none", False, True, "horner", 0.02),
    (128, "simpson", False, True, "horner", 0.4), (70, "none", False, True, None, 0.3)])
def test_hermitian_square_in_batched_gemm(pb, monkeypatch, dim, quad, complex_amps, hermitian, series, dt):
    """Real amplitudes + Hermitian matrices take the mirrored-tile GEMM for Y Y (and W W in the blocks-of-four form); complex
    amplitudes and non-Hermitian inputs must not.  Both against the oracle, and against each other through PARAMENT_K4_HERM=0."""
    rng = np.random.default_rng(dim)
    ct = np.complex128
    H0 = (0.5 * rand_herm(rng, dim)).astype(ct)
    H1 = np.stack([(0.3 * rand_herm(rng, dim)).astype(ct) for _ in range(2)])
    if not hermitian:
        G = rng.normal(size=(dim, dim)) + 1j * rng.normal(size=(dim, dim))
        H1[0] = H1[0] + 0.02 * G / np.linalg.norm(G, 2)
    pts = 31
    carr = rng.uniform(-1, 1, (2, pts))
    if complex_amps:
        carr = carr + 1j * rng.uniform(-0.3, 0.3, (2, pts))
    carr = carr.astype(ct)
    if series:
        monkeypatch.setenv("PARAMENT_SERIES", series)
    res, forms = {}, {}
    for mode in ("1", "0"):
        monkeypatch.setenv("PARAMENT_K4_HERM", mode)
        with pb.Parament("fp64") as ctx:
            ctx.set_hamiltonian(H0, *H1, quadrature_mode=quad)
            res[mode] = ctx.equiprop(dt, *carr)
            assert ctx.stat(5) == 3
            forms[mode] = ctx.stat(9)
    assert rel_frobenius(res["1"], equiprop_oracle(H0, H1, carr, dt, quad, False, "fp64")) < TOL["fp64"], forms
    assert rel_frobenius(res["1"], res["0"]) < 1e-13
