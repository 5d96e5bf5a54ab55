"""Acceptance tests of the installed `parament` package (`pytest --pyargs parament`, the reference's CI command, README.md:83-87).
They restate what the reference's own parament/test/ checks -- numerics against scipy.linalg.expm at the reference's thresholds
(test_numerics.py:32-117), the error paths (test_error.py:26-54) and the wrapper life cycle (test_wrapper.py:20-45) -- against
this library.  A CUDA device is required, as for the reference ("parament itself requires a GPU")."""
import ctypes

import numpy as np
import pytest
import scipy.linalg

import parament


def _have_gpu():
    try:
        cuda = ctypes.CDLL("libcuda.so.1")
    except OSError:
        return False
    n = ctypes.c_int(0)
    return cuda.cuInit(0) == 0 and cuda.cuDeviceGetCount(ctypes.byref(n)) == 0 and n.value > 0


pytestmark = pytest.mark.skipif(not _have_gpu(), reason="no CUDA device")


def random_hamiltonians(dim):
    rng = np.random.default_rng(27)
    H0 = rng.uniform(-1, 1, (dim, dim)) / dim + 1j * rng.uniform(-1, 1, (dim, dim)) / 2
    return H0 + H0.conj().T, np.zeros((dim, dim))


@pytest.mark.parametrize("precision,eps", [("fp32", np.finfo(np.float32).eps), ("fp64", np.finfo(np.float64).eps)])
@pytest.mark.parametrize("dim", [2, 16])
def test_expm_scipy_random(precision, eps, dim):
    H0, H1 = random_hamiltonians(dim)
    dt = 0.01
    with parament.Parament(precision=precision) as ctx:
        ctx.set_hamiltonian(H0, H1)
        U = ctx.equiprop(dt, np.zeros(1))
    ct = np.complex64 if precision == "fp32" else np.complex128
    ref = scipy.linalg.expm(-1j * dt * H0.astype(ct).astype(np.complex128))
    assert np.sum(np.abs(U - ref)) < eps * dim ** 2


@pytest.mark.parametrize("dim", [2, 4])
def test_expm_of_a_general_matrix(dim):
    rng = np.random.default_rng(dim)
    m = 0.2 * (rng.standard_normal((dim, dim)) + 1j * rng.standard_normal((dim, dim)))
    assert np.sum(np.abs(parament.debug_functions.expm(m) - scipy.linalg.expm(m))) < np.finfo(np.float32).eps * dim ** 2 * 4


@pytest.mark.parametrize("dim", [2, 16])
def test_multi_fields(dim):
    rng = np.random.default_rng(dim)
    herm = lambda: (lambda g: (g + g.conj().T) / (2 * dim))(rng.standard_normal((dim, dim)) + 1j * rng.standard_normal((dim, dim)))
    H0, H1, H2 = herm(), herm(), herm()
    a, b = rng.uniform(0, 1, 10), rng.uniform(0, 1, 10)
    dt = 0.1
    with parament.Parament() as ctx:
        ctx.set_hamiltonian(H0, H1, H2)
        U = ctx.equiprop(dt, a, b)
    ref = np.eye(dim, dtype=np.complex128)
    c64 = lambda m: m.astype(np.complex64).astype(np.complex128)
    for x, y in zip(a.astype(np.float32).astype(np.float64), b.astype(np.float32).astype(np.float64)):
        ref = scipy.linalg.expm(-1j * dt * (c64(H0) + x * c64(H1) + y * c64(H2))) @ ref      # later step on the left
    assert np.linalg.norm(U - ref) < 1e-6


def test_docstring_known_answer():
    """parament.py:61-68: H0 = sigma_z, H1 = sigma_x, dt = 1, zero amplitude -> diag(exp(-+i))."""
    with parament.Parament(precision="fp64") as ctx:
        ctx.set_hamiltonian(np.array([[1, 0], [0, -1]]), np.array([[0, 1], [1, 0]]))
        U = ctx.equiprop(1.0, np.zeros(1))
    assert np.allclose(U, np.diag([np.exp(-1j), np.exp(1j)]), atol=1e-14)


def test_error_paths():
    ctx = parament.Parament()
    ctx.destroy()
    with pytest.raises(RuntimeError, match="Attempting to use a context that has been destroyed"):
        ctx.set_hamiltonian(np.eye(2), np.eye(2))
    with parament.Parament() as ctx:
        with pytest.raises(RuntimeError, match="No hamiltonian set"):
            ctx.equiprop(1.0)
        for mode in ("none", "midpoint"):
            with pytest.raises(ValueError, match="Invalid quadrature selection"):
                ctx.set_hamiltonian(np.eye(2), np.eye(2), use_magnus=True, quadrature_mode=mode)
        ctx.set_hamiltonian(np.eye(2), np.eye(2))
        with pytest.raises(ValueError, match="Got 2 amplitude arrays, but there are only 1 Hamiltonians."):
            ctx.equiprop(1.0, np.zeros(4), np.zeros(4))
        ctx.set_hamiltonian(np.eye(2), np.eye(2), np.eye(2))
        with pytest.raises(ValueError, match="All amplitude arrays must have the same length."):
            ctx.equiprop(1.0, np.zeros(4), np.zeros(5))


def test_life_cycle_and_device_info(capfd):
    ctx = parament.Parament()
    ctx.set_hamiltonian(np.eye(2), np.eye(2))
    ctx.destroy()
    with parament.Parament(precision="fp64") as ctx:
        ctx.set_hamiltonian(np.eye(3), np.eye(3), quadrature_mode="simpson")
    parament.device_info()
    assert "Total number of CUDA devices" in capfd.readouterr().out
