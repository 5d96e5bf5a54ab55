"""CPU checks (numpy) of the algebraic identities the round-2 kernels rely on -- the GPU parity tests prove the kernels, these pin
the reasoning behind them so that a maintainer can change a kernel and see which identity it was built on:
  * packed small systems (k1_warp.cu p.pack, k1_common.cuh unpack_blocks): block-diagonal tiles, sub-ranges of the step range per block,
    log2(pack) shift-and-multiply products to combine the blocks in time order;
  * Hermitian step matrices: right-operand layout by conjugation, mirrored tiles of Y Y with a fused epilogue (k4_gemm.cu), the
    transposed assembly of the TF32 kernel from conjugated tables (k1_tf32.cu HERM).
"""
import numpy as np
import pytest

from oracle.equiprop_oracle import equiprop_oracle
from workloads import rand_herm


def _step_propagators(H0, H1, carr, dt):
    """Exact step propagators exp(-i dt (H0 + sum_k c_k(j) H_k)) by eigendecomposition (Hermitian case)."""
    out = []
    for j in range(carr.shape[1]):
        X = H0 + np.tensordot(carr[:, j], H1, axes=1)
        w, V = np.linalg.eigh(X)
        out.append((V * np.exp(-1j * dt * w)) @ V.conj().T)
    return out


@pytest.mark.parametrize("dim,pack", [(2, 4), (1, 4), (4, 2), (3, 2)])
@pytest.mark.parametrize("L", [1, 2, 3, 4, 5, 7, 8, 13, 64])
def test_packed_blocks_combine_in_time_order(dim, pack, L):
    rng = np.random.default_rng(10 * dim + L)
    nb = 8 // pack
    H0 = 0.7 * rand_herm(rng, dim)
    H1 = np.stack([0.4 * rand_herm(rng, dim) for _ in range(2)])
    carr = rng.uniform(-1, 1, (2, L))
    U = _step_propagators(H0, H1, carr, 0.05)
    lo, hi = 0, L
    # the kernel: block b walks [lo + L b / pack, lo + L (b + 1) / pack); the tile holds the TRANSPOSED running products on its diagonal
    Q = np.eye(8, dtype=complex)
    iters = (L + pack - 1) // pack
    jb = [lo + L * b // pack for b in range(pack)]
    hb = [lo + L * (b + 1) // pack for b in range(pack)]
    assert max(h - j for j, h in zip(jb, hb)) == iters and jb[0] == lo and hb[-1] == hi
    for it in range(iters):
        E = np.zeros((8, 8), dtype=complex)                     # E = U - I per block, zero for a block that ran out of steps
        for b in range(pack):
            j = jb[b] + it
            if j < hb[b]:
                E[b * nb:b * nb + dim, b * nb:b * nb + dim] = U[j] - np.eye(dim)
        Q = Q + Q @ E.T                                          # Q <- Q (I + E)^T
    # unpack_blocks: Q <- Q shift(Q), shift by nb, 2 nb, ... rows and columns (block b + s moves to block b)
    sh = nb
    while sh < 8:
        idx = (np.arange(8) + sh) % 8
        Q = Q @ Q[np.ix_(idx, idx)]
        sh *= 2
    got = Q[:dim, :dim].T                                        # the stored propagator is the transpose of block 0
    exp = np.eye(dim, dtype=complex)
    for j in range(L):
        exp = U[j] @ exp
    assert np.linalg.norm(got - exp) < 1e-13 * max(1.0, np.linalg.norm(exp))


def test_packed_result_equals_the_oracle():
    """The same construction against the oracle's Chebyshev evaluation (the reference's algorithm), dim 2, four blocks."""
    rng = np.random.default_rng(3)
    H0 = (0.7 * rand_herm(rng, 2)).astype(np.complex128)
    H1 = np.stack([(0.4 * rand_herm(rng, 2)).astype(np.complex128)])
    carr = rng.uniform(-1, 1, (1, 23)).astype(np.complex128)
    Uo = equiprop_oracle(H0, H1, carr, 0.05, "none", False, "fp64")
    exp = np.eye(2, dtype=complex)
    for Uj in _step_propagators(H0, H1, carr.real, 0.05):
        exp = Uj @ exp
    assert np.linalg.norm(Uo - exp) < 1e-12


@pytest.mark.parametrize("n,BM,BN", [(256, 64, 32), (128, 64, 32), (128, 64, 64), (96, 32, 32)])
def test_mirrored_tiles_of_a_hermitian_square(n, BM, BN):
    """k4_gemm.cu GemmArgs::herm: only tiles with BM i < BN (j + 1) are computed; a computed element (r, c) whose transposed position
    lies in a skipped tile also writes there every output of the fused epilogue evaluated on the conjugated product and addends."""
    rng = np.random.default_rng(n)
    Y = rand_herm(rng, n)
    c4, c3 = 0.37, -0.21
    P = Y @ Y
    W_full, T_full = P, c4 * P + 1j * c3 * Y                     # Dprod = W, D = alpha P + beta Y with beta = i c3
    W = np.full((n, n), np.nan, dtype=complex)
    T = np.full((n, n), np.nan, dtype=complex)
    writes = np.zeros((n, n), dtype=int)
    skipped = lambda i, j: BM * i >= BN * (j + 1)
    for i in range(n // BM):
        for j in range(n // BN):
            if skipped(i, j):
                continue
            for r in range(BM * i, BM * i + BM):
                for c in range(BN * j, BN * j + BN):
                    W[r, c] = P[r, c]
                    T[r, c] = c4 * P[r, c] + 1j * c3 * Y[r, c]
                    writes[r, c] += 1
                    if skipped(c // BM, r // BN):
                        W[c, r] = np.conj(P[r, c])
                        T[c, r] = c4 * np.conj(P[r, c]) + 1j * c3 * np.conj(Y[r, c])
                        writes[c, r] += 1
    assert (writes == 1).all()                                   # every element written exactly once
    assert np.linalg.norm(W - W_full) < 1e-12 * np.linalg.norm(W_full)
    assert np.linalg.norm(T - T_full) < 1e-12 * np.linalg.norm(T_full)


def test_right_operand_of_a_hermitian_matrix_is_its_conjugated_transpose_layout():
    """frag.cuh: BFrag(E^T) is a relabeling of the accumulator registers of E; for Hermitian X, X = conj(X^T), so BFrag(X) is that
    relabeling with the imaginary parts negated (conj_transpose_as_bfrag).  Written out on matrices: element (k, n) of the right
    operand is X[k][n] = conj(X[n][k])."""
    rng = np.random.default_rng(1)
    X = rand_herm(rng, 16)
    W = X @ X
    assert np.abs(X - X.conj().T).max() == 0.0
    assert np.abs(W - W.conj().T).max() < 1e-14                  # W is Hermitian to rounding: its conjugated transpose stands in for it
    # the three-real-product form keeps Re W exactly symmetric (same products in the same order) but not Im W
    Ar, Ai = X.real, X.imag
    P1, P2 = Ar @ Ar, Ai @ Ai
    assert np.abs((P1 - P2) - (P1 - P2).T).max() < 1e-15


@pytest.mark.parametrize("complex_coeffs", [False, True])
def test_transposed_assembly_from_conjugated_tables(complex_coeffs):
    """k1_tf32.cu HERM: with Hermitian tables Z_k the registers of Z_k^T are conj(Z_k), so X^T = conj(X) + D with
    D = sum_k 2 Im(c_k) (Im Z_k + i Re Z_k) -- zero for real coefficients, where X is Hermitian itself."""
    rng = np.random.default_rng(8)
    H0 = rand_herm(rng, 8)
    Z = [rand_herm(rng, 8) for _ in range(3)]
    c = rng.uniform(-1, 1, 3) + (1j * rng.uniform(-1, 1, 3) if complex_coeffs else 0)
    X = H0 + sum(ck * Zk for ck, Zk in zip(c, Z))
    D = sum(2 * ck.imag * (Zk.imag + 1j * Zk.real) for ck, Zk in zip(np.atleast_1d(c).astype(complex), Z))
    assert np.linalg.norm(X.T - (X.conj() + D)) < 1e-14
    if not complex_coeffs:
        assert np.abs(D).max() == 0.0 and np.abs(X - X.conj().T).max() < 1e-15
