"""The reference's solver tests restated against the oracle: dense LCP (lcp.cc:412-528), dense and
matrix-free Jacobi / Gauss-Seidel / SOR (sparse_iterations.cc:355-748) and the matrix-free block
products / triangular solves against the explicit J M^-1 J^T (sparse_iterations_utils.cc:938-1248)."""
import numpy as np
import pytest

from oracle import pyoracle as O

TOL = 1e-9   # kAllowNumericalError


def _spd(rng, n):
    """GenerateSPDMatrix, utils.cc:203-216 (well-conditioned m^T m)."""
    while True:
        m = rng.uniform(-1, 1, (n, n))
        A = m.T @ m
        if np.linalg.cond(A) <= 1e7:
            return A


def _diag_dominant(rng, n):
    """GenerateDiagonalDominantMatrix, utils.cc:218-231."""
    while True:
        A = rng.uniform(-1, 1, (n, n))
        scale = np.abs(A).max() / np.abs(A).min()
        A[np.diag_indices(n)] *= scale
        if np.linalg.cond(A) <= 1e7:
            return A


def test_murty_random_spd_no_bounds():          # lcp.cc:412-437
    rng = np.random.default_rng(20)
    for _ in range(25):
        A, b = _spd(rng, 50), rng.uniform(-1, 1, 50)
        ok, x, w, it, S = O.murty(A, b)
        assert ok and np.linalg.norm(A @ x - b - w) < TOL
        assert np.all(x >= 0) and np.all(w[x == 0] >= -TOL)


def test_murty_random_spd_with_bounds():        # lcp.cc:439-469
    rng = np.random.default_rng(21)
    for _ in range(25):
        A, b = _spd(rng, 50), rng.uniform(-1, 1, 50)
        lo, hi = -rng.uniform(0, 2), rng.uniform(1e-3, 2)
        ok, x, w, it, S = O.murty(A, b, lo, hi)
        assert ok and np.linalg.norm(A @ x - b - w) < TOL
        assert np.all(x >= lo) and np.all(x <= hi)


@pytest.mark.parametrize("honour", [False, True])
def test_mixed_constraints_solver(honour):      # lcp.cc:471-528
    rng = np.random.default_rng(22)
    for _ in range(20):
        n = 50
        A, b = _spd(rng, n), rng.uniform(-1, 1, n)
        C = rng.integers(0, 2, n)
        lo, hi = (np.full(n, -10.0), np.full(n, 10.0)) if honour else (np.zeros(n), np.full(n, np.inf))
        ok, x, w, it = O.mixed_solver(A, b, C, lo, hi, honour_bounds=honour)
        assert ok and np.linalg.norm(A @ x - b - w) < TOL
        ineq = C == 0
        assert np.all(x[ineq] >= lo[ineq]) and np.all(x[ineq] <= hi[ineq])
        assert np.all(w[C == 1] == 0)


def test_mixed_solver_ignores_bounds_quirk_q1():
    """lcp.cc:298 calls the 4-argument Murty: the passed x_lo/x_hi never reach the solver."""
    rng = np.random.default_rng(23)
    A, b = _spd(rng, 12), rng.uniform(-1, 1, 12)
    C = np.zeros(12, dtype=int)
    r1 = O.mixed_solver(A, b, C, np.full(12, -1.0), np.full(12, 1.0), honour_bounds=False)
    r2 = O.mixed_solver(A, b, C, np.zeros(12), np.full(12, np.inf), honour_bounds=False)
    assert np.array_equal(r1[1], r2[1])


def check_mixed(A, b, x, C, lo, hi):
    """CheckMixedConstraintSolutions, sparse_iterations.cc:308-353."""
    r = A @ x - b
    if not np.linalg.norm(r[C == 1]) < TOL:
        return False
    for i in np.nonzero(C == 0)[0]:
        if lo[i] < x[i] < hi[i]:
            if not abs(r[i]) < TOL:
                return False
        elif x[i] == lo[i]:
            if not r[i] > -TOL:
                return False
        elif x[i] == hi[i]:
            if not r[i] < TOL:
                return False
        else:
            return False
    return True


@pytest.mark.parametrize("itype", [O.IT_JACOBI, O.IT_GS, O.IT_SOR])
def test_dense_iterations_equalities(itype):    # sparse_iterations.cc:355-460
    rng = np.random.default_rng(30 + itype)
    for _ in range(10):
        n = int(rng.integers(3, 51))
        A = _diag_dominant(rng, n) if itype == O.IT_JACOBI else _spd(rng, n) + (0.0 if itype == O.IT_GS else 0.0) * np.eye(n)
        if itype != O.IT_JACOBI:
            A = A + 0.5 * np.eye(n)              # the reference shifts its SPD test matrices (":SPD(+shift)")
        b = rng.uniform(-1, 1, n)
        x, sweeps = O.dense_iteration(A, b, itype, k_max=5000)
        if sweeps < 5000:
            assert np.linalg.norm(A @ x - b) < 1e-8


def _stepped_worlds(kind, steps=20):
    """Chain(4,(0,0,2)) / Cairn(4,...) stepped with Ensemble::Step as in sparse_iterations.cc:600-640."""
    W = O.World()
    if kind == "chain":
        W.build_chain(4, [0, 0, 2])
    else:
        W.build_cairn(4, [-0.2, 0.2], [-0.2, 0.2], [0.16, 0.6])
    W.set_params(solver=O.SOLVER_DENSE_MURTY)
    assert W.init() == 0
    yield W
    for _ in range(steps):
        W.step(0.001)
        W.update_contacts()
        yield W


@pytest.mark.parametrize("kind", ["chain", "cairn"])
def test_matrix_free_products_match_dense(kind):      # sparse_iterations_utils.cc:760-935
    rng = np.random.default_rng(40)
    eps = 0.01                                        # kEpsilonDiagonal
    seen_contacts = False
    for W in _stepped_worlds(kind, steps=12):
        nr = 3 * (W.n_joints + W.n_contacts)
        if nr == 0:
            continue
        seen_contacts |= W.n_contacts > 0
        A = W.dense_A(0.0)
        x = rng.uniform(-1, 1, nr)
        L, U = np.tril(A, -1), np.triu(A, 1)
        assert np.allclose(W.sparse_product(0, x, cfm=eps), (A + eps * np.eye(nr)) @ x, atol=TOL)
        assert np.allclose(W.sparse_product(1, x), U @ x, atol=TOL)
        assert np.allclose(W.sparse_product(2, x), L @ x, atol=TOL)
        assert np.allclose(W.sparse_product(4, x), (L + U) @ x, atol=TOL)
        assert np.allclose(W.sparse_product(3, x, cfm=eps, scale=0.7), (np.diag(A) + eps) * 0.7 * x, atol=TOL)
        # triangular / diagonal solves without projection quirks on an all-equality reading
        Ae = A + eps * np.eye(nr)
        rows = W.rows()
        if np.all(rows["type"] == 1):
            assert np.allclose(W.sparse_solve(0, x, cfm=eps), np.linalg.solve(np.tril(Ae), x), atol=1e-8)
            assert np.allclose(W.sparse_solve(1, x, cfm=eps), np.linalg.solve(np.triu(Ae), x), atol=1e-8)
            assert np.allclose(W.sparse_solve(2, x, cfm=eps), x / np.diag(Ae), atol=1e-8)
    if kind == "cairn":
        assert seen_contacts


@pytest.mark.parametrize("kind", ["chain", "cairn"])
@pytest.mark.parametrize("itype", [O.IT_JACOBI, O.IT_GS, O.IT_SOR])
def test_matrix_free_iterations_on_ensembles(kind, itype):   # sparse_iterations.cc:515-748
    rng = np.random.default_rng(50 + itype)
    cfm = 0.1                                                # kCfmCoeff of the reference tests
    for k, W in enumerate(_stepped_worlds(kind, steps=10)):
        nr = 3 * (W.n_joints + W.n_contacts)
        if nr == 0:
            continue
        rows = W.rows()
        A = W.dense_A(cfm)
        rhs = rng.uniform(-1, 1, nr)
        # quirk-free run: honest per-row bounds => must satisfy the reference's solution check
        x, sweeps = W.sparse_iteration(itype, rhs, cfm=cfm, quirk=False, k_max=3000)
        if itype == O.IT_JACOBI and sweeps >= 3000:
            continue                                         # Jacobi need not converge on every A
        assert check_mixed(A, rhs, x, rows["type"].astype(int), rows["lo"], rows["hi"]), (kind, itype, k, sweeps)
        # matrix-free == dense-matrix variant of the same iteration (same sweeps, same x)
        xd, sd = O.dense_iteration(A, rhs, itype, rows["type"], rows["lo"], rows["hi"], k_max=3000)
        assert sd == sweeps and np.allclose(x, xd, atol=1e-9)


def test_gs_bounds_shift_quirk_q2():
    """With joints followed by contacts the reference projects block i with the bounds of block
    i-1 (sparse_iterations_utils.cc:169,180,229-235): the first contact is left unclamped."""
    W = O.World()
    W.build_chain(3, [0, 0, 0.2])          # hangs low enough for ground contacts
    W.set_params(solver=O.SOLVER_PGS)
    assert W.init() == 0
    W.update_contacts()
    assert W.n_contacts > 0
    nr = 3 * (W.n_joints + W.n_contacts)
    rhs = -np.ones(nr)                     # pushes every multiplier negative
    xq, _ = W.sparse_iteration(O.IT_GS, rhs, cfm=0.1, quirk=True)
    xn, _ = W.sparse_iteration(O.IT_GS, rhs, cfm=0.1, quirk=False)
    first = 3 * W.n_joints
    assert xn[first + 2] >= 0              # honest bounds clamp the first contact's normal row
    assert xq[first + 2] < 0               # the reference leaves it unclamped
    assert np.all(xq[first + 5::3] >= 0)   # later contacts are clamped in both
