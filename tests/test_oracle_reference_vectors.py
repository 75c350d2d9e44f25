"""Pins the CPU oracle against every literal known-answer vector the reference's own tests hold for
the step path: /root/reference/eggshell/lcp.cc:348-389 (5x5 LCPs) and utils.cc:398-497 (integer
gather/scatter), plus the exact-arithmetic identities of utils.cc:329-395,499-515."""
import numpy as np

from oracle import pyoracle as O

A1 = np.array([2.1104, 1.4090, 1.5055, 1.3060, 1.1413, 1.4090, 1.9846, 1.7126, 1.0858, 1.9358, 1.5055, 1.7126,
               2.1673, 1.3226, 1.5765, 1.3060, 1.0858, 1.3226, 1.2704, 0.8927, 1.1413, 1.9358, 1.5765, 0.8927,
               2.1211]).reshape(5, 5)
B1 = np.array([0.6691, 0.1904, 0.3689, 0.4607, 0.9816])
A2 = np.array([2.7345, 1.8859, 2.0785, 1.9442, 1.9567, 1.8859, 2.2340, 2.0461, 2.3164, 2.0875, 2.0785, 2.0461,
               2.7591, 2.4606, 1.9473, 1.9442, 2.3164, 2.4606, 2.5848, 2.2768, 1.9567, 2.0875, 1.9473, 2.2768,
               2.4853]).reshape(5, 5)
B2 = np.array([0.7577, 0.7431, 0.3922, 0.6555, 0.1712])


def test_check_murty_solution_lcp_cc_348():
    x = np.array([0.0942, 0, 0, 0, 0.4121])
    w = np.array([0, 0.7401, 0.4226, 0.0302, 0])
    ok, _ = O.check_murty_solution(A1, B1, x, w, [1, 0, 0, 0, 1], err=1e-4)
    assert ok
    x2 = np.array([0.0942, 0, 0.5678, 0, 0.4121])
    w2 = np.array([0, 0.7401, 0.4226, -0.0302, 0])
    ok, _ = O.check_murty_solution(A1, B1, x2, w2, [1, 0, 0, 0, 1], err=1e-4)
    assert not ok


def test_murty_simple_lcp_cc_367():
    ok, x, w, it, S = O.murty(A1, B1)
    assert ok
    assert np.linalg.norm(x - [0.0942, 0, 0, 0, 0.4121]) <= 5e-4
    assert np.linalg.norm(w - [0, 0.7401, 0.4226, 0.0302, 0]) <= 5e-4
    assert S.tolist() == [1, 0, 0, 0, 1]
    # second case: the reference's asserts are commented out (lcp.cc:403,408); we still meet them
    ok, x, w, it, S = O.murty(A2, B2)
    assert ok
    assert np.linalg.norm(x - [0.1141, 0.2363, 0, 0, 0]) <= 5e-4
    assert np.linalg.norm(w - [0, 0, 0.3285, 0.1138, 0.5454]) <= 5e-4


def test_select_submatrix_utils_cc_398():
    S = np.array([1, 1, 0, 0, 1, 0, 1, 1])
    A = np.array([44, 23, 81, 97, 37, 34, 72, 51, 12, 12, 3, 55, 99, 68, 91, 48, 26, 30, 93, 53, 4, 14, 90, 91, 41, 32, 74,
                  24, 89, 73, 34, 61, 60, 43, 49, 49, 92, 11, 70, 62, 27, 51, 58, 63, 80, 66, 20, 86, 61, 9, 24, 68, 10, 50,
                  4, 81, 72, 27, 46, 40, 27, 78, 75, 58], dtype=float).reshape(8, 8)
    A_SS = np.array([44, 23, 37, 72, 51, 12, 12, 99, 91, 48, 60, 43, 92, 70, 62, 61, 9, 10, 4, 81, 72, 27, 27, 75, 58], dtype=float).reshape(5, 5)
    A_ScS = np.array([26, 30, 4, 90, 91, 41, 32, 89, 34, 61, 27, 51, 80, 20, 86], dtype=float).reshape(3, 5)
    A_SSc = np.array([81, 97, 34, 3, 55, 68, 49, 49, 11, 24, 68, 50, 46, 40, 78], dtype=float).reshape(5, 3)
    A_ScSc = np.array([93, 53, 14, 74, 24, 73, 58, 63, 66], dtype=float).reshape(3, 3)
    Sc = 1 - S
    assert np.array_equal(O.select_submatrix(A, S, S), A_SS)
    assert np.array_equal(O.select_submatrix(A, Sc, S), A_ScS)
    assert np.array_equal(O.select_submatrix(A, S, Sc), A_SSc)
    assert np.array_equal(O.select_submatrix(A, Sc, Sc), A_ScSc)


def test_update_submatrix_utils_cc_425():
    S = np.array([1, 1, 0, 0, 1, 0, 1, 1])
    Sc = 1 - S
    A = np.array([90, 82, 36, 39, 57, 17, 23, 11, 96, 25, 84, 57, 47, 61, 92, 97, 55, 93, 59, 8, 2, 27, 16, 1, 14, 35, 55, 6,
                  34, 66, 83, 78, 15, 20, 92, 54, 17, 69, 54, 82, 26, 26, 29, 78, 80, 75, 100, 87, 85, 62, 76, 94, 32, 46, 8,
                  9, 26, 48, 76, 13, 53, 9, 45, 40], dtype=float).reshape(8, 8)
    m0 = np.arange(1, 26, dtype=float).reshape(5, 5)
    m1 = np.arange(1, 16, dtype=float).reshape(5, 3)
    m2 = np.arange(1, 16, dtype=float).reshape(3, 5)
    m3 = np.arange(1, 10, dtype=float).reshape(3, 3)
    A = O.update_submatrix(A, S, S, m0)
    res = np.array([1, 2, 36, 39, 3, 17, 4, 5, 6, 7, 84, 57, 8, 61, 9, 10, 55, 93, 59, 8, 2, 27, 16, 1, 14, 35, 55, 6, 34, 66,
                    83, 78, 11, 12, 92, 54, 13, 69, 14, 15, 26, 26, 29, 78, 80, 75, 100, 87, 16, 17, 76, 94, 18, 46, 19, 20,
                    21, 22, 76, 13, 23, 9, 24, 25], dtype=float).reshape(8, 8)
    assert np.array_equal(A, res)
    A = O.update_submatrix(A, S, Sc, m1)
    A = O.update_submatrix(A, Sc, S, m2)
    A = O.update_submatrix(A, Sc, Sc, m3)
    res = np.array([1, 2, 1, 2, 3, 3, 4, 5, 6, 7, 4, 5, 8, 6, 9, 10, 1, 2, 1, 2, 3, 3, 4, 5, 6, 7, 4, 5, 8, 6, 9, 10, 11, 12, 7,
                    8, 13, 9, 14, 15, 11, 12, 7, 8, 13, 9, 14, 15, 16, 17, 10, 11, 18, 12, 19, 20, 21, 22, 13, 14, 23, 15, 24,
                    25], dtype=float).reshape(8, 8)
    assert np.array_equal(A, res)


def test_select_update_subvector_utils_cc_464():
    v = np.array([79, 9, 93, 78, 49, 44, 45, 31, 51, 52, 82, 80, 65, 38, 82, 54, 36, 94, 88, 56], dtype=float)
    S = np.array([1, 1, 0, 0, 1, 0, 0, 0, 1, 1, 1, 0, 0, 1, 0, 0, 1, 0, 1, 1])
    assert np.array_equal(O.select_subvector(v, S), [79, 9, 49, 51, 52, 82, 38, 36, 88, 56])
    assert np.array_equal(O.select_subvector(v, 1 - S), [93, 78, 44, 45, 31, 80, 65, 82, 54, 94])
    n = np.arange(1, 11, dtype=float)
    v = O.update_subvector(v, S, n)
    assert np.array_equal(v, [1, 2, 93, 78, 3, 44, 45, 31, 4, 5, 6, 80, 65, 7, 82, 54, 8, 94, 9, 10])
    v = O.update_subvector(v, 1 - S, n)
    assert np.array_equal(v, [1, 2, 1, 2, 3, 3, 4, 5, 4, 5, 6, 6, 7, 7, 8, 9, 8, 10, 9, 10])
    v = O.update_subvector(v, S, 0.0)
    assert np.array_equal(v, [0, 0, 1, 2, 0, 3, 4, 5, 0, 0, 0, 6, 7, 0, 8, 9, 0, 10, 0, 0])


def test_cross_mat_utils_cc_329():
    rng = np.random.default_rng(0)
    for _ in range(10):
        a, b = rng.uniform(-1, 1, 3), rng.uniform(-1, 1, 3)
        assert np.allclose(O.cross_mat(a) @ b, np.cross(a, b), atol=1e-15)


def test_w_to_q_is_rotation_and_zero_w_is_identity():
    rng = np.random.default_rng(1)
    for _ in range(10):
        R = O.w_to_q_matrix(rng.uniform(-20, 20, 3), 0.01)
        assert np.allclose(R.T @ R, np.eye(3), atol=1e-9)        # IsOrthonormal, utils.cc:11-14
        assert abs(np.linalg.det(R) - 1) < 1e-9
    assert np.array_equal(O.w_to_q_matrix([0, 0, 0], 0.001), np.eye(3))    # utils.cc:83-85 comment


def test_align_vectors_utils_cc_499():
    rng = np.random.default_rng(2)
    for _ in range(100):
        a, b = rng.uniform(-1, 1, 3), rng.uniform(-1, 1, 3)
        R = O.align_vectors(a, b)
        assert (b / np.linalg.norm(b)) @ (R @ a) - np.linalg.norm(a) < 1e-9
        assert np.allclose(R @ a / np.linalg.norm(a), b / np.linalg.norm(b), atol=1e-9)
    # the contact frame of a +z normal is exactly the identity (ground contacts, contact.cc:50-51)
    assert np.array_equal(O.align_vectors([0, 0, 1], [0, 0, 1]), np.eye(3))
    # anti-parallel branch: our pinned rule still returns a proper rotation taking a to b
    R = O.align_vectors([0, 0, -1], [0, 0, 1])
    assert np.allclose(R @ [0, 0, -1], [0, 0, 1], atol=1e-12) and np.allclose(R.T @ R, np.eye(3), atol=1e-12)


def test_linear_algebra_against_numpy():
    rng = np.random.default_rng(3)
    for n in (1, 3, 7, 30):
        M = rng.uniform(-1, 1, (n, n))
        A = M.T @ M + 0.1 * np.eye(n)
        b = rng.uniform(-1, 1, n)
        assert np.allclose(O.ldlt_solve(A, b), np.linalg.solve(A, b), rtol=1e-9, atol=1e-11)
        assert np.allclose(O.lu_inverse(A), np.linalg.inv(A), rtol=1e-8, atol=1e-10)
        assert abs(O.condition_number(A) / np.linalg.cond(A) - 1) < 1e-6
    # semidefinite / indefinite systems: LDLT (not Cholesky) semantics
    A = np.diag([2.0, -1.0, 3.0])
    assert np.allclose(O.ldlt_solve(A, [2, 1, 3]), [1, -1, 1])
    assert O.condition_number(np.diag([1.0, 0.0])) == np.inf
