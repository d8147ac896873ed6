"""The oracle of the STATE-AUGMENTED variants (oracle/kfbo_oracle_aug.c, N = 3, 4) against independent known answers.
The reference has no such variant (include/system_models.hpp:11-12 fixes N = 2), so these are the pins: scipy's matrix
exponential, the closed forms of the integrator chain, a numpy restatement of the filter, N = 2 against the main oracle
bit for bit, and chi-square consistency."""
import math

import numpy as np
import pytest
import scipy.linalg

from conftest import rel_err


def _van_loan(N, q, dt):
    Fc = np.diag(np.ones(N - 1), 1)
    Qc = np.zeros((N, N))
    Qc[-1, -1] = q
    A = np.zeros((2 * N, 2 * N))
    A[:N, :N], A[:N, N:], A[N:, N:] = -Fc * dt, Qc * dt, Fc.T * dt          # continuous2discrete.hpp:15-21
    B = scipy.linalg.expm(A)
    Fd = B[N:, N:].T
    return A, Fd, Fd @ B[:N, N:]


@pytest.mark.parametrize("N", [2, 3, 4])
@pytest.mark.parametrize("dt", [0.1, 0.5, 1.0])
def test_discretisation_against_scipy_expm_and_closed_form(oracle, N, dt):
    q = 0.7
    A, Fd_s, Qd_s = _van_loan(N, q, dt)
    np.testing.assert_allclose(oracle.aug_expm(A), scipy.linalg.expm(A), rtol=1e-13, atol=1e-300)
    Fd, Qd = oracle.aug_discretize(N, q, dt)
    np.testing.assert_allclose(Fd, Fd_s, rtol=1e-13, atol=1e-300)
    np.testing.assert_allclose(Qd, Qd_s, rtol=1e-13)
    f = math.factorial
    for i in range(N):
        for j in range(N):
            assert Fd[i, j] == pytest.approx(dt ** (j - i) / f(j - i) if j >= i else 0.0, rel=1e-14, abs=0)
            assert Qd[i, j] == pytest.approx(q * dt ** (2 * N - 1 - i - j) / (f(N - 1 - i) * f(N - 1 - j) * (2 * N - 1 - i - j)), rel=1e-13)
    np.testing.assert_allclose(oracle.aug_gvector(N, dt), [dt ** (N - i) / f(N - i) for i in range(N)], rtol=1e-15)
    L = oracle.aug_cholesky(Qd)
    np.testing.assert_allclose(L, np.linalg.cholesky(Qd), rtol=1e-10)


def test_order_two_is_the_main_oracle_bit_for_bit(oracle):
    rng = np.random.default_rng(3)
    dt, T, B = 0.1, 57, 21
    Fd, Qd = oracle.aug_discretize(2, 1.3, dt)
    Fd2, Qd2 = oracle.discretize(1.3, dt)
    assert np.array_equal(Fd, Fd2) and np.array_equal(Qd, Qd2)
    x0n, w, v = rng.standard_normal((2, B)), 0.3 * rng.standard_normal((T, 2, B)), rng.standard_normal((T, B))
    assert np.array_equal(oracle.aug_eval_supplied(2, dt, T, 0.8, 0.15, x0n, w, v), oracle.eval_supplied(dt, T, 0.8, 0.15, x0n, w, v))


def _numpy_filter(oracle, N, dt, T, q_est, r_est, x0n, w, v):
    """Textbook restatement with numpy linear algebra (np.linalg.solve for the NEES, not cofactors)."""
    Fd, Qd = oracle.aug_discretize(N, q_est, dt)
    g, u, R = oracle.aug_gvector(N, dt), oracle.control_table(dt, 2000), r_est / dt
    H = np.zeros((1, N))
    H[0, 0] = 1.0
    B = x0n.shape[1]
    out = np.empty((4, B))
    for b in range(B):
        x, xh, P = x0n[:, b].copy(), np.zeros(N), np.eye(N)
        nees, nis = np.empty(T), np.empty(T)
        for s in range(T):
            x = Fd @ x + g * u[s] + w[s, :, b]
            z = x[0] + v[s, b]
            xp, Pp = Fd @ xh + g * u[s], Fd @ P @ Fd.T + Qd
            S = Pp[0, 0] + R
            K = Pp[:, 0] / S
            nu = z - xp[0]
            xh, P = xp + K * nu, (np.eye(N) - np.outer(K, H[0])) @ Pp
            e = x - xh
            nees[s], nis[s] = e @ np.linalg.solve(P, e), nu * nu / S
        out[:, b] = nees.sum(), nis.sum(), nees.var(ddof=1), nis.var(ddof=1)
    return out


@pytest.mark.parametrize("N", [3, 4])
def test_filter_against_a_numpy_restatement(oracle, N):
    rng = np.random.default_rng(10 + N)
    for dt, T in [(0.1, 201), (0.5, 60)]:
        B = 6
        _, Qd = oracle.aug_discretize(N, 1.0, dt)
        L = np.linalg.cholesky(Qd)
        x0n = rng.standard_normal((N, B))
        w = np.ascontiguousarray(np.einsum("ij,tjb->tib", L, rng.standard_normal((T, N, B))))
        v = np.sqrt(0.1 / dt) * rng.standard_normal((T, B))
        got = oracle.aug_eval_supplied(N, dt, T, 0.6, 0.2, x0n, w, v)
        assert rel_err(got, _numpy_filter(oracle, N, dt, T, 0.6, 0.2, x0n, w, v)) < 1e-8   # np.linalg.solve vs cofactors on P ~ 1e-9 conditioned


@pytest.mark.parametrize("N", [3, 4])
def test_matched_filter_is_chi_square_consistent(oracle, N):
    # E[NEES] = N, E[NIS] = 1 when the estimator's noise levels are the truth's (the reference's matched-filter smoke case,
    # launch/robot1d_test_trail.launch:4-5, for the longer chain); on-device noise layout, 3000 trials
    T, B = 201, 3000
    pt = oracle.aug_eval_philox(N, 99, 5, 0.1, T, 1.0, 0.1, 1.0, 0.1, 0, B)
    nees, nis = pt[0] / T, pt[1] / T
    assert abs(nees.mean() - N) < 5 * nees.std() / np.sqrt(B)
    assert abs(nis.mean() - 1.0) < 5 * nis.std() / np.sqrt(B)
    # a mistuned filter is not: q_est ten times too small -> NEES well above N
    pt = oracle.aug_eval_philox(N, 99, 5, 0.1, T, 1.0, 0.1, 0.1, 0.1, 0, 500)
    assert (pt[0] / T).mean() > 1.5 * N
    # streams: 62-bit ids, the upper half matters
    a = oracle.aug_eval_philox(N, 99, 5, 0.1, 9, 1.0, 0.1, 1.0, 0.1, 0, 4)
    b = oracle.aug_eval_philox(N, 99, 5 + (1 << 32), 0.1, 9, 1.0, 0.1, 1.0, 0.1, 0, 4)
    assert not np.array_equal(a, b)
