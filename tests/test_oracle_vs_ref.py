"""The CPU oracle against the REFERENCE'S OWN CODE: oracle/_ref/libkfbo_ref.so is built (oracle/Makefile) from
trial.cpp, simulator.cpp, kalman_filter.cpp, read_para_vehicle.cpp and include/*.hpp where they lie under
/root/reference, with oracle/ref_shim/ standing in for Eigen3 / ROS / matplotlibcpp (absent here) and a counter
for the wall clock the reference seeds its RNG from.  What this pins is the reference's own control flow and
operation order -- models, Van Loan discretisation, KF step, NEES/NIS, time variance, thread averages, and the
RNG discipline of eigenmvn.h; what it cannot pin is Eigen's internal arithmetic, which the shim restates.

Skipped where the reference tree / _ref is absent (the GPU box); tests/golden/ref_vectors.npz carries outputs of
the same library for those runs (tests/golden/make_ref_vectors.py)."""
import os

import numpy as np
import pytest

from conftest import ROOT, rel_err

from oracle import ref

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")

CASES = [(1.0, 0.1, 0.1), (0.3, 0.7, 0.5), (5.0, 0.01, 1.0), (0.1, 1.0, 0.1)]   # (pnoise, onoise, dt)


def test_models_match_the_reference_code(oracle):
    for q, r, dt in CASES:
        Fd, Qd, gu, H, R = ref.models(q, r, dt)
        oFd, oQd = oracle.discretize(q, dt)
        np.testing.assert_allclose(Fd, oFd, rtol=0, atol=0)                   # [[1, dt], [0, 1]] exactly
        np.testing.assert_allclose(Qd, oQd, rtol=4e-16)                       # two different expm evaluations, <= 2 ulp
        assert np.array_equal(gu, oracle.gu_table(dt))                        # accumulated j += dt, 2 cos(0.75 j), g*u: bit-exact
        assert np.array_equal(H, [1.0, 0.0]) and R == r / dt                  # system_models.hpp:141-143


def test_van_loan_discretisation_general_matrices(oracle):
    # continuousToDiscrete<Matrix2d> on the model's own (Fc, Qc) equals the oracle; on a non-nilpotent Fc it equals scipy
    from scipy.linalg import expm
    rng = np.random.default_rng(2)
    for _ in range(5):
        Fc = rng.normal(size=(2, 2)) * 0.5
        A = rng.normal(size=(2, 2))
        Qc = A @ A.T
        dt = float(rng.uniform(0.05, 1.0))
        Fd, Qd = ref.continuous_to_discrete(Fc, Qc, dt)
        big = np.block([[-Fc, Qc], [np.zeros((2, 2)), Fc.T]]) * dt
        E = expm(big)
        np.testing.assert_allclose(Fd, E[2:, 2:].T, rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(Qd, E[2:, 2:].T @ E[:2, 2:], rtol=1e-11, atol=1e-14)


def test_kalman_filter_step_matches_the_reference_code(oracle):
    # KalmanFilter::Step (kalman_filter.cpp:18-38) on the same measurement sequence: estimate, all four covariance
    # entries (never symmetrised), innovation and S, step by step
    rng = np.random.default_rng(3)
    for q, r, dt in CASES:
        for T in (1, 2, 201, 2000):
            z = np.cumsum(rng.normal(size=T)) * 0.3
            want, got = ref.kf_run(q, r, dt, z), oracle.kf_run(q, r, dt, z)
            assert rel_err(got[:, 2:6], want[:, 2:6]) < 1e-12                 # P
            assert rel_err(got[:, 7], want[:, 7]) < 1e-12                     # S
            np.testing.assert_allclose(got[:, :2], want[:, :2], rtol=1e-9, atol=1e-11)   # xest (sign changes: atol)
            np.testing.assert_allclose(got[:, 6], want[:, 6], rtol=1e-9, atol=1e-11)     # nu
            assert not np.array_equal(want[:, 3], want[:, 4]) or T < 3        # P01 != P10 in the reference too


@pytest.mark.parametrize("dt,t_end", [(0.1, 20.1), (1.0, 201.0), (0.5, 105.5)])
def test_trial_run_matches_the_reference_code(oracle, dt, t_end):
    # Trial::Run end to end (trial.cpp:16-113) under the counter clock: same seeds, same draws (eigenmvn.h's functor
    # copy discards half a polar pair per observation draw), same NEES/NIS, time variances and thread averages
    for q_est, r_est, nsim, cores, clock0 in [(1.0, 0.1, 1, 1, 1000), (0.1, 0.1, 8, 1, 77), (5.0, 1.0, 10, 2, 123456),
                                              (1.0, 0.1, 25, 10, 2**32 - 3)]:
        want, calls, T = ref.trial_run(dt, t_end, nsim, cores, 1.0, 0.1, q_est, r_est, clock0)
        assert T == int((t_end - 0.0) / dt)                                   # read_para_vehicle.cpp:145
        assert calls == 2 + nsim // cores                                     # 2 sampler seeds + one per trial
        got = oracle.trial_run_refclock(dt, T, 1.0, 0.1, q_est, r_est, nsim, cores, clock0)
        assert rel_err(got, want) < 1e-9, (dt, q_est, got, want)


def test_matched_filter_is_consistent_in_the_reference_code():
    # chi-square consistency straight from the reference's code: E[NEES] = 2, E[NIS] = 1 at matched parameters
    out, _, _ = ref.trial_run(0.1, 20.1, 400, 1, 1.0, 0.1, 1.0, 0.1, clock0=5)
    assert abs(out[0] - 2.0) < 0.1 and abs(out[1] - 1.0) < 0.03
