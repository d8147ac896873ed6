"""Golden vectors produced by THE REFERENCE'S OWN CODE (tests/golden/ref_vectors.npz, written by
tests/golden/make_ref_vectors.py from oracle/_ref -- the reference's trial.cpp / simulator.cpp / kalman_filter.cpp
compiled where they lie): runs everywhere, including the GPU box where the reference tree is absent.
CPU: the oracle reproduces them.  GPU: the CUDA path, fed the same noise draws through the C ABI, reproduces
Trial::Run's getters."""
import os

import numpy as np
import pytest

from conftest import ROOT, rel_err

RTOL = 1e-9


@pytest.fixture(scope="module")
def refvec():
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_vectors.npz"))


def _trial_cases(refvec):
    for i, (dt, t_end, qt, rt, qe, re, clock0) in enumerate(refvec["trial_cases"]):
        yield i, dt, int(refvec[f"trial_T_{i}"]), qt, rt, qe, re, int(clock0)


def test_oracle_reproduces_reference_models_and_filter(oracle, refvec):
    for i, (q, r, dt, T) in enumerate(refvec["kf_cases"]):
        m = refvec[f"models_{i}"]
        Fd, Qd = oracle.discretize(q, dt)
        assert np.array_equal(Fd.ravel(), m[:4])
        np.testing.assert_allclose(Qd.ravel(), m[4:8], rtol=4e-16)
        assert np.array_equal(m[8:10], [1.0, 0.0]) and m[10] == r / dt
        assert np.array_equal(oracle.gu_table(dt), refvec[f"models_gu_{i}"])
        got, want = oracle.kf_run(q, r, dt, refvec[f"kf_z_{i}"]), refvec[f"kf_out_{i}"]
        assert rel_err(got[:, 2:6], want[:, 2:6]) < 1e-12 and rel_err(got[:, 7], want[:, 7]) < 1e-12
        np.testing.assert_allclose(got[:, [0, 1, 6]], want[:, [0, 1, 6]], rtol=RTOL, atol=1e-11)


def test_oracle_reproduces_reference_trials(oracle, refvec):
    for i, dt, T, qt, rt, qe, re, clock0 in _trial_cases(refvec):
        want = refvec[f"trial_getters_{i}"]
        assert rel_err(oracle.trial_run_refclock(dt, T, qt, rt, qe, re, 1, 1, clock0), want) < RTOL
        # the same trial through the supplied-noise entry point: the draws are the whole coupling
        x0n, w, v = refvec[f"trial_x0n_{i}"], refvec[f"trial_w_{i}"], refvec[f"trial_v_{i}"]
        assert np.array_equal(np.concatenate(oracle.refclock_draws(dt, T, qt, rt, clock0, 0)[1:2]), w)
        pt = oracle.eval_supplied(dt, T, qe, re, x0n.reshape(2, 1), w.reshape(T, 2, 1), v.reshape(T, 1))[:, 0]
        assert rel_err([pt[0] / T, pt[1] / T, pt[2], pt[3]], want) < RTOL      # trial.cpp:110-113 with one trial


@pytest.mark.gpu
def test_cuda_path_reproduces_reference_trials(kfbo, refvec):
    # Trial::Run of the reference's own code vs the CUDA kernels on identical noise draws, through the C ABI:
    # average_nees_ = row_nees / T, average_nis_ = row_nis / T, average_var_* = the time variances (one trial)
    from kf_bayesopt_b200 import _lib
    for i, dt, T, qt, rt, qe, re, clock0 in _trial_cases(refvec):
        want = refvec[f"trial_getters_{i}"]
        x0n, w, v = refvec[f"trial_x0n_{i}"], refvec[f"trial_w_{i}"], refvec[f"trial_v_{i}"]
        rv = kfbo.ReadParaVehicle(nsimruns_=2, cpu_core_number_=1, pnoise_=[qt], onoise_=[rt])
        # two identical columns: an even trial count takes the bulk-copy kernel, then the per-thread-load kernel
        X0, W, V = np.repeat(x0n.reshape(2, 1), 2, 1), np.repeat(w.reshape(T, 2, 1), 2, 2), np.repeat(v.reshape(T, 1), 2, 1)
        with kfbo.CostEvaluator(rv, [(dt, T)]) as ev:
            for variant in (_lib.SUPPLIED_AUTO, _lib.SUPPLIED_LDG):
                ev.set_option(_lib.OPT_SUPPLIED_KERNEL, variant)
                pt, sums = ev.eval_supplied(0, qe, re, X0, W, V)
                for col in range(2):
                    got = [pt[0, 0, col] / T, pt[0, 1, col] / T, pt[0, 2, col], pt[0, 3, col]]
                    assert rel_err(got, want) < RTOL, (i, variant, got, want)
                # and the host-side cost assembly on those sums: J_NEES of the reference's formula (trial_node.cpp:77)
                res = kfbo.finalize(kfbo.make_params(rv, [(dt, T)]), sums[0].reshape(1, 4))
                assert res.cost[0] == pytest.approx(abs(np.log(want[0] / 2)), rel=1e-8, abs=1e-10)


# ---------------------------------------------------------------------------------------------
# Statistical anchors from the reference's OWN noise path (mt19937 + eigenmvn.h; stat_* arrays): north_star's second
# correctness clause -- the counter-based Philox noise must agree with the reference within the Monte Carlo standard
# error.  5 sigma of the combined standard error (the reference's comes from 10 blocks, so it is itself +-25%).
# ---------------------------------------------------------------------------------------------
def _stat_cases(refvec):
    for i, (dt, t_end, qe, re) in enumerate(refvec["stat_cases"]):
        yield i, float(dt), int((float(t_end) - 0.0) / float(dt)), float(qe), float(re), refvec[f"stat_mean_{i}"], refvec[f"stat_se_{i}"]


def _check_against_anchor(per_trial, T, mean, se, what):
    B = per_trial.shape[1]
    got = np.array([per_trial[0].mean() / T, per_trial[1].mean() / T, per_trial[2].mean(), per_trial[3].mean()])
    own = np.array([per_trial[0].std(ddof=1) / T, per_trial[1].std(ddof=1) / T, per_trial[2].std(ddof=1), per_trial[3].std(ddof=1)]) / np.sqrt(B)
    z = (got - mean) / np.sqrt(se ** 2 + own ** 2)
    assert np.all(np.abs(z) < 5.0), (what, got, mean, z)


def test_oracle_philox_noise_agrees_statistically_with_the_reference_rng(oracle, refvec):
    for i, dt, T, qe, re, mean, se in _stat_cases(refvec):
        pt = oracle.eval_philox(20261019, i, dt, T, 1.0, 0.1, qe, re, 0, 20000)
        _check_against_anchor(pt, T, mean, se, ("oracle philox", i))


@pytest.mark.gpu
def test_cuda_philox_noise_agrees_statistically_with_the_reference_rng(kfbo, refvec):
    B = 1 << 17
    for i, dt, T, qe, re, mean, se in _stat_cases(refvec):
        rv = kfbo.ReadParaVehicle(nsimruns_=B, cpu_core_number_=1)
        with kfbo.CostEvaluator(rv, [(dt, T)], seed=20261019 + i) as ev:
            pt = ev.eval_philox_trials(0, qe, re, 0, 0, B)
            _check_against_anchor(pt, T, mean, se, ("cuda philox", i))
            # and the whole-evaluation path (in-kernel reduction -> finalize) gives the same averages as its own per-trial rows
            ev.set_eval_counter(0)
            res = ev.eval([qe], [re])
            assert res.average_nees[0] == pytest.approx(pt[0].sum() / (B * T), rel=1e-9)
            assert res.average_var_nis[0] == pytest.approx(pt[3].mean(), rel=1e-9)
