"""GPU parity tests of the STATE-AUGMENTED variants (kfbo_params.state_dim = 3, 4; kf_bayesopt_b200/csrc/kfbo_aug.cuh: one
trajectory per thread -- the default with on-device noise -- or one covariance row per lane of a 4-lane group,
KFBO_OPT_AUG_LAYOUT) against their oracle (oracle/kfbo_oracle_aug.c) through the C ABI.

The reference has no such variant (include/system_models.hpp:11-12 fixes N = 2): the oracle states the generalisation
and tests/test_oracle_aug.py pins it (scipy expm, closed forms, a numpy filter, N = 2 == the main oracle bit for bit).
Tolerance as everywhere: 1e-9 relative on identical noise draws, selection identical."""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu

RTOL = 1e-9
Q_TRUTH, R_TRUTH = 1.0, 0.1
TIMES = [(0.1, 201), (0.5, 211)]


def _noise(rng, oracle, N, dt, T, B):
    _, Qd = oracle.aug_discretize(N, Q_TRUTH, dt)
    L = np.linalg.cholesky(Qd)
    x0n = rng.standard_normal((N, B))
    w = np.ascontiguousarray(np.einsum("ij,tjb->tib", L, rng.standard_normal((T, N, B))))
    v = np.sqrt(R_TRUTH / dt) * rng.standard_normal((T, B))
    return x0n, w, v


@pytest.mark.parametrize("N", [3, 4])
def test_supplied_noise_matches_the_oracle(kfbo, oracle, N):
    rng = np.random.default_rng(100 + N)
    # whole and ragged CTAs (32 trajectories each), one trajectory, the minimum and maximum step counts.  The horizon T dt
    # is kept where the problem is well conditioned: a chain of N integrators driven by white noise wanders like t^(N-1/2),
    # and NEES is formed from e = x - xhat -- beyond |x| / |e| ~ 1e6 (N = 4: T dt > ~150) ANY two FP64 evaluations of
    # these formulas differ by more than 1e-9 (for N = 2 the same effect is 2e-12 at T dt = 2000)
    for B, T, dt in [(200, 201, 0.1), (33, 211, 0.5), (1, 2, 0.1), (31, 3, 1.0), (64, 2000, 0.1 if N == 3 else 0.05), (257, 64, 0.25)]:
        rv = kfbo.ReadParaVehicle(nsimruns_=B, cpu_core_number_=1, state_dof_=N)
        x0n, w, v = _noise(rng, oracle, N, dt, T, B)
        q, r = np.array([1.0, 0.3, 4.0]), np.array([0.1, 0.5, 0.02])
        with kfbo.CostEvaluator(rv, [(dt, T)], max_candidates=3, state_dim=N) as ev:
            pt, sums = ev.eval_supplied(0, q, r, x0n, w, v)
        for c in range(3):
            want = oracle.aug_eval_supplied(N, dt, T, q[c], r[c], x0n, w, v)
            assert rel_err(pt[c], want) < RTOL, (N, B, T, c)
            assert rel_err(sums[c], want.sum(axis=1)) < RTOL, (N, B, T, c)


LAYOUTS = ["thread", "lanes"]


def _set_layout(kfbo, ev, layout):
    from kf_bayesopt_b200 import _lib
    ev.set_option(_lib.OPT_AUG_LAYOUT, {"thread": _lib.AUG_LAYOUT_THREAD, "lanes": _lib.AUG_LAYOUT_LANES}[layout])


@pytest.mark.parametrize("layout", LAYOUTS)
@pytest.mark.parametrize("N", [3, 4])
def test_philox_trials_match_the_oracle_draw_for_draw(kfbo, oracle, N, layout):
    seed = 0x0123456789ABCDEF
    rv = kfbo.ReadParaVehicle(nsimruns_=100, cpu_core_number_=1, state_dof_=N)
    for times in (TIMES, [(0.1, 2), (0.5, 3)], [(0.1, 7), (0.05, 1999)]):      # even and odd step counts (the noise comes in step pairs)
        with kfbo.CostEvaluator(rv, times, seed=seed, state_dim=N) as ev:
            _set_layout(kfbo, ev, layout)
            for k, (dt, T) in enumerate(times):
                for qe, re, stream, trial0, n in [(1.0, 0.1, 0, 0, 100), (0.3, 0.7, (1 << 40) + 77, 0xFFFFFF00, 37)]:
                    got = ev.eval_philox_trials(k, qe, re, stream, trial0, n)
                    want = oracle.aug_eval_philox(N, seed, stream, dt, T, Q_TRUTH, R_TRUTH, qe, re, trial0, n)
                    assert rel_err(got, want) < RTOL, (N, times, k, qe)


@pytest.mark.parametrize("N", [3, 4])
def test_the_two_layouts_give_the_same_bits(kfbo, N):
    # one trajectory per thread performs lane i's arithmetic for every row i, on the same Philox words: the per-trial outputs
    # are bit-identical; the cost sums add them in another order (32 / 128 trajectories per CTA) and agree to rounding.
    # Ragged tiles of either kernel, odd and even step counts, the shortest trial (T = 2: one step pair).
    from kf_bayesopt_b200 import _lib
    seed = 0xFEEDFACE12345678
    for B, times in [(1000, TIMES), (33, [(0.1, 7), (1.0, 64)]), (129, [(0.25, 2)])]:
        rv = kfbo.ReadParaVehicle(nsimruns_=B, cpu_core_number_=1, state_dof_=N)
        q, r = np.array([1.0, 0.2, 3.0]), np.array([0.1, 0.9, 0.03])
        got = {}
        for layout in LAYOUTS:
            with kfbo.CostEvaluator(rv, times, seed=seed, max_candidates=len(q), state_dim=N) as ev:
                _set_layout(kfbo, ev, layout)
                ev.eval_batch(q, r)
                sums = ev.last_sums(len(q)).copy()
                pt = [ev.eval_philox_trials(k, 0.7, 0.3, 5 + k, 11, B) for k in range(len(times))]
                launches = ev.last_kernel_ms()[1]
            got[layout] = (sums, pt, launches)
        assert rel_err(got["thread"][0], got["lanes"][0]) < 1e-13, (N, B)
        for a, b in zip(got["thread"][1], got["lanes"][1]):
            assert np.array_equal(a, b), (N, B)
        assert np.all(np.isfinite(got["thread"][0]))


@pytest.mark.parametrize("layout", LAYOUTS)
@pytest.mark.parametrize("N", [3, 4])
def test_batch_costs_match_the_oracle_and_the_selection_is_identical(kfbo, oracle, N, layout):
    B, cores, seed = 96, 4, 7
    rv = kfbo.ReadParaVehicle(nsimruns_=B, cpu_core_number_=cores, state_dof_=N)
    q, r = np.array([1.0, 0.1, 5.0, 0.5]), np.array([0.1, 0.1, 1.0, 0.01])
    D = len(TIMES)
    with kfbo.CostEvaluator(rv, TIMES, seed=seed, max_candidates=len(q), state_dim=N) as ev:
        _set_layout(kfbo, ev, layout)
        res = ev.eval_batch(q, r)
        sums = ev.last_sums(len(q))
        ev.set_eval_counter(0)
        cost, mx, idx = ev.eval_batch_costs(q, r, with_index=True)
        for c in range(len(q)):
            want_cost = []
            for k, (dt, T) in enumerate(TIMES):
                pt = oracle.aug_eval_philox(N, seed, c * D + k, dt, T, Q_TRUTH, R_TRUTH, q[c], r[c], 0, B)
                assert rel_err(sums[c, k], pt.sum(axis=1)) < RTOL
                avg_nees = pt[0].sum() / (T * B // cores) / cores          # trial.cpp:19,112, trial_node.cpp:74
                want_cost.append(abs(np.log(avg_nees / N)))                  # J_NEES with stateDOF = N (trial_node.cpp:77)
            assert res[c].cost == pytest.approx(want_cost, rel=RTOL, abs=1e-12)
            assert cost[c] == pytest.approx(want_cost, rel=RTOL, abs=1e-12)
            assert res[c].index_max == int(np.argmax(want_cost)) == idx[c]
        # one candidate alone == the same candidate in a batch (same streams)
        ev.set_eval_counter(0)
        ev.set_stream_base(2 * D)
        one = ev.eval([q[2]], [r[2]])
        assert np.array_equal(one.cost, res[2].cost)


@pytest.mark.parametrize("N", [3, 4])
def test_matched_filter_is_consistent_at_size(kfbo, N):
    # E[NEES] = N, E[NIS] = 1 for estimator == truth (the matched-filter smoke case of launch/robot1d_test_trail.launch:4-5)
    B, T = 1 << 16, 201
    rv = kfbo.ReadParaVehicle(nsimruns_=B, cpu_core_number_=1, state_dof_=N)
    with kfbo.CostEvaluator(rv, [(0.1, T), (1.0, T)], seed=5, state_dim=N) as ev:
        res = ev.eval([Q_TRUTH], [R_TRUTH])
        again = (ev.set_eval_counter(0), ev.eval([Q_TRUTH], [R_TRUTH]))[1]
    assert np.all(np.abs(res.average_nees - N) < 0.02) and np.all(np.abs(res.average_nis - 1.0) < 0.01)
    assert res.max_cost < 0.01
    assert np.array_equal(res.cost, again.cost)                       # deterministic run to run


def test_unsupported_combinations_are_refused(kfbo):
    rv = kfbo.ReadParaVehicle(nsimruns_=8, cpu_core_number_=1, state_dof_=3)
    with kfbo.CostEvaluator(rv, [(0.1, 21)], state_dim=3) as ev:
        with pytest.raises(kfbo.KfboError):
            ev.trace_trial(0, 1.0, 0.1, 0, 0)                          # trial.cpp:69-78 plots two states
    with pytest.raises(kfbo.KfboError):
        kfbo.CostEvaluator(rv, [(0.1, 21)], state_dim=5)               # beyond one 4-lane group
