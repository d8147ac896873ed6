"""Property tests (hypothesis) of the host side of the path against the oracle: the sharding arithmetic, the
discretisation (row a1) and the cost assembly (row a9: trial.cpp:110-113 thread averages with the reference's integer
semantics, trial_node.cpp:70-119 cost formulas, first-maximum selection) on random shapes and values -- the fixed cases
live in tests/test_host_abi.py.  CPU only; the oracle is the checker."""
import ctypes as C

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from conftest import rel_err

SETTINGS = dict(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])


@settings(**SETTINGS)
@given(n=st.integers(0, 1 << 40), g=st.integers(1, 64))
def test_shard_range_is_a_balanced_partition(kfbo, n, g):
    spans = [kfbo.shard_range(n, r, g) for r in range(g)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    assert all(spans[i][1] == spans[i + 1][0] for i in range(g - 1))
    sizes = [e - b for b, e in spans]
    assert min(sizes) >= 0 and max(sizes) - min(sizes) <= 1
    assert sizes == sorted(sizes, reverse=True)                 # the remainder goes to the low ranks


@settings(**SETTINGS)
@given(q=st.floats(1e-3, 1e3), dt=st.sampled_from([0.01, 0.05, 0.1, 0.2, 0.25, 0.5, 1.0, 2.0]))
def test_discretisation_equals_the_oracle_and_the_closed_form(kfbo, oracle, q, dt):
    Fd, Qd = kfbo.discretize(q, dt)
    Fo, Qo = oracle.discretize(q, dt)
    assert np.array_equal(Fd, Fo) and rel_err(Qd, Qo) < 1e-14
    # Van Loan for Fc = [[0,1],[0,0]], Qc = diag(0, q) (continuous2discrete.hpp:8-32): the series ends after three terms
    assert np.array_equal(Fd, np.array([[1.0, dt], [0.0, 1.0]]))
    want = q * np.array([[dt**3 / 3, dt**2 / 2], [dt**2 / 2, dt]])
    assert rel_err(Qd, want) < 1e-14


@settings(**SETTINGS)
@given(data=st.data(), choice=st.sampled_from(["JNEES", "CNEES", "JNIS", "CNIS"]), cores=st.integers(1, 12),
       per_thread=st.integers(1, 9), extra=st.integers(0, 11), D=st.integers(1, 4))
def test_cost_assembly_equals_the_oracle_on_random_shapes(kfbo, oracle, data, choice, cores, per_thread, extra, D):
    # nsimruns need not divide by the thread count: the reference runs nsimruns / cores trials per thread (trial.cpp:19)
    # but averages with k_average = (T * nsimruns) / cores (trial.cpp:28) -- both truncating
    nsim = cores * per_thread + (extra % cores)
    run = cores * per_thread
    times = [(data.draw(st.sampled_from([0.1, 0.5, 1.0])), data.draw(st.integers(2, 2000))) for _ in range(D)]
    seed = data.draw(st.integers(0, 2**32 - 1))
    rng = np.random.default_rng(seed)
    rv = kfbo.ReadParaVehicle(nsimruns_=nsim, cpu_core_number_=cores, cost_choice_=choice)
    params = kfbo.make_params(rv, times)
    assert kfbo._lib.load().kfbo_effective_trials(C.byref(params)) == run
    # per-trial {row_nees, row_nis, var_nees, var_nis}: sums over T steps of chi-square-like values, and their time variances
    per_trial = [np.stack([rng.gamma(1.0, 2.0, run) * T, rng.gamma(0.5, 2.0, run) * T, rng.gamma(2.0, 2.0, run), rng.gamma(2.0, 1.0, run)])
                 for _, T in times]
    res = kfbo.finalize(params, np.stack([p.sum(axis=1) for p in per_trial]))
    costs = []
    for d, (_, T) in enumerate(times):
        want = oracle.cost_from_per_trial(per_trial[d], T, nsim, cores, cost_choice=choice)
        got = [res.average_nees[d], res.average_nis[d], res.average_var_nees[d], res.average_var_nis[d]]
        # the product averages the SUM over all trials, the reference the per-thread averages: the same number up to
        # the order of the additions
        assert rel_err(got, want[:4]) < 1e-12
        assert res.cost[d] == pytest.approx(want[4], rel=1e-10, abs=1e-12)
        assert res.cost[d] >= 0.0                               # |log|
        costs.append(res.cost[d])
    m, i = oracle.max_cost(costs)
    assert res.index_max == i and res.max_cost == m             # first maximum, bit-identical given the costs
