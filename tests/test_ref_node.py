"""THE REFERENCE'S NODE against the oracle and the product's host side.  oracle/_ref/libkfbo_ref.so also holds the
reference's trial_node.cpp, read_para_vehicle.cpp and read_para_bayes.cpp compiled where they lie (oracle/Makefile;
main() renamed, oracle/ref_shim/ros/ros.h standing in for the ROS topics and parameter server), so
RunTrial::ProcessNoiseCallback (trial_node.cpp:21-132) itself runs here: "noise" messages in, "cost" messages out.
That pins SURVEY.md 8 row a9 -- the two passes with their (dt, t_end) overwrite, the thread averages, the four cost
formulas, max_element and its index -- and the yaml front-end against the reference's own code.

The `ref`-marked tests need the reference tree (skipped on the GPU box); the golden-vector tests run everywhere
(tests/golden/ref_vectors.npz, node_* arrays, written by tests/golden/make_ref_vectors.py)."""
import os

import numpy as np
import pytest

from conftest import ROOT, rel_err

from oracle import ref

REF = "/root/reference"
needs_ref = pytest.mark.skipif(not (ref.available() and os.path.exists(f"{REF}/config/vehicleParams.yaml")),
                               reason="oracle/_ref not built (needs /root/reference)")
CHOICES = ("JNEES", "CNEES", "JNIS", "CNIS")


@pytest.fixture(scope="module")
def refvec():
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_vectors.npz"))


def _yaml_params(**over):
    return dict(ref.load_yaml_params(f"{REF}/config/vehicleParams.yaml", f"{REF}/config/bayesParams.yaml"), **over)


def _oracle_node_cost(oracle, getters_per_pass, choice):
    """trial_node.cpp:70-119 restated by the oracle: one thread's getters per pass -> (max cost, index)."""
    pic = [oracle.cost_from_threads(np.asarray(g, dtype=np.float64).reshape(1, 4), 2, 1, choice)[4] for g in getters_per_pass]
    return oracle.max_cost(pic)


# ------------------------------------------------------------------ yaml front-end
@needs_ref
def test_yaml_front_end_matches_the_reference_readers(kfbo):
    ref.set_params(_yaml_params())
    rv = kfbo.ReadParaVehicle.from_yaml(f"{REF}/config/vehicleParams.yaml")
    for show in (False, True):                      # both branches of read_para_vehicle.cpp:5-158 read the same values
        want = ref.read_para_vehicle(show)
        for f in ("cpu_core_number_", "state_dof_", "observation_dof_", "t_start_", "dt_", "t_end_", "xi0_", "xidot0_",
                  "nsimruns_", "max_iterations_", "dim_pn_", "dim_on_"):
            assert getattr(rv, f) == want[f], f
        assert rv.pnoise_[0] == want["pnoise0"] and rv.onoise_[0] == want["onoise0"]
        assert rv.optimization_choice_ == want["optimization_choice_"] and rv.cost_choice_ == want["cost_choice_"]
    rb = ref.read_para_bayes()                     # what the evaluator side uses of bayesParams.yaml (trial_node.cpp:122)
    assert rb["n_init_samples_"] == 20 and rb["lower_bound_"] == [0.1, 0.01] and rb["upper_bound_"] == [5.0, 1.0]


@needs_ref
def test_truncating_max_iterations_matches_the_reference_reader(kfbo):
    # read_para_vehicle.cpp:145: (t_end - t_start)/dt truncated -- including the cases where the quotient is 1 ulp off
    for dt, t_end in [(0.1, 20.1), (0.5, 105.5), (1.0, 201.0), (0.1, 100.0), (0.3, 60.0), (0.7, 7.0), (0.02, 4.02)]:
        ref.set_params(_yaml_params(dt=dt, t_end=t_end))
        want = ref.read_para_vehicle()["max_iterations_"]
        assert kfbo.ReadParaVehicle(dt_=dt, t_end_=t_end).max_iterations_ == want, (dt, t_end)


# ------------------------------------------------------------------ the node vs the oracle (reference tree present)
@needs_ref
@pytest.mark.parametrize("choice", CHOICES)
def test_node_cost_matches_the_oracle_replay(oracle, choice):
    nsim, clock0 = 5, 31337
    msgs = [([0.3], [0.5]), ([2.0], [0.05]), ([1.0], [0.1]), ([0.1], [1.0])]
    ref.set_params(_yaml_params(CPUcoreNumber=1, Nsimruns=nsim, costChoice=choice))
    costs, calls = ref.node_run(msgs, clock0=clock0)
    assert len(costs) == len(msgs)                                   # one "cost" per "noise"
    assert calls == len(msgs) * 2 * (2 + nsim)                       # per pass: 2 sampler seeds + one re-seed per trial
    clock = clock0
    for (pn, on), cost in zip(msgs, costs):
        getters = []
        for dt, T in [(0.1, 201), (1.0, 201)]:                       # yaml pass, then trial_node.cpp:105-114
            getters.append(oracle.trial_run_refclock(dt, T, 1.0, 0.1, pn[0], on[0], nsim, 1, clock))
            clock += 2 + nsim
        want, _ = _oracle_node_cost(oracle, getters, choice)
        assert cost == pytest.approx(want, rel=1e-9, abs=1e-12), (choice, pn, on)


@needs_ref
def test_node_with_threads_is_statistically_the_same(oracle):
    # CPUcoreNumber = 4 races on the process-wide static generator exactly as the reference does, so only the
    # statistics are comparable: the matched filter is consistent (cost -> 0), a mismatched one is not
    ref.set_params(_yaml_params(CPUcoreNumber=4, Nsimruns=400, costChoice="JNEES"))
    costs, _ = ref.node_run([([1.0], [0.1]), ([0.1], [0.1])], clock0=99)
    assert costs[0] < 0.08 and 1.3 < costs[1] < 1.9                 # SURVEY 8c anchors: J(0.1) ~ 1.6
    avg, T = ref.eval_threads(0.5, 105.5, 400, 4, 1.0, 0.1, 1.0, 0.1, clock0=7)
    assert T == 211 and abs(avg[0] - 2.0) < 0.1 and abs(avg[1] - 1.0) < 0.05


# ------------------------------------------------------------------ golden vectors of the node (run everywhere)
def _node_cases(refvec):
    nsim = int(refvec["node_meta"][0])
    for m, (q, r) in enumerate(refvec["node_msgs"]):
        passes = []
        for k in range(2):
            T = int(refvec[f"node_T_{m}_{k}"])
            dt = float(refvec["node_meta"][2 + 2 * k])
            passes.append((dt, T, refvec[f"node_x0n_{m}_{k}"], refvec[f"node_w_{m}_{k}"], refvec[f"node_v_{m}_{k}"],
                           refvec[f"node_getters_{m}_{k}"]))
        yield m, float(q), float(r), nsim, passes


@pytest.mark.parametrize("choice", CHOICES)
def test_oracle_reproduces_the_node_costs(oracle, refvec, choice):
    for m, q, r, nsim, passes in _node_cases(refvec):
        pic = []
        for dt, T, x0n, w, v, getters in passes:
            pt = oracle.eval_supplied(dt, T, q, r, x0n, w, v)
            assert rel_err(oracle.thread_averages(pt, 0, T, nsim, 1), getters) < 1e-9
            pic.append(oracle.cost_from_per_trial(pt, T, nsim, 1, cost_choice=choice)[4])
        want = refvec[f"node_costs_{choice}"][m]
        assert oracle.max_cost(pic)[0] == pytest.approx(want, rel=1e-9, abs=1e-12)


@pytest.mark.parametrize("choice", CHOICES)
def test_host_finalize_reproduces_the_node_costs(kfbo, oracle, refvec, choice):
    # the product's cost assembly (kfbo_finalize) on the sums of those trials: same cost, same winning pass
    for m, q, r, nsim, passes in _node_cases(refvec):
        rv = kfbo.ReadParaVehicle(nsimruns_=nsim, cpu_core_number_=1, cost_choice_=choice)
        sums = np.stack([oracle.eval_supplied(dt, T, q, r, x0n, w, v).sum(axis=1) for dt, T, x0n, w, v, _ in passes])
        res = kfbo.finalize(kfbo.make_params(rv, [(dt, T) for dt, T, *_ in passes]), sums)
        want = refvec[f"node_costs_{choice}"][m]
        assert res.max_cost == pytest.approx(want, rel=1e-9, abs=1e-12)
        assert res.index_max == int(np.argmax(res.cost)) and res.cost[res.index_max] == res.max_cost


@pytest.mark.gpu
@pytest.mark.parametrize("choice", CHOICES)
def test_cuda_path_reproduces_the_node_costs(kfbo, refvec, choice):
    # "noise" -> "cost" of the reference's node vs the CUDA kernels on the identical draws, through the C ABI:
    # both passes of the callback, the host cost assembly, the max over the sample times
    for m, q, r, nsim, passes in _node_cases(refvec):
        rv = kfbo.ReadParaVehicle(nsimruns_=nsim, cpu_core_number_=1, cost_choice_=choice)
        st = [(dt, T) for dt, T, *_ in passes]
        with kfbo.CostEvaluator(rv, st) as ev:
            sums = []
            for k, (dt, T, x0n, w, v, getters) in enumerate(passes):
                pt, s = ev.eval_supplied(k, q, r, x0n, w, v)
                got = [pt[0, 0].sum() / (nsim * T), pt[0, 1].sum() / (nsim * T), pt[0, 2].mean(), pt[0, 3].mean()]
                assert rel_err(got, getters) < 1e-9, (m, k, got, getters)
                sums.append(s[0])
            res = kfbo.finalize(kfbo.make_params(rv, st), np.stack(sums))
        want = refvec[f"node_costs_{choice}"][m]
        assert res.max_cost == pytest.approx(want, rel=1e-9, abs=1e-11), (choice, m)
