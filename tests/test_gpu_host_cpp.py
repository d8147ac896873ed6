"""GPU tests of the C++ host programs (kf_bayesopt_b200/host/): the reference-shaped callers
(test_trial, the "noise" -> "cost" node, the BO loop) running on the CUDA evaluator."""
import json
import os
import subprocess

import numpy as np
import pytest

from test_host_cpp import BIN, host_bins, yamls  # noqa: F401  (fixtures)

pytestmark = pytest.mark.gpu


def _run(exe, *args, stdin=None):
    out = subprocess.run([os.path.join(BIN, exe), *map(str, args)], input=stdin, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    return out.stdout


def test_test_trial_matched_filter_is_consistent(host_bins, yamls):
    # launch/robot1d_test_trail.launch:4-5: estimator == truth -> J_NEES ~ 0; both the
    # reference-shaped caller (CPUcoreNumber Trial objects on std::threads) and the fused launch
    for extra in ([], ["--fused"]):
        out = _run("kfbo_test_trial", "--vehicle", yamls[0], "--nsimruns", 4000, *extra)
        j = float(next(l for l in out.splitlines() if l.startswith("cost JNEES")).split()[-1])
        var = float(next(l for l in out.splitlines() if l.startswith("variance NEES")).split()[-1])
        assert j < 0.02, out                      # SE(avg NEES) ~ 0.4/sqrt(4000) -> J ~ 0.003
        assert 3.0 < var < 5.0, out               # anchor: avgVarNEES ~ 3.86 at dt = 0.1
    out = _run("kfbo_test_trial", "--vehicle", yamls[0], "--nsimruns", 4000, "--pnoise_est", 0.1, "--fused")
    j = float(next(l for l in out.splitlines() if l.startswith("cost JNEES")).split()[-1])
    assert abs(j - 1.613) < 0.05, out             # anchor: q_est = 0.1 -> J_NEES ~ 1.613 at dt = 0.1


def test_node_stdin_contract_matches_python_run_trial(host_bins, yamls, kfbo):
    seed = 424242
    msgs = [([0.7], [0.2]), ([1.0], [0.1]), ([3.0], [0.05])]
    lines = "".join(f"noise 1 1 {pn[0]} {on[0]}\n" for pn, on in msgs)
    out = _run("kfbo_robot1d_kf", "--vehicle", yamls[0], "--seed", seed, "--quiet", stdin=lines)
    got = [l.split() for l in out.splitlines() if l.startswith("cost ")]
    rv = kfbo.ReadParaVehicle.from_yaml(yamls[0])
    rt = kfbo.RunTrial(rv, seed=seed)
    assert len(got) == len(msgs)
    for g, (pn, on) in zip(got, msgs):
        cost = rt.ProcessNoiseCallback(kfbo.Float64MultiArray.noise(pn, on))
        assert float(g[1]) == cost.data                       # same Philox streams -> bit-identical cost
        assert int(g[2]) == rt.last_result.index_max
        assert [float(x) for x in g[3:]] == list(rt.last_result.cost)
    # the batched message: candidate c of a batch == evaluation with the same stream bookkeeping
    out = _run("kfbo_robot1d_kf", "--vehicle", yamls[0], "--seed", seed, "--quiet", "--max_candidates", 3,
               stdin="batch 3 0.7 0.2 1.0 0.1 3.0 0.05\n")
    got = [l.split() for l in out.splitlines() if l.startswith("cost ")]
    with kfbo.CostEvaluator(rv, kfbo.sample_times_from(rv), seed=seed, max_candidates=3) as ev:
        want = ev.eval_batch([0.7, 1.0, 3.0], [0.2, 0.1, 0.05])
    assert [float(g[1]) for g in got] == [w.max_cost for w in want]


def test_bo_loop_runs_the_topic_contract_end_to_end(host_bins, yamls):
    out = _run("kfbo_bo_loop", "--vehicle", yamls[0], "--bayes", yamls[1], "--iterations", 25, "--init", 10,
               "--nsimruns", 2000, "--quiet")
    line = json.loads(next(l for l in out.splitlines() if l.startswith("{")))
    assert line["iterations"] == 35 and line["nsimruns"] == 2000 and line["sample_times"] == 2
    assert line["best_cost"] < 0.05                           # a consistent filter was found
    assert 0.1 <= line["best_x"][0] <= 5.0 and 0.01 <= line["best_x"][1] <= 1.0
    assert line["dt0_count"] + line["dt1_count"] == 25        # trial_node.cpp:121-126 bookkeeping
    assert line["evaluator_s"] <= line["wall_s"]


def test_bo_loop_consumes_the_settings_of_bayes_params_yaml(host_bins, yamls, tmp_path):
    """The stand-in optimiser honours config/bayesParams.yaml the way the reference's node does
    (robot1d_bayesopt.cpp:242-249 copies n_init_samples, n_iterations, ...; :251-262 the bounds): the number of
    evaluations is n_init_samples + n_iterations FROM THE FILE, every sample lies inside the file's bounds, and
    optimizationChoice of vehicleParams.yaml decides which half of the "noise" message the sample fills."""
    import re
    bayes = tmp_path / "bayesParams.yaml"
    text = open(yamls[1]).read()
    text = re.sub(r"n_init_samples: *\d+", "n_init_samples: 7", text)
    text = re.sub(r"n_iterations: *\d+", "n_iterations: 9", text)
    text = re.sub(r"lower_bound: *\[[^\]]*\]", "lower_bound: [0.5, 0.05]", text)
    text = re.sub(r"upper_bound: *\[[^\]]*\]", "upper_bound: [2.0, 0.5]", text)
    bayes.write_text(text)
    out = _run("kfbo_bo_loop", "--vehicle", yamls[0], "--bayes", bayes, "--nsimruns", 1000)
    line = json.loads(next(l for l in out.splitlines() if l.startswith("{")))
    assert (line["n_init_samples"], line["n_iterations"], line["iterations"]) == (7, 9, 16)
    assert line["lower_bound"] == [0.5, 0.05] and line["upper_bound"] == [2.0, 0.5]
    xs = np.array([[float(v) for v in re.match(r"iteration \d+ x (\S+) (\S+) cost", l).groups()] for l in out.splitlines() if l.startswith("iteration ")])
    assert xs.shape == (16, 2)
    assert np.all(xs[:, 0] >= 0.5) and np.all(xs[:, 0] <= 2.0) and np.all(xs[:, 1] >= 0.05) and np.all(xs[:, 1] <= 0.5)
    assert line["dt0_count"] + line["dt1_count"] == 16 - 7    # trial_node.cpp:121-126 counts after the initial samples
    # optimizationChoice = processNoise: a 1-D search over pnoise inside the FIRST bound pair, onoise stays at the yaml value
    vehicle = tmp_path / "vehicleParams.yaml"
    vehicle.write_text(open(yamls[0]).read().replace("optimizationChoice: all", "optimizationChoice: processNoise"))
    out = _run("kfbo_bo_loop", "--vehicle", vehicle, "--bayes", bayes, "--nsimruns", 1000)
    line = json.loads(next(l for l in out.splitlines() if l.startswith("{")))
    assert line["optimization_choice"] == "processNoise" and len(line["best_x"]) == 1 and 0.5 <= line["best_x"][0] <= 2.0
    assert line["iterations"] == 16


def test_trace_csv_and_cost_surface_cli(host_bins, yamls, tmp_path):
    trace = tmp_path / "trace.csv"
    out = _run("kfbo_test_trial", "--vehicle", yamls[0], "--nsimruns", 10, "--fused", "--trace", trace)
    assert "trace of trajectory 0" in out
    t = np.loadtxt(trace, delimiter=",", skiprows=1)
    assert t.shape == (201, 9) and np.all(t[:, 4] > 0) and np.all(t[:, 7] >= 0) and np.all(t[:, 8] >= 0)
    np.testing.assert_allclose(t[:, 0], np.arange(201) * 0.1, rtol=1e-12)
    surf = tmp_path / "surface.csv"
    out = _run("kfbo_cost_surface", "--vehicle", yamls[0], "--bayes", yamls[1], "--grid", 12, "--nsimruns", 2000, "--out", surf)
    s = np.loadtxt(surf, delimiter=",", skiprows=1)
    assert s.shape == (144, 6)
    assert np.all(s[:, 4] == np.maximum(s[:, 2], s[:, 3]))        # max_cost = max over the sample times
    assert np.all((s[:, 5] == 0) | (s[:, 5] == 1))
    best = s[np.argmin(s[:, 4])]
    assert best[4] < 0.2 and 0.2 < best[0] < 5.0
