"""GPU tests of the reference-shaped host interface (Trial, RunTrial "noise" -> "cost")."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_trial_adapter_like_test_trial_cpp(kfbo):
    # src/test_trial.cpp:41-91 with launch/robot1d_test_trail.launch:4-5 (estimator == truth): J_NEES ~ 0
    rv_gt = kfbo.ReadParaVehicle(nsimruns_=4000, cpu_core_number_=10)
    rv = rv_gt.copy()
    rv.pnoise_[0], rv.onoise_[0] = 1.0, 0.1
    t = [kfbo.Trial(rv_gt, rv) for _ in range(rv_gt.cpu_core_number_)]
    for x in t:
        x.Run()
    average_nees = sum(x.get_average_nees() / rv_gt.cpu_core_number_ for x in t)
    average_nis = sum(x.get_average_nis() / rv_gt.cpu_core_number_ for x in t)
    average_var_nees = sum(x.get_average_var_nees() / rv_gt.cpu_core_number_ for x in t)
    assert abs(average_nees - 2.0) < 0.04 and abs(average_nis - 1.0) < 0.015
    assert abs(average_var_nees - 3.857) < 0.2
    assert abs(np.log(average_nees / rv.state_dof_)) < 0.02
    # every Trial drew its own noise
    assert len({x.get_average_nees() for x in t}) == len(t)


def test_run_trial_noise_to_cost_contract(kfbo):
    published = []
    rt = kfbo.RunTrial(kfbo.ReadParaVehicle(nsimruns_=2000), n_init_samples=1, publish=published.append)
    msg = kfbo.Float64MultiArray.noise([1.0], [0.1])
    assert msg.layout.dim[0].size == 1 and msg.layout.dim[1].size == 1 and msg.data == [1.0, 0.1]
    cost = rt.ProcessNoiseCallback(msg)
    assert isinstance(cost, kfbo.Float64) and published == [cost]
    assert 0 <= cost.data < 0.1                                   # matched filter, |log| cost
    assert cost.data == max(rt.last_result.cost)                  # max over the two sample times
    bad = rt.ProcessNoiseCallback(kfbo.Float64MultiArray.noise([0.1], [0.1]))
    assert bad.data > 1.0                                         # q_est = 0.1 -> J_NEES ~ 1.6
    assert rt.it_count_ == 2 and rt.dt0_count_ + rt.dt1_count_ == 1


def test_cost_choices(kfbo):
    rv = kfbo.ReadParaVehicle(nsimruns_=2000)
    out = {}
    for choice in ("JNEES", "CNEES", "JNIS", "CNIS"):
        with kfbo.CostEvaluator(rv, seed=5, cost_choice=choice) as ev:
            out[choice] = ev.eval([0.5], [0.3])
    a = out["JNEES"]
    for k in range(2):
        assert out["CNEES"].cost[k] == pytest.approx(a.cost[k] + abs(np.log(0.5 * a.average_var_nees[k] / 2)), rel=1e-12)
        assert out["JNIS"].cost[k] == pytest.approx(abs(np.log(a.average_nis[k] / 1)), rel=1e-12)
        assert out["CNIS"].cost[k] == pytest.approx(out["JNIS"].cost[k] + abs(np.log(0.5 * a.average_var_nis[k] / 1)), rel=1e-12)


@pytest.mark.gpu
def test_every_kernel_runs_on_small_ragged_shapes():
    # tools/all_kernels_small.py: Philox (one candidate, batch, per-trial), the state-augmented kernels, both supplied-noise kernels on whole + ragged tiles,
    # odd trial counts, the gain recursion and the trace, with finite outputs everywhere
    import os
    import subprocess
    import sys
    from conftest import ROOT
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "all_kernels_small.py")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "all_kernels_small ok" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
