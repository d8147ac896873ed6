#!/usr/bin/env python
"""Generate tests/golden/ref_vectors.npz: outputs of THE REFERENCE'S OWN CODE (oracle/_ref/libkfbo_ref.so, built by
oracle/Makefile from /root/reference's trial.cpp, simulator.cpp, kalman_filter.cpp, read_para_vehicle.cpp and headers)
for fixed inputs, so that machines without the reference tree (the GPU box) can still check against it.

  kf_*     KalmanFilter::Step on seeded measurement sequences: {xest, P (4 entries), nu, S} per step
  models_* Fd, Qd, g*u table, R of ProcessModel / ObservationModel
  trial_*  Trial::Run with Nsimruns = CPUcoreNumber = 1 under the counter clock: the four getters, plus the noise
           draws of that trial (x0n, w, v) as replayed by the oracle -- the replay is what
           tests/test_oracle_vs_ref.py checks against the reference end to end, so feeding these draws to a
           supplied-noise evaluation must reproduce the reference's getters.
  node_*   THE NODE: trial_node.cpp's main() -> RunTrial::ProcessNoiseCallback on queued "noise" messages (CPUcoreNumber = 1
           so that the static generator is not raced; the reference's own vehicleParams.yaml / bayesParams.yaml otherwise):
           the "cost" messages it published for each costChoice, plus the noise draws of every trial of every pass.
  stat_*   statistical anchors FROM THE REFERENCE'S OWN RNG PATH (mt19937 + eigenmvn): Trial::Run on 10 independent
           blocks of 2000 trials per case -> mean and standard error over the blocks of the four getters; the on-device
           Philox path has to agree with these within the Monte Carlo standard error (north_star).
Run here (the reference tree must be present):  python tests/golden/make_ref_vectors.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as o   # noqa: E402
from oracle import ref           # noqa: E402

out = {}
rng = np.random.default_rng(20261019)
kf_cases = [(1.0, 0.1, 0.1, 201), (0.3, 0.7, 0.5, 211), (5.0, 0.01, 1.0, 201), (0.1, 1.0, 0.1, 2000)]
out["kf_cases"] = np.array(kf_cases)
for i, (q, r, dt, T) in enumerate(kf_cases):
    z = np.cumsum(rng.normal(size=T)) * 0.3
    out[f"kf_z_{i}"] = z
    out[f"kf_out_{i}"] = ref.kf_run(q, r, dt, z)
    Fd, Qd, gu, H, R = ref.models(q, r, dt)
    out[f"models_{i}"] = np.concatenate([Fd.ravel(), Qd.ravel(), H, [R]])
    out[f"models_gu_{i}"] = gu
trial_cases = [  # dt, t_end, q_truth, r_truth, q_est, r_est, clock0
    (0.1, 20.1, 1.0, 0.1, 1.0, 0.1, 1000), (0.1, 20.1, 1.0, 0.1, 0.1, 0.1, 1001), (0.1, 20.1, 1.0, 0.1, 5.0, 1.0, 1002),
    (1.0, 201.0, 1.0, 0.1, 1.0, 0.1, 2000), (1.0, 201.0, 1.0, 0.1, 2.5, 0.01, 2001),
    (0.5, 105.5, 1.0, 0.1, 1.0, 0.1, 3000), (0.5, 105.5, 1.0, 0.1, 0.37, 0.62, 3001), (0.1, 200.0, 1.0, 0.1, 1.3, 0.07, 4000)]
out["trial_cases"] = np.array(trial_cases, dtype=np.float64)
for i, (dt, t_end, qt, rt, qe, re, clock0) in enumerate(trial_cases):
    getters, calls, T = ref.trial_run(dt, t_end, 1, 1, qt, rt, qe, re, int(clock0))
    assert calls == 3
    replay = o.trial_run_refclock(dt, T, qt, rt, qe, re, 1, 1, int(clock0))
    assert np.max(np.abs(replay - getters) / np.abs(getters)) < 1e-9, (i, replay, getters)
    x0n, w, v = o.refclock_draws(dt, T, qt, rt, int(clock0), 0)
    out[f"trial_T_{i}"] = np.array(T)
    out[f"trial_getters_{i}"] = getters        # average_nees_, average_nis_, average_var_nees_, average_var_nis_
    out[f"trial_x0n_{i}"], out[f"trial_w_{i}"], out[f"trial_v_{i}"] = x0n, w, v

# ---- the node (trial_node.cpp): "noise" messages in, "cost" messages out
REF = "/root/reference"
NODE_NSIM, NODE_CLOCK0 = 6, 5000
node_msgs = [([0.3], [0.5]), ([2.0], [0.05]), ([1.0], [0.1])]
yaml_params = ref.load_yaml_params(f"{REF}/config/vehicleParams.yaml", f"{REF}/config/bayesParams.yaml")
out["node_msgs"] = np.array([[pn[0], on[0]] for pn, on in node_msgs])
out["node_meta"] = np.array([NODE_NSIM, NODE_CLOCK0, yaml_params["dt"], yaml_params["t_end"], 1.0, 201.0])   # + trial_node.cpp:105-108
qt, rt = yaml_params["pnoise"][0], yaml_params["onoise"][0]
passes = [(yaml_params["dt"], yaml_params["t_end"]), (1.0, 201.0)]            # (dt, t_end) of the callback's two passes
for choice in ("JNEES", "CNEES", "JNIS", "CNIS"):
    ref.set_params(dict(yaml_params, CPUcoreNumber=1, Nsimruns=NODE_NSIM, costChoice=choice))
    costs, calls = ref.node_run(node_msgs, clock0=NODE_CLOCK0)
    assert len(costs) == len(node_msgs) and calls == len(node_msgs) * 2 * (2 + NODE_NSIM)
    out[f"node_costs_{choice}"] = costs
clock = NODE_CLOCK0
for m, (pn, on) in enumerate(node_msgs):
    for k, (dt, t_end) in enumerate(passes):
        want, _, T = ref.trial_run(dt, t_end, NODE_NSIM, 1, qt, rt, pn[0], on[0], clock)
        g = o.trial_run_refclock(dt, T, qt, rt, pn[0], on[0], NODE_NSIM, 1, clock)
        assert np.max(np.abs(g - want) / np.abs(want)) < 1e-9
        draws = [o.refclock_draws(dt, T, qt, rt, clock, i) for i in range(NODE_NSIM)]
        out[f"node_x0n_{m}_{k}"] = np.stack([d[0] for d in draws], axis=1)            # [2][B]
        out[f"node_w_{m}_{k}"] = np.stack([d[1] for d in draws], axis=2)              # [T][2][B]
        out[f"node_v_{m}_{k}"] = np.stack([d[2] for d in draws], axis=1)              # [T][B]
        out[f"node_getters_{m}_{k}"] = want
        out[f"node_T_{m}_{k}"] = np.array(T)
        clock += 2 + NODE_NSIM

# ---- statistical anchors of the reference's own noise path
stat_cases = [  # dt, t_end, q_est, r_est      (truth q = 1.0, r = 0.1: config/vehicleParams.yaml)
    (0.1, 20.1, 1.0, 0.1), (0.1, 20.1, 0.1, 0.1), (0.1, 20.1, 5.0, 0.1), (0.1, 20.1, 1.0, 0.01), (0.1, 20.1, 1.0, 1.0),
    (1.0, 201.0, 1.0, 0.1), (1.0, 201.0, 0.1, 0.1), (0.5, 105.5, 1.0, 0.1), (0.5, 105.5, 5.0, 1.0)]
STAT_BLOCKS, STAT_BLOCK_TRIALS = 10, 2000
out["stat_cases"] = np.array(stat_cases)
out["stat_meta"] = np.array([STAT_BLOCKS, STAT_BLOCK_TRIALS])
for i, (dt, t_end, qe, re) in enumerate(stat_cases):
    blocks = np.stack([ref.trial_run(dt, t_end, STAT_BLOCK_TRIALS, 1, qt, rt, qe, re, 10_000_000 * (i + 1) + 100_000 * b)[0]
                       for b in range(STAT_BLOCKS)])
    out[f"stat_mean_{i}"] = blocks.mean(axis=0)
    out[f"stat_se_{i}"] = blocks.std(axis=0, ddof=1) / np.sqrt(STAT_BLOCKS)
    print("stat", stat_cases[i], out[f"stat_mean_{i}"], out[f"stat_se_{i}"])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_vectors.npz"), **out)
print("wrote ref_vectors.npz:", len(out), "arrays")
