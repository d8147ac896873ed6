"""GPU parity tests: the CUDA path, called through the C ABI (libkfbo.so), against the CPU oracle
on the same seeded inputs, against the committed golden vectors, and -- at BASELINE.json's full
sizes -- through size-independent properties.

Tolerances (north_star): per-trial NEES/NIS sums and J costs within 1e-9 RELATIVE in FP64 when both
sides see identical noise draws; the max-over-sample-times selection identical; on-device RNG
statistically within the Monte Carlo standard error."""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu

RTOL = 1e-9
Q_TRUTH, R_TRUTH = 1.0, 0.1
CODE_TIMES = [(0.1, 201), (1.0, 201)]       # trial_node.cpp:105-114
README_TIMES = [(0.1, 201), (0.5, 211)]     # README.md:96 / BASELINE.json configs[1]


def _noise(rng, oracle, dt, T, B):
    """cfg2 parity inputs: standard normals mapped by the Cholesky factor of the truth Qd / sqrt(R)."""
    _, Qd = oracle.discretize(Q_TRUTH, dt)
    L = np.linalg.cholesky(Qd)
    x0n = rng.standard_normal((2, B))
    w = np.einsum("ij,tjb->tib", L, rng.standard_normal((T, 2, B)))
    v = np.sqrt(R_TRUTH / dt) * rng.standard_normal((T, B))
    return x0n, np.ascontiguousarray(w), v


def test_supplied_noise_matches_golden_vectors(kfbo, golden):
    B, cores = int(golden["meta"][1]), int(golden["meta"][2])
    st = [(float(dt), int(T)) for dt, T in golden["sample_times"]]
    rv = kfbo.ReadParaVehicle(nsimruns_=B, cpu_core_number_=cores)
    q, r = golden["candidates"][:, 0], golden["candidates"][:, 1]
    with kfbo.CostEvaluator(rv, st, max_candidates=len(q)) as ev:
        for d in range(len(st)):
            pt, sums = ev.eval_supplied(d, q, r, golden[f"x0noise_{d}"], golden[f"w_{d}"], golden[f"v_{d}"])
            for c in range(len(q)):
                assert rel_err(pt[c], golden[f"per_trial_{d}_{c}"]) < RTOL, (d, c)
                assert rel_err(sums[c], golden[f"per_trial_{d}_{c}"].sum(axis=1)) < RTOL


@pytest.mark.parametrize("times", [CODE_TIMES, README_TIMES])
def test_supplied_noise_matches_oracle_cfg2(kfbo, oracle, times):
    # BASELINE parity case: B=4096, 2-D candidates, PCG64(20260101), J_NEES and J_NIS, argmax identical
    B, cores = 4096, 8
    rng = np.random.Generator(np.random.PCG64(20260101))
    cands = np.array([[1.0, 0.1], [0.1, 0.1], [5.0, 1.0], [2.5, 0.01], [0.37, 0.62]])
    rv = kfbo.ReadParaVehicle(nsimruns_=B, cpu_core_number_=cores)
    params = {c: kfbo.make_params(rv, times, cost_choice=c) for c in ("JNEES", "CNEES", "JNIS", "CNIS")}
    sums = np.zeros((len(cands), len(times), 4))
    want_pt = {}
    with kfbo.CostEvaluator(rv, times, max_candidates=len(cands)) as ev:
        for d, (dt, T) in enumerate(times):
            x0n, w, v = _noise(rng, oracle, dt, T, B)
            pt, s = ev.eval_supplied(d, cands[:, 0], cands[:, 1], x0n, w, v)
            sums[:, d] = s
            for c, (qe, re) in enumerate(cands):
                want = oracle.eval_supplied(dt, T, qe, re, x0n, w, v)
                want_pt[c, d] = want
                assert rel_err(pt[c], want) < RTOL, (d, c)        # per-trial NEES/NIS sums and time-variances
                assert rel_err(s[c], want.sum(axis=1)) < RTOL
    for choice, p in params.items():
        for c in range(len(cands)):
            got = kfbo.finalize(p, sums[c])
            want_cost = [oracle.cost_from_per_trial(want_pt[c, d], T, B, cores, cost_choice=choice)[4]
                         for d, (dt, T) in enumerate(times)]
            np.testing.assert_allclose(got.cost, want_cost, rtol=RTOL, atol=1e-12)
            m, i = oracle.max_cost(want_cost)
            assert got.index_max == i                              # selection identical
            assert got.max_cost == pytest.approx(m, rel=RTOL, abs=1e-12)


def test_supplied_noise_edge_shapes(kfbo, oracle):
    # ragged tile (B not a multiple of the CTA size), a single trial, the minimum and maximum step counts
    rng = np.random.default_rng(5)
    for B, T, dt in [(1, 2, 0.1), (33, 3, 0.5), (129, 7, 1.0), (200, 2000, 0.1), (257, 201, 0.1)]:
        rv = kfbo.ReadParaVehicle(nsimruns_=max(B, 1), cpu_core_number_=1)
        x0n, w, v = _noise(rng, oracle, dt, T, B)
        with kfbo.CostEvaluator(rv, [(dt, T)]) as ev:
            pt, s = ev.eval_supplied(0, 0.8, 0.15, x0n, w, v)
        want = oracle.eval_supplied(dt, T, 0.8, 0.15, x0n, w, v)
        assert rel_err(pt[0], want) < RTOL, (B, T)
        assert rel_err(s[0], want.sum(axis=1)) < RTOL


def test_philox_trials_match_oracle_draw_for_draw(kfbo, oracle):
    # the device RNG path restated by the oracle: same counters -> same normals -> same per-trial outputs
    seed = 0x0123456789ABCDEF
    rv = kfbo.ReadParaVehicle(nsimruns_=300, cpu_core_number_=1)
    # step counts of every residue mod 4 (the noise comes in groups of four steps), including less than one group
    # ... and both sides of the 32-step bound below which the kernel shifts its variance sums by the first sample
    for times in (CODE_TIMES, README_TIMES, [(0.1, 200), (0.5, 2)], [(0.1, 3), (0.5, 4), (1.0, 5), (0.2, 6)], [(0.1, 7), (0.5, 1999)],
                  [(0.1, 31), (0.5, 40)], [(0.1, 32), (0.5, 33)]):
        with kfbo.CostEvaluator(rv, times, seed=seed) as ev:
            for k, (dt, T) in enumerate(times):
                # stream ids are 62 bits wide (kfbo.h: KFBO_STREAM_BITS): the upper 30 bits ride in the counter word of the call index
                for qe, re, stream, trial0 in [(1.0, 0.1, 0, 0), (0.3, 0.7, 77, 1000), (4.0, 0.02, 0xFFFFFFFE, 0xFFFFFF00),
                                               (1.0, 0.1, (1 << 32) + 77, 0), (2.0, 0.3, (1 << 62) - 1, 5), (0.5, 0.05, kfbo_trial_base() + 3, 9)]:
                    n = 255 if trial0 else 300
                    got = ev.eval_philox_trials(k, qe, re, stream, trial0, n)
                    want = oracle.eval_philox(seed, stream, dt, T, Q_TRUTH, R_TRUTH, qe, re, trial0, n)
                    assert rel_err(got, want) < RTOL, (times, k, qe, stream)


def kfbo_trial_base():
    from kf_bayesopt_b200 import _lib
    return _lib.TRIAL_STREAM_BASE


def test_stream_id_upper_bits_select_other_streams(kfbo):
    rv = kfbo.ReadParaVehicle(nsimruns_=64, cpu_core_number_=1)
    with kfbo.CostEvaluator(rv, [(0.1, 21)], seed=3) as ev:
        a = ev.eval_philox_trials(0, 1.0, 0.1, 9, 0, 64)
        b = ev.eval_philox_trials(0, 1.0, 0.1, 9 + (1 << 32), 0, 64)
        assert not np.array_equal(a, b)
        # the evaluation counter no longer wraps the stream id at 2^32: evaluation 2^32 is not evaluation 0 again
        ev.set_eval_counter(0)
        s0 = (ev.eval([1.0], [0.1]), ev.last_sums(1).copy())[1]
        ev.set_eval_counter(1 << 32)
        s1 = (ev.eval([1.0], [0.1]), ev.last_sums(1).copy())[1]
        ev.set_eval_counter(0)
        s2 = (ev.eval([1.0], [0.1]), ev.last_sums(1).copy())[1]
        assert np.array_equal(s0, s2) and not np.array_equal(s0, s1)


def test_parity_envelope(kfbo, oracle):
    """The corners the reference can reach beyond the yaml bounds (tests/envelope_model.py: CASES).  Inside the envelope the
    CUDA path matches the oracle's reference arithmetic to 1e-9; where r_est -> 0 pushes the REFERENCE's (I - K H) P-
    past 1e-9 of its own exact value, the CUDA path (P = R K) is checked against the long-double evaluation instead.
    The measured errors are recorded in DESIGN.md section 6."""
    import envelope_model as em
    rng = np.random.default_rng(20261021)
    B = 48
    for name, dt, T, qt, rt, qe, re_, inside in em.CASES:
        x0n, w, v = em.noise(oracle, rng, dt, T, B, qt, rt)
        rv = kfbo.ReadParaVehicle(nsimruns_=B, cpu_core_number_=1, pnoise_=[qt], onoise_=[rt])
        with kfbo.CostEvaluator(rv, [(dt, T)]) as ev:
            pt, _ = ev.eval_supplied(0, qe, re_, x0n, w, v)
        want = oracle.eval_supplied(dt, T, qe, re_, x0n, w, v)
        exact = em.run(oracle, "device", dt, T, qe, re_, x0n, w, v, em.LD)   # P = R K in long double: no cancellation, the most accurate value at hand
        vs_oracle, vs_exact, oracle_vs_exact = rel_err(pt[0], want), em.rel(pt[0], exact), em.rel(want, exact)
        print(f"envelope {name}: cuda-vs-oracle {vs_oracle:.1e}, cuda-vs-long-double {vs_exact:.1e}, oracle-vs-long-double {oracle_vs_exact:.1e}")
        assert vs_exact < 1e-10, (name, vs_exact)                  # the CUDA path itself is at rounding level everywhere
        if inside:
            assert vs_oracle < RTOL, (name, vs_oracle)
        else:
            assert oracle_vs_exact > RTOL                          # the reference formulas are the inaccurate side here


def test_candidates_outside_the_domain_give_a_nan_cost_like_the_reference(kfbo):
    """Non-finite or <= 0 pnoise / onoise: the reference's filter degenerates and it publishes a NaN "cost"
    (trial_node.cpp:77,127-129), which ends the optimiser's wait loop (robot1d_bayesopt.cpp:107).  Same here: the call
    succeeds, that candidate's result is NaN with index_max 0, its neighbours in a batch are untouched."""
    rv = kfbo.ReadParaVehicle(nsimruns_=256, cpu_core_number_=1)
    with kfbo.CostEvaluator(rv, seed=11, max_candidates=6) as ev:
        for pn, on in [([0.0], [0.1]), ([-1.0], [0.1]), ([1.0], [0.0]), ([float("nan")], [0.1]), ([1.0], [float("inf")])]:
            r = ev.eval(pn, on)
            assert np.isnan(r.max_cost) and r.index_max == 0 and np.all(np.isnan(r.cost)), (pn, on)
            assert np.all(np.isnan(ev.last_sums(1)))
        ev.set_eval_counter(0)
        good = ev.eval_batch_raw([0.5, 1.0, 2.0, 3.0, 4.0, 5.0], [0.1] * 6)
        ev.set_eval_counter(0)
        mixed = ev.eval_batch_raw([0.5, -1.0, 2.0, float("nan"), 4.0, 5.0], [0.1, 0.1, 0.0, 0.1, 0.1, 0.1])
        for c in (0, 4, 5):
            assert np.array_equal(mixed["cost"][c], good["cost"][c]) and mixed["index_max"][c] == good["index_max"][c]
        for c in (1, 2, 3):
            assert np.isnan(mixed["max_cost"][c]) and mixed["index_max"][c] == 0
        with pytest.raises(kfbo.KfboError) as e:                   # positive but absurd: refused, not guessed at
            ev.eval([1e-200], [0.1])
        assert e.value.code == -2
        with pytest.raises(kfbo.KfboError):
            ev.eval_philox_trials(0, -1.0, 0.1, 0, 0, 10)          # the parity entry points have no "cost" message to fill
        assert ev.eval([1.0], [0.1]).max_cost >= 0                 # the handle stays usable


def test_reserved_scratch_makes_the_parity_entry_points_allocation_free(kfbo, oracle):
    import torch
    from kf_bayesopt_b200 import _lib
    rng = np.random.default_rng(2)
    B, T = 300, 201
    rv = kfbo.ReadParaVehicle(nsimruns_=B, cpu_core_number_=1)
    x0n, w, v = _noise(rng, oracle, 0.1, T, B)
    with kfbo.CostEvaluator(rv, [(0.1, T)], max_candidates=3) as ev:
        ev.set_option(_lib.OPT_RESERVE_SUPPLIED_TRIALS, B)
        torch.cuda.synchronize()
        free0 = torch.cuda.mem_get_info()[0]
        pt, _ = ev.eval_supplied(0, [0.5, 1.0, 2.0], [0.1, 0.1, 0.2], x0n, w, v)
        ev.eval_philox_trials(0, 1.0, 0.1, 0, 0, B)
        ev.trace_trial(0, 1.0, 0.1, 0, 0)
        assert torch.cuda.mem_get_info()[0] == free0               # nothing was allocated on the device by the calls
        assert rel_err(pt[1], oracle.eval_supplied(0.1, T, 1.0, 0.1, x0n, w, v)) < RTOL


def test_random_configurations_match_oracle(kfbo, oracle):
    # seeded sweep over what the fixed cases above hold constant: the TRUTH noise levels (pnoise_ / onoise_ of the
    # simulator's ReadParaVehicle), the sample time, the step count, a ragged trial count and the candidate
    rng = np.random.default_rng(20261020)
    for case in range(16):
        dt = float(rng.choice([0.05, 0.1, 0.2, 0.25, 0.5, 1.0]))
        T = int(rng.integers(2, 400))
        B = int(rng.integers(1, 300))
        q_truth, r_truth = float(np.exp(rng.uniform(np.log(0.05), np.log(5.0)))), float(np.exp(rng.uniform(np.log(0.01), np.log(1.0))))
        q_est, r_est = float(np.exp(rng.uniform(np.log(0.05), np.log(5.0)))), float(np.exp(rng.uniform(np.log(0.01), np.log(1.0))))
        stream, trial0 = int(rng.integers(0, 2**32 - 1)), int(rng.integers(0, 2**32 - 512))
        seed = int(rng.integers(0, 2**63))
        rv = kfbo.ReadParaVehicle(nsimruns_=B, cpu_core_number_=1, pnoise_=[q_truth], onoise_=[r_truth])
        with kfbo.CostEvaluator(rv, [(dt, T)], seed=seed) as ev:
            got = ev.eval_philox_trials(0, q_est, r_est, stream, trial0, B)
            want = oracle.eval_philox(seed, stream, dt, T, q_truth, r_truth, q_est, r_est, trial0, B)
            assert rel_err(got, want) < RTOL, (case, dt, T, B, q_truth, r_truth, q_est, r_est)
            # the same candidate against supplied draws of that truth model
            _, Qd = oracle.discretize(q_truth, dt)
            L = np.linalg.cholesky(Qd)
            x0n = rng.standard_normal((2, B))
            w = np.ascontiguousarray(np.einsum("ij,tjb->tib", L, rng.standard_normal((T, 2, B))))
            v = np.sqrt(r_truth / dt) * rng.standard_normal((T, B))
            pt, sm = ev.eval_supplied(0, q_est, r_est, x0n, w, v)
            want = oracle.eval_supplied(dt, T, q_est, r_est, x0n, w, v)
            assert rel_err(pt[0], want) < RTOL, (case, dt, T, B)
            assert rel_err(sm[0], want.sum(axis=1)) < RTOL


def test_eval_batch_sums_match_oracle_and_selection_is_identical(kfbo, oracle):
    B, cores, seed = 512, 8, 7
    rv = kfbo.ReadParaVehicle(nsimruns_=B, cpu_core_number_=cores)
    q = np.array([1.0, 0.1, 5.0, 0.5, 2.0, 1.3])
    r = np.array([0.1, 0.1, 1.0, 0.01, 0.3, 0.07])
    for times in (CODE_TIMES, README_TIMES):
        D = len(times)
        with kfbo.CostEvaluator(rv, times, seed=seed, max_candidates=len(q)) as ev:
            ev.set_stream_base(1000)
            res = ev.eval_batch(q, r)
            sums = ev.last_sums(len(q))
            for c in range(len(q)):
                want_cost = []
                for k, (dt, T) in enumerate(times):
                    stream = 1000 + (0 * len(q) + c) * D + k          # include/kfbo.h stream layout, evaluation 0
                    pt = oracle.eval_philox(seed, stream, dt, T, Q_TRUTH, R_TRUTH, q[c], r[c], 0, B)
                    assert rel_err(sums[c, k], pt.sum(axis=1)) < RTOL
                    out = oracle.cost_from_per_trial(pt, T, B, cores)
                    assert rel_err([res[c].average_nees[k], res[c].average_nis[k], res[c].average_var_nees[k],
                                    res[c].average_var_nis[k]], out[:4]) < RTOL
                    assert res[c].cost[k] == pytest.approx(out[4], rel=RTOL, abs=1e-12)
                    want_cost.append(out[4])
                m, i = oracle.max_cost(want_cost)
                assert res[c].index_max == i
            # the second evaluation draws fresh streams (evaluation counter advanced)
            res2 = ev.eval_batch(q, r)
            assert res2[0].average_nees[0] != res[0].average_nees[0]
            # ... and is reproducible when the counter is rewound
            ev.set_eval_counter(0)
            res3 = ev.eval_batch(q, r)
            assert all(np.array_equal(a.cost, b.cost) for a, b in zip(res, res3))


def test_device_built_cases_are_bit_identical_to_host_built(kfbo):
    """A batch of more than one candidate sends (pnoise, onoise) to the device and build_cases_kernel makes Fd, Qd, R and the
    Philox stream ids there; one candidate is built on the host (kfbo_api.cu: build_case).  Same arithmetic operation for
    operation (kfbo_host.h: RoundedOps), so candidate c of a batch and a single evaluation on the same streams are the
    same bits -- over awkward values of pnoise / dt too."""
    rng = np.random.default_rng(9)
    q = np.concatenate([[1.0, 0.1, 5.0, 1.0 / 3.0, 0.7], np.exp(rng.uniform(np.log(1e-3), np.log(1e3), 11))])
    r = np.concatenate([[0.1, 0.1, 1.0, 0.01, 0.3], np.exp(rng.uniform(np.log(1e-4), np.log(10.0), 11))])
    times = [(0.1, 201), (0.5, 211), (1.0 / 3.0, 37), (0.07, 64)]
    C, D = len(q), len(times)
    rv = kfbo.ReadParaVehicle(nsimruns_=300, cpu_core_number_=1)
    with kfbo.CostEvaluator(rv, times, seed=5, max_candidates=C) as ev:
        ev.set_stream_base((1 << 40) + 17)                         # the device forms the 62-bit stream ids too
        batch = ev.eval_batch_raw(q, r)
        sums = ev.last_sums(C)
        assert ev.last_transfer_bytes() == (16 * C, 32 * D * C)   # 16 bytes per candidate in, the [C][D][4] sums out
        for c in range(C):
            ev.set_stream_base((1 << 40) + 17 + c * D)            # candidate c's streams as candidate 0 of a single evaluation
            ev.set_eval_counter(0)
            one = ev.eval([q[c]], [r[c]])
            assert np.array_equal(ev.last_sums(1)[0], sums[c]), c
            assert np.array_equal(one.cost, batch["cost"][c, :D]) and one.index_max == batch["index_max"][c]


def test_device_side_cost_assembly_matches_the_host(kfbo):
    """kfbo_eval_batch_costs: trial_node.cpp:70-119 per candidate on the device.  Same divisions in the same order as
    kfbo_finalize; log() is CUDA's, within 1 ulp of libm's -- costs to 1e-9 (observed ~1e-16), selection identical."""
    rng = np.random.default_rng(4)
    C = 700
    q, r = np.exp(rng.uniform(np.log(0.1), np.log(5.0), C)), np.exp(rng.uniform(np.log(0.01), np.log(1.0), C))
    rv = kfbo.ReadParaVehicle(nsimruns_=130, cpu_core_number_=10)
    for choice in ("JNEES", "CNEES", "JNIS", "CNIS"):
        with kfbo.CostEvaluator(rv, README_TIMES, seed=21, max_candidates=C, cost_choice=choice) as ev:
            full = ev.eval_batch_raw(q, r)
            ev.set_eval_counter(0)
            cost, mx, idx = ev.eval_batch_costs(q, r, with_index=True)
            h2d, d2h = ev.last_transfer_bytes()
            assert (h2d, d2h) == (16 * C, C * (8 * 3 + 4))
            assert rel_err(cost, full["cost"][:, :2]) < RTOL and rel_err(mx, full["max_cost"]) < RTOL, choice
            assert np.array_equal(idx, full["index_max"]), choice
            # a candidate outside the domain in the middle of the batch: NaN there, its neighbours untouched
            q2 = q.copy()
            q2[5] = -1.0
            ev.set_eval_counter(0)
            cost2, mx2, idx2 = ev.eval_batch_costs(q2, r, with_index=True)
            assert np.isnan(mx2[5]) and np.all(np.isnan(cost2[5])) and idx2[5] == 0
            keep = np.arange(C) != 5
            assert np.array_equal(cost2[keep], cost[keep]) and np.array_equal(idx2[keep], idx[keep])


def test_graph_replay_gives_the_same_bits_as_call_by_call_enqueueing(kfbo):
    """(On one GPU the library enqueues call by call whatever the option says -- the replay only pays with the exchange kernel
    behind the rollout, tools/multi_gpu_check.py covers that -- so here this checks the option plumbing and the lazily read times.)
    A single-candidate evaluation is replayed from a CUDA graph (constants copy + rollout + timing events in one launch;
    only the by-value estimator constants of the kernel node change between replays).  Same kernels, same arguments:
    the results are the same bits as with KFBO_OPT_CUDA_GRAPH = 0 -- across changing candidates, a batch in between,
    a change of the CUDA stream (the graph is rebuilt) and the lazily read kernel time."""
    import torch
    from kf_bayesopt_b200 import _lib
    rv = kfbo.ReadParaVehicle(nsimruns_=3000, cpu_core_number_=10)
    cands = [(0.7, 0.2), (1.0, 0.1), (3.0, 0.05), (0.7, 0.2), (0.2, 0.9)]
    with kfbo.CostEvaluator(rv, seed=3, max_candidates=4) as a, kfbo.CostEvaluator(rv, seed=3, max_candidates=4) as b:
        b.set_option(_lib.OPT_CUDA_GRAPH, 0)
        stream = torch.cuda.Stream()
        for i, (q, r) in enumerate(cands):
            if i == 2:        # a batch between two replays ...
                assert np.array_equal(a.eval_batch_raw([0.5, 1.5], [0.1, 0.2])["cost"], b.eval_batch_raw([0.5, 1.5], [0.1, 0.2])["cost"])
            if i == 3:        # ... and another stream: the graph is captured anew on it
                a.set_cuda_stream(stream.cuda_stream)
                b.set_cuda_stream(stream.cuda_stream)
            ra, rb = a.eval([q], [r]), b.eval([q], [r])
            assert np.array_equal(a.last_sums(1), b.last_sums(1)), i
            assert np.array_equal(ra.cost, rb.cost) and ra.max_cost == rb.max_cost and ra.index_max == rb.index_max
            (ms_a, n_a), (ms_b, n_b) = a.last_kernel_ms(), b.last_kernel_ms()
            assert n_a == n_b == 1 and 0 < ms_a < 50 and 0 < ms_b < 50
        a.kernel_ms_total(reset=True)
        for q, r in cands:
            a.eval([q], [r])
        total, n = a.kernel_ms_total()
        assert n == len(cands) and 0 < total < 50 * n


def test_latency_form_of_the_rollout_gives_the_same_bits(kfbo, oracle):
    """rollout_philox_small (KFBO_OPT_SMALL_LAUNCH; launches of at most 2 CTAs per SM, e.g. the reference's default 200 x 201 x 2
    workload): 32 trajectories per CTA, one warp runs the filter, three make its noise a step group ahead.  Same streams, same
    filter_step, partial sums combined four at a time like the warps of a rollout_philox CTA: the sums are THE SAME BITS as the
    throughput kernel's -- whole and ragged tiles, one trajectory, every T mod 4, both variance-shift policies, batches --
    and the per-trial outputs still match the oracle draw for draw."""
    from kf_bayesopt_b200 import _lib
    shapes = [(200, [(0.1, 201), (1.0, 201)]), (31, [(0.1, 2), (0.5, 3)]), (33, [(0.1, 7), (0.5, 31)]), (1, [(0.1, 32), (0.5, 33)]),
              (300, [(0.1, 2000)]), (129, [(0.1, 201), (0.5, 211), (1.0, 64), (0.25, 5)])]
    for B, times in shapes:
        rv = kfbo.ReadParaVehicle(nsimruns_=B, cpu_core_number_=1)
        with kfbo.CostEvaluator(rv, times, seed=23, max_candidates=3) as a, kfbo.CostEvaluator(rv, times, seed=23, max_candidates=3) as b:
            b.set_option(_lib.OPT_SMALL_LAUNCH, 0)
            for q, r in [(1.0, 0.1), (0.4, 0.3)]:
                ra, rb = a.eval([q], [r]), b.eval([q], [r])
                assert np.array_equal(a.last_sums(1), b.last_sums(1)), (B, times, q)
                assert np.array_equal(ra.cost, rb.cost) and ra.index_max == rb.index_max
            ba, bb = a.eval_batch_raw([0.5, 1.0, 2.0], [0.1, 0.2, 0.05]), b.eval_batch_raw([0.5, 1.0, 2.0], [0.1, 0.2, 0.05])
            assert np.array_equal(a.last_sums(3), b.last_sums(3)) and np.array_equal(ba["cost"], bb["cost"])
            for k, (dt, T) in enumerate(times):
                pa, pb = a.eval_philox_trials(k, 0.7, 0.2, 5, 3, B), b.eval_philox_trials(k, 0.7, 0.2, 5, 3, B)
                assert np.array_equal(pa, pb), (B, k)
                assert rel_err(pa, oracle.eval_philox(23, 5, dt, T, Q_TRUTH, R_TRUTH, 0.7, 0.2, 3, B)) < RTOL


def test_eval_single_candidate_equals_batch_of_one(kfbo):
    rv = kfbo.ReadParaVehicle()                                   # yaml defaults: B=200, 10 "cores"
    with kfbo.CostEvaluator(rv, seed=11) as a, kfbo.CostEvaluator(rv, seed=11, max_candidates=1) as b:
        ra = a.eval([0.7], [0.2])
        rb = b.eval_batch([0.7], [0.2])[0]
        assert np.array_equal(ra.cost, rb.cost) and ra.index_max == rb.index_max
        assert ra.max_cost == max(ra.cost)
        ms, launches = a.last_kernel_ms()
        assert launches == 1 and ms > 0


def test_reference_integer_semantics_on_device(kfbo, oracle):
    # nsimruns % cores != 0: (nsimruns/cores)*cores trials run (trial.cpp:28), k_average = (T*nsimruns)/cores (:19)
    B, cores, seed = 203, 10, 3
    rv = kfbo.ReadParaVehicle(nsimruns_=B, cpu_core_number_=cores)
    with kfbo.CostEvaluator(rv, CODE_TIMES, seed=seed) as ev:
        assert ev.effective_trials == 200
        res = ev.eval([1.0], [0.1])
        for k, (dt, T) in enumerate(CODE_TIMES):
            pt = oracle.eval_philox(seed, k, dt, T, Q_TRUTH, R_TRUTH, 1.0, 0.1, 0, 200)
            out = oracle.cost_from_per_trial(pt, T, B, cores)
            assert res.cost[k] == pytest.approx(out[4], rel=RTOL, abs=1e-12)


def test_errors_are_reported_not_thrown(kfbo):
    with kfbo.CostEvaluator(kfbo.ReadParaVehicle(), max_candidates=2) as ev:
        with pytest.raises(kfbo.KfboError) as e:
            ev.eval_batch([1.0, 2.0, 3.0], [0.1, 0.1, 0.1])       # more than max_candidates
        assert e.value.code == -1
        with pytest.raises(kfbo.KfboError):
            ev.eval([], [0.1])                                    # pnoise_[0] would be out of range
        with pytest.raises(kfbo.KfboError):
            ev.eval_philox_trials(5, 1.0, 0.1, 0, 0, 10)          # sample time out of range
        assert ev.eval([1.0], [0.1]).max_cost >= 0                # handle still usable afterwards


# ---------------------------------------------------------------------------------------------
# full-size properties (BASELINE.json configs[2] and configs[3] shapes)
# ---------------------------------------------------------------------------------------------
def test_full_size_cfg3_properties(kfbo):
    B, T = 1 << 20, 1000
    rv = kfbo.ReadParaVehicle(nsimruns_=B, cpu_core_number_=1)
    times = [(0.1, T), (0.5, T)]
    with kfbo.CostEvaluator(rv, times, seed=2026) as ev:
        res = ev.eval([1.0], [0.1])
        s1 = ev.last_sums(1)
        # matched filter: E[NEES] = stateDOF, E[NIS] = observationDOF (chi-square consistency, SURVEY 8c anchors)
        # SE(avg NEES) = per-trial std of the time-mean / sqrt(B) < 0.4/1024
        assert abs(res.average_nees[0] - 2.0) < 5 * 0.4 / 1024
        assert abs(res.average_nees[1] - 2.0) < 5 * 0.4 / 1024
        assert abs(res.average_nis[0] - 1.0) < 5 * 0.1 / 1024
        assert res.cost[0] < 2e-3 and res.cost[1] < 2e-3
        # determinism: same streams -> bit-identical sums
        ev.set_eval_counter(0)
        ev.eval([1.0], [0.1])
        assert np.array_equal(s1, ev.last_sums(1))
        # checksum of checksums: per-trial outputs of a slice add up to that slice's contribution;
        # and a mistuned filter moves the cost the right way (NEES > dof when q_est is too small)
        res_small_q = ev.eval([0.1], [0.1])
        assert res_small_q.average_nees[0] > 5 and res_small_q.max_cost > 1.0


def test_trial_shards_add_up(kfbo):
    # evaluating the two halves of the trial range separately and adding equals the whole: the reduction
    # is linear and the Philox counters use the global trial index (what SHARD_TRIALS relies on)
    B, T, seed = 100_000, 201, 5
    rv = kfbo.ReadParaVehicle(nsimruns_=B, cpu_core_number_=1)
    with kfbo.CostEvaluator(rv, [(0.1, T)], seed=seed) as ev:
        ev.eval([0.6], [0.2])
        whole = ev.last_sums(1)[0, 0]
        lo = ev.eval_philox_trials(0, 0.6, 0.2, 0, 0, B // 2)
        hi = ev.eval_philox_trials(0, 0.6, 0.2, 0, B // 2, B - B // 2)
        np.testing.assert_allclose(lo.sum(axis=1) + hi.sum(axis=1), whole, rtol=1e-12)


def test_full_size_cfg4_sweep_shape(kfbo, oracle):
    # 4096 candidates x 1024 trials x 2 sample times; spot-check candidates against the oracle
    n, B, seed = 64, 1024, 9
    q = np.repeat(np.geomspace(0.1, 5.0, n), n)
    r = np.tile(np.geomspace(0.01, 1.0, n), n)
    rv = kfbo.ReadParaVehicle(nsimruns_=B, cpu_core_number_=1)
    with kfbo.CostEvaluator(rv, README_TIMES, seed=seed, max_candidates=n * n) as ev:
        full = ev.eval_batch_raw(q, r)
        sums = ev.last_sums(n * n)
        ev.set_eval_counter(0)
        cost, mx, idx = ev.eval_batch_costs(q, r, with_index=True)      # the same sweep, costs assembled on the device
    assert cost.shape == (n * n, 2) and np.all(np.isfinite(cost)) and np.all(mx == cost.max(axis=1))
    assert rel_err(cost, full["cost"][:, :2]) < RTOL and np.array_equal(idx, full["index_max"])
    for c in (0, 777, 2049, 4095):
        for k, (dt, T) in enumerate(README_TIMES):
            pt = oracle.eval_philox(seed, c * 2 + k, dt, T, Q_TRUTH, R_TRUTH, q[c], r[c], 0, B)
            assert rel_err(sums[c, k], pt.sum(axis=1)) < RTOL
    # the surface has its valley near the truth (q=1, r=0.1)
    best = np.argmin(mx)
    assert 0.3 < q[best] < 3.0 and 0.03 < r[best] < 0.4


def test_supplied_device_pointers_and_linearity(kfbo, oracle):
    import torch
    B, T, dt = 8192, 201, 0.1
    rng = np.random.default_rng(8)
    x0n, w, v = _noise(rng, oracle, dt, T, B)
    rv = kfbo.ReadParaVehicle(nsimruns_=B, cpu_core_number_=1)
    dev = torch.device("cuda:0")
    tx, tw, tv = (torch.from_numpy(a).to(dev) for a in (x0n, w, v))
    out = torch.empty((1, 4, B), dtype=torch.float64, device=dev)
    with kfbo.CostEvaluator(rv, [(dt, T)]) as ev:
        torch.cuda.synchronize()
        s = ev.eval_supplied_dev(0, 0.9, 0.12, B, tx.data_ptr(), tw.data_ptr(), tv.data_ptr(), out.data_ptr())
        pt_host, s_host = ev.eval_supplied(0, 0.9, 0.12, x0n, w, v)
    assert np.array_equal(out.cpu().numpy(), pt_host) and np.array_equal(s, s_host)
    assert rel_err(out.cpu().numpy()[0], oracle.eval_supplied(dt, T, 0.9, 0.12, x0n, w, v)) < RTOL


def test_supplied_kernels_ldg_and_bulk_agree_bit_for_bit(kfbo, oracle):
    # the bulk-copy (cp.async.bulk + mbarrier) kernel and the per-thread-load kernel run the same
    # arithmetic: identical per-trial outputs, for full tiles, a ragged last tile, and step counts that
    # are / are not a multiple of the staging depth
    from kf_bayesopt_b200 import _lib
    rng = np.random.default_rng(11)
    for B, T, dt in [(256, 201, 0.1), (130, 7, 0.5), (2, 2, 1.0), (1026, 2000, 0.1), (4096, 211, 0.5)]:
        rv = kfbo.ReadParaVehicle(nsimruns_=B, cpu_core_number_=1)
        x0n, w, v = _noise(rng, oracle, dt, T, B)
        out = {}
        with kfbo.CostEvaluator(rv, [(dt, T)], max_candidates=2) as ev:
            for name, variant in (("ldg", _lib.SUPPLIED_LDG), ("bulk", _lib.SUPPLIED_BULK)):
                ev.set_option(_lib.OPT_SUPPLIED_KERNEL, variant)
                out[name] = ev.eval_supplied(0, [0.8, 2.0], [0.15, 0.05], x0n, w, v)
        assert np.array_equal(out["ldg"][0], out["bulk"][0]), (B, T)
        assert np.array_equal(out["ldg"][1], out["bulk"][1]), (B, T)
        want = oracle.eval_supplied(dt, T, 0.8, 0.15, x0n, w, v)
        assert rel_err(out["bulk"][0][0], want) < RTOL, (B, T)


def test_single_trial_trace_matches_rollout_and_oracle_covariance(kfbo, oracle):
    # the Nsimruns == 1 record (trial.cpp:69-78): replaying one trajectory step by step gives the per-step
    # NEES/NIS whose sums / time-variances are that trial's rollout outputs, and +-2 sqrt(P_ii) bands that
    # follow the oracle's (data-independent) covariance recursion
    seed = 99
    rv = kfbo.ReadParaVehicle(nsimruns_=64, cpu_core_number_=1)
    times = [(0.1, 201), (0.5, 212)]
    with kfbo.CostEvaluator(rv, times, seed=seed) as ev:
        for k, (dt, T) in enumerate(times):
            for qe, re, stream, trial in [(1.0, 0.1, 0, 0), (0.4, 0.3, 11, 37)]:
                tr = ev.trace_trial(k, qe, re, stream, trial)
                assert tr.shape == (T, 9)
                np.testing.assert_allclose(tr[:, 0], np.arange(T) * dt, rtol=1e-15)
                pt = ev.eval_philox_trials(k, qe, re, stream, trial, 1)[:, 0]
                want = [tr[:, 7].sum(), tr[:, 8].sum(), tr[:, 7].var(ddof=1), tr[:, 8].var(ddof=1)]   # trial.cpp:62-63,84-96
                np.testing.assert_allclose(pt, want, rtol=1e-9)
                ric = oracle.riccati_trace(dt, qe, re, T)        # columns: S, K0, K1, P00, P01, P10, P11
                np.testing.assert_allclose(tr[:, 4], 2 * np.sqrt(ric[:, 3]), rtol=1e-9)
                np.testing.assert_allclose(tr[:, 6], 2 * np.sqrt(ric[:, 6]), rtol=1e-9)
                assert np.array_equal(tr[:, 3], -tr[:, 4]) and np.array_equal(tr[:, 5], -tr[:, 6])
                want_pt = oracle.eval_philox(seed, stream, dt, T, Q_TRUTH, R_TRUTH, qe, re, trial, 1)[:, 0]
                np.testing.assert_allclose(want, want_pt, rtol=1e-9)
                # a consistent filter keeps the error inside its 95 % band most of the time
                if (qe, re) == (1.0, 0.1):
                    assert np.mean(np.abs(tr[:, 1]) <= tr[:, 4]) > 0.85
