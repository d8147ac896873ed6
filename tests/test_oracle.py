"""CPU tests: the oracle (oracle/kfbo_oracle.c) against the committed golden vectors
(tests/golden/kf_golden.npz, made by the independent numpy/scipy restatement in make_golden.py),
published known answers of the RNGs it restates, and statistical anchors.

The reference ships no tests (CMakeLists.txt:221-232 commented out); these are the pins."""
import numpy as np
import pytest

from conftest import rel_err

TOL = 1e-12  # oracle vs numpy restatement: same arithmetic, different summation/BLAS rounding only


def test_discretisation_matches_scipy_expm(oracle, golden):
    for (q, dt), Fd, Qd in zip(golden["disc_in"], golden["disc_Fd"], golden["disc_Qd"]):
        F, Q = oracle.discretize(q, dt)
        assert np.array_equal(F, Fd), (q, dt)           # [[1,dt],[0,1]] exactly
        assert rel_err(Q, Qd) < 4 * np.finfo(float).eps, (q, dt)


def test_discretisation_closed_form(oracle):
    # continuous2discrete.hpp on the white-noise-acceleration model: Fd=[[1,dt],[0,1]], Qd=q[[dt^3/3,dt^2/2],[.,dt]]
    for q, dt in [(1.0, 0.1), (1.0, 0.5), (1.0, 1.0), (3.3, 0.25)]:
        F, Q = oracle.discretize(q, dt)
        assert np.array_equal(F, [[1, dt], [0, 1]])
        np.testing.assert_allclose(Q, q * np.array([[dt**3 / 3, dt**2 / 2], [dt**2 / 2, dt]]), rtol=1e-15)
    # SURVEY 8(c) discretisation KAT
    _, Q = oracle.discretize(1.0, 0.1)
    assert Q[0, 0] == pytest.approx(3.3333333333333343e-4, rel=1e-15)


def test_control_table_accumulates_dt(oracle, golden):
    for dt in (0.1, 0.5, 1.0):
        u = oracle.control_table(dt)
        assert u.shape == (2000,)
        assert np.array_equal(u, golden[f"u_{dt}"])
    # j += dt, NOT i*dt (system_models.hpp:54): differs from the closed form at the 1e-11 level
    u = oracle.control_table(0.1)
    closed = 2 * np.cos(0.75 * 0.1 * np.arange(2000))
    assert 0 < np.abs(u - closed).max() < 1e-10


def test_riccati_known_answers(oracle, golden):
    tr = oracle.riccati_trace(0.1, 1.0, 0.1, 201)
    assert rel_err(tr, golden["riccati_0.1"]) < TOL
    # SURVEY 8(c) Riccati KAT (q=1, r=0.1, dt=0.1, P0=I)
    assert tr[0, 0] == pytest.approx(2.0103333333333335, rel=1e-14)
    assert tr[0, 1] == pytest.approx(0.5025700547172939, rel=1e-14)
    assert tr[0, 2] == pytest.approx(0.05223014425468413, rel=1e-14)
    assert tr[1, 0] == pytest.approx(1.5242945752500967, rel=1e-14)
    assert tr[200, 0] == pytest.approx(1.2859356657862273, rel=1e-13)
    assert tr[200, 6] == pytest.approx(0.7473678281766549, rel=1e-13)
    tr = oracle.riccati_trace(1.0, 0.1, 0.01, 201)
    assert rel_err(tr, golden["riccati_1.0"]) < TOL


def test_supplied_noise_trials_match_numpy_restatement(oracle, golden):
    B = int(golden["meta"][1])
    for d, (dt, T) in enumerate(golden["sample_times"]):
        for c, (qe, re) in enumerate(golden["candidates"]):
            pt = oracle.eval_supplied(dt, int(T), qe, re, golden[f"x0noise_{d}"], golden[f"w_{d}"], golden[f"v_{d}"])
            assert pt.shape == (4, B)
            assert rel_err(pt, golden[f"per_trial_{d}_{c}"]) < TOL, (d, c)


def test_cost_assembly_matches_numpy_restatement(oracle, golden):
    B, cores = int(golden["meta"][1]), int(golden["meta"][2])
    for d, (dt, T) in enumerate(golden["sample_times"]):
        for c in range(len(golden["candidates"])):
            pt = golden[f"per_trial_{d}_{c}"]
            for i, choice in enumerate(["JNEES", "CNEES", "JNIS", "CNIS"]):
                out = oracle.cost_from_per_trial(pt, int(T), B, cores, cost_choice=choice)
                assert rel_err(out[:4], golden[f"avg_{d}_{c}"]) < TOL
                assert out[4] == pytest.approx(golden[f"J_{d}_{c}"][i], rel=1e-11, abs=1e-13)


def test_thread_chunking_integer_semantics(oracle):
    # trial.cpp:19,28: n = nsimruns/cores trials run, but k_average = (T*nsimruns)/cores (int division of the product)
    rng = np.random.default_rng(0)
    T, nsim, cores = 201, 23, 4
    n = nsim // cores
    pt = rng.uniform(1, 3, size=(4, n * cores))
    avg = oracle.thread_averages(pt, 0, T, nsim, cores)
    assert avg[0] == pytest.approx(pt[0, :n].sum() / float(T * nsim // cores), rel=1e-14)
    assert avg[2] == pytest.approx(pt[2, :n].sum() * cores / nsim, rel=1e-14)
    assert float(T * nsim // cores) != T * n  # the reference's averages are biased when nsim % cores != 0


def test_max_cost_takes_first_maximum(oracle):
    assert oracle.max_cost([0.3, 0.7]) == (0.7, 1)
    assert oracle.max_cost([0.7, 0.3]) == (0.7, 0)
    assert oracle.max_cost([0.5, 0.5]) == (0.5, 0)     # std::max_element: first on ties (trial_node.cpp:118-119)
    assert oracle.max_cost([0.1, 0.9, 0.9, 0.2]) == (0.9, 1)


def test_philox4x32_10_known_answers(oracle):
    # Random123 kat_vectors (philox4x32 10)
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, want in kat:
        assert tuple(int(x) for x in oracle.philox4x32_10(ctr, key)) == want


def test_mt19937_known_answer(oracle):
    # ISO C++ [rand.predef]: the 10000th output of a default-constructed mt19937 is 4123659995
    assert oracle.mt19937_nth(5489, 10000) == 4123659995


def test_normal_pair_mapping(oracle):
    # the spec of kfbo_math.cuh::normal_pair: radius from m = (w0 << 12 | w1 >> 20), u1 = 1 - (m + 1/2) 2^-44;
    # direction dir = w1 & 0xFFFFF, theta = 2 pi (dir + 1/2) / 2^20
    def want(w0, w1):
        m = (w0 << 12) | (w1 >> 20)
        u1 = 1.0 - (m + 0.5) * 2.0**-44
        r = np.sqrt(-2 * np.log(u1))
        th = 2 * np.pi * ((w1 & 0xFFFFF) + 0.5) / 2.0**20
        return r * np.cos(th), r * np.sin(th)
    # m = 0: the smallest radius, sqrt(-2 ln(1 - 2^-45)) ~ 2^-22; m max: u1 = 2^-45, the largest, 7.9 sigma
    a, b = oracle.normal_pair(0, 0)
    assert np.hypot(a, b) == pytest.approx(np.sqrt(2.0**-44), rel=1e-12)
    a, b = oracle.normal_pair(0xffffffff, 0xfff00000)
    assert np.hypot(a, b) == pytest.approx(np.sqrt(-2 * np.log(2.0**-45)), rel=1e-15) and 7.89 < np.hypot(a, b) < 7.9
    # the direction only depends on the low 20 bits of w1, the radius on the rest
    r0 = np.hypot(*oracle.normal_pair(0x12345678, 0xABC00000))
    for d in (0, 1, 0x3FFFF, 0x40000, 0x80000, 0xC0000, 0xFFFFF, 0x5A5A5):
        a, b = oracle.normal_pair(0x12345678, 0xABC00000 | d)
        wa, wb = want(0x12345678, 0xABC00000 | d)
        assert a == pytest.approx(wa, rel=1e-13, abs=1e-15) and b == pytest.approx(wb, rel=1e-13, abs=1e-15)
        assert np.hypot(a, b) == pytest.approx(r0, rel=1e-14)
    # every octant
    for o in range(8):
        w1 = 0x7FF00000 | (o << 17) | 0x0BCDE
        a, b = oracle.normal_pair(0x80000000, w1)
        wa, wb = want(0x80000000, w1)
        assert a == pytest.approx(wa, rel=1e-13, abs=1e-15) and b == pytest.approx(wb, rel=1e-13, abs=1e-15)
    # moments of the stream (a Philox call holds two pairs)
    n = 20000
    z = []
    for i in range(n):
        w = oracle.philox4x32_10((i, 7, 1, 0), (11, 22))
        z.append(oracle.normal_pair(w[0], w[1]))
        z.append(oracle.normal_pair(w[2], w[3]))
    z = np.array(z).ravel()
    assert abs(z.mean()) < 4 / np.sqrt(z.size)
    assert abs(z.var() - 1) < 4 * np.sqrt(2 / z.size)
    assert abs((z**4).mean() - 3) < 0.15
    assert abs(np.corrcoef(z[0::2], z[1::2])[0, 1]) < 4 / np.sqrt(z.size / 2)


def test_direction_grid_reproduces_the_trigonometric_moments():
    # why 2^20 equally spaced directions are enough: on a uniform grid of N angles every moment E[cos^p sin^q] with
    # p + q < N equals the continuous one exactly (aliasing starts at order N), so the joint moments of a Box-Muller pair
    # are those of two independent normals up to order 2^20 - 1.  Checked numerically on the grid of the spec.
    from math import gamma, pi
    N = 1 << 20
    th = 2 * np.pi * (np.arange(N) + 0.5) / N
    c, s = np.cos(th), np.sin(th)
    for p, q in [(2, 0), (0, 2), (1, 1), (4, 0), (2, 2), (6, 2), (8, 8), (3, 1), (20, 10), (40, 40)]:
        got = np.mean(c**p * s**q)
        exact = 0.0 if (p % 2 or q % 2) else gamma((p + 1) / 2) * gamma((q + 1) / 2) / (pi * gamma((p + q) / 2 + 1))
        assert got == pytest.approx(exact, rel=1e-11, abs=1e-15), (p, q)


def test_noise_transforms_reproduce_covariance(oracle):
    for q, dt in [(1.0, 0.1), (1.0, 1.0), (2.5, 0.5)]:
        _, Qd = oracle.discretize(q, dt)
        L, sr = oracle.noise_transform(Qd, 0.1 / dt)
        Lm = np.array([[L[0], 0], [L[1], L[2]]])
        np.testing.assert_allclose(Lm @ Lm.T, Qd, rtol=1e-14)
        assert sr == pytest.approx(np.sqrt(0.1 / dt))
        Tm = oracle.eig_transform2(Qd)      # eigenmvn.h:126-130
        np.testing.assert_allclose(Tm @ Tm.T, Qd, rtol=1e-13, atol=1e-18)


def test_philox_eval_is_shard_invariant(oracle):
    # counters use the global trial index: evaluating [0,24) at once == [0,10) and [10,24) separately
    args = (0x1234, 5, 0.1, 61, 1.0, 0.1, 0.7, 0.2)
    full = oracle.eval_philox(*args, 0, 24)
    a, b = oracle.eval_philox(*args, 0, 10), oracle.eval_philox(*args, 10, 14)
    assert np.array_equal(full, np.concatenate([a, b], axis=1))
    # different stream -> different draws
    other = oracle.eval_philox(0x1234, 6, 0.1, 61, 1.0, 0.1, 0.7, 0.2, 0, 24)
    assert not np.allclose(full, other)


def test_statistical_anchors_philox(oracle):
    # matched filter -> E[NEES]=2 (state dof), E[NIS]=1; q_est=0.1 -> NEES ~ 10.0 at dt=0.1
    B, T = 4000, 201
    pt = oracle.eval_philox(42, 0, 0.1, T, 1.0, 0.1, 1.0, 0.1, 0, B)
    nees = pt[0] / T
    se = nees.std(ddof=1) / np.sqrt(B)
    assert abs(nees.mean() - 1.9988) < 5 * se
    assert abs((pt[1] / T).mean() - 1.0) < 5 * (pt[1] / T).std(ddof=1) / np.sqrt(B)
    assert abs(pt[2].mean() - 3.857) < 0.15 and abs(pt[3].mean() - 2.001) < 0.08
    pt = oracle.eval_philox(42, 1, 0.1, T, 1.0, 0.1, 0.1, 0.1, 0, B)
    nees = pt[0] / T
    assert abs(nees.mean() - 10.035) < 5 * nees.std(ddof=1) / np.sqrt(B)


def test_statistical_anchors_threaded_mt(oracle):
    # the timed CPU baseline path (mt19937 + polar normals + eigen transform, threaded like trial_node.cpp:54-67)
    out = oracle.eval_mt(0.1, 201, 1.0, 0.1, 1.0, 0.1, 8000, 8, seed=3)
    assert abs(out[0] - 1.9988) < 0.03 and abs(out[1] - 1.0) < 0.01
    assert out[4] == pytest.approx(abs(np.log(out[0] / 2)), rel=1e-14)
    out = oracle.eval_mt(1.0, 201, 1.0, 0.1, 0.1, 0.1, 8000, 8, seed=4, cost_choice="JNIS")
    assert abs(out[0] - 9.776) < 0.15
    assert out[4] == pytest.approx(abs(np.log(out[1] / 1)), rel=1e-14)
    with pytest.raises(ValueError):
        oracle.eval_mt(0.1, 2001, 1.0, 0.1, 1.0, 0.1, 80, 8)   # T > 2000: control table limit


# ---------------------------------------------------------------------------------------------
# The parity envelope: how far outside the yaml bounds the 1e-9 bound can hold at all.
# ---------------------------------------------------------------------------------------------
def test_envelope_reference_formulas_lose_accuracy_as_r_est_goes_to_zero(oracle):
    import envelope_model as em
    rng = np.random.default_rng(20261021)
    B = 48
    worst = {}
    for name, dt, T, qt, rt, qe, re_, inside in em.CASES:
        x0n, w, v = em.noise(oracle, rng, dt, T, B, qt, rt)
        want = oracle.eval_supplied(dt, T, qe, re_, x0n, w, v)
        # the numpy restatement in float64 IS the oracle, bit for bit: the model of the two formulations is faithful
        assert np.array_equal(em.run(oracle, "reference", dt, T, qe, re_, x0n, w, v), want), name
        exact = em.run(oracle, "device", dt, T, qe, re_, x0n, w, v, em.LD)   # P = R K in long double: no cancellation, the most accurate value at hand
        dev64 = em.run(oracle, "device", dt, T, qe, re_, x0n, w, v)
        worst[name] = (em.rel(want, exact), em.rel(dev64, exact), em.rel(dev64, want))
        # the device's formulation (P = R K: no 1 - k0 cancellation) stays at rounding level everywhere ...
        assert worst[name][1] < 1e-10, (name, worst[name])
        # ... and agrees with the reference arithmetic to 1e-9 wherever the reference itself is that accurate
        if inside:
            assert worst[name][2] < 1e-9, (name, worst[name])
    # r_est -> 0: (I - K H) P- cancels (1 - k0 = R/S ~ 1e-9 of k0), the REFERENCE side drifts past 1e-9
    assert worst["r_est 1e-10"][0] > 1e-9 > worst["r_est 1e-10"][1]
    assert worst["r_est 1e-8"][0] > 100 * worst["r_est 1e-8"][1]


def test_philox_stream_ids_are_62_bits_wide(oracle):
    seed, T = 0xABCDEF, 9
    a = oracle.eval_philox(seed, 5, 0.1, T, 1.0, 0.1, 1.0, 0.1, 0, 8)
    b = oracle.eval_philox(seed, 5 + (1 << 32), 0.1, T, 1.0, 0.1, 1.0, 0.1, 0, 8)      # same low word, other stream
    c = oracle.eval_philox(seed, 5 + (1 << 61), 0.1, T, 1.0, 0.1, 1.0, 0.1, 0, 8)
    d = oracle.eval_philox(seed, 5 + (1 << 62), 0.1, T, 1.0, 0.1, 1.0, 0.1, 0, 8)      # bit 62 is outside the id
    assert not np.array_equal(a, b) and not np.array_equal(a, c) and not np.array_equal(b, c)
    assert np.array_equal(a, d)
