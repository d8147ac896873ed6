"""Test infrastructure: a numpy restatement of one supplied-noise trial batch in a chosen floating-point type, in the
two algebraic forms that exist in this repository, for the ENVELOPE tests (how far from the usual parameter range the
1e-9 parity bound holds, and which side is the inaccurate one where it does not):

  form "reference": the operation order of kalman_filter.cpp:18-38 / trial.cpp:53-60 as the oracle restates it --
                    P = (I - K H) P-, all four covariance entries, P^-1 by adjugate x 1/det.  In float64 this reproduces
                    oracle.eval_supplied BIT FOR BIT (checked in tests/test_oracle.py), so in long double it is "what the
                    reference's formulas give without rounding error".
  form "device":    what filter_step (kf_bayesopt_b200/csrc/kfbo_device.cuh) evaluates -- P = R K (no 1 - k0
                    cancellation), symmetric covariance, the division applied once to the quadratic form.

Returns per_trial[4][B] = row_nees, row_nis, var_nees, var_nis (trial.cpp:62-63,95-96)."""
import numpy as np

LD = np.longdouble


def run(oracle, form, dt, T, q_est, r_est, x0n, w, v, dtype=np.float64):
    B = x0n.shape[1]
    Fd, Qd = oracle.discretize(q_est, dt)
    f = dtype(Fd[0, 1])
    q00, q01, q10, q11 = [dtype(x) for x in Qd.ravel()]
    R = dtype(r_est) / dtype(dt)                       # system_models.hpp:143
    u = oracle.control_table(dt, 2000)
    g0, g1 = 0.5 * dt * dt, dt                         # system_models.hpp:46-47 (the products g*u are rounded doubles)
    x0, x1 = x0n[0].astype(dtype), x0n[1].astype(dtype)
    xh0, xh1 = np.zeros(B, dtype), np.zeros(B, dtype)
    p00, p01, p10, p11 = np.ones(B, dtype), np.zeros(B, dtype), np.zeros(B, dtype), np.ones(B, dtype)
    nees, nis = np.zeros((T, B), dtype), np.zeros((T, B), dtype)
    for s in range(T):
        gu0, gu1 = dtype(g0 * u[s]), dtype(g1 * u[s])
        x0, x1 = x0 + f * x1 + gu0 + w[s, 0].astype(dtype), x1 + gu1 + w[s, 1].astype(dtype)
        z = x0 + v[s].astype(dtype)
        xp0, xp1 = xh0 + f * xh1 + gu0, xh1 + gu1
        a00, a01, a10, a11 = p00 + f * p10, p01 + f * p11, p10, p11              # Fd P
        pp00, pp01, pp10, pp11 = a00 + a01 * f + q00, a01 + q01, a10 + a11 * f + q10, a11 + q11
        S = pp00 + R
        sinv = 1 / S
        k0, k1 = pp00 * sinv, pp10 * sinv
        nu = z - xp0
        xh0, xh1 = xp0 + k0 * nu, xp1 + k1 * nu
        if form == "reference":
            i00, i10 = 1 - k0, -k1
            p00, p01, p10, p11 = i00 * pp00, i00 * pp01, i10 * pp00 + pp10, i10 * pp01 + pp11
        else:
            p00, p01, p11 = R * k0, R * k1, pp11 - k1 * pp01
            p10 = p01
        e0, e1 = x0 - xh0, x1 - xh1
        idet = 1 / (p00 * p11 - p01 * p10)
        if form == "reference":
            i00, i01, i10, i11 = p11 * idet, -p01 * idet, -p10 * idet, p00 * idet
            nees[s] = (e0 * i00 + e1 * i10) * e0 + (e0 * i01 + e1 * i11) * e1
        else:
            nees[s] = (e0 * (p11 * e0 - p01 * e1) + e1 * (p00 * e1 - p01 * e0)) * idet
        nis[s] = nu * sinv * nu
    return np.stack([nees.sum(0), nis.sum(0), nees.var(0, ddof=1), nis.var(0, ddof=1)])


def rel(a, b):
    a, b = np.asarray(a, dtype=LD), np.asarray(b, dtype=LD)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), LD(1e-300))))


# (name, dt, T, q_truth, r_truth, q_est, r_est, inside): the corners the reference can reach beyond the yaml bounds
# [0.1, 5] x [0.01, 1] (config/bayesParams.yaml:18-19).  inside = the 1e-9 bound against the ORACLE is asserted.
CASES = [
    ("yaml centre", 0.1, 201, 1.0, 0.1, 1.0, 0.1, True),
    ("r_est 1e-4", 0.1, 201, 1.0, 0.1, 1.0, 1e-4, True),
    ("r_est 1e-6", 0.1, 201, 1.0, 0.1, 1.0, 1e-6, True),
    ("r_est 1e-8", 0.1, 201, 1.0, 0.1, 1.0, 1e-8, True),
    ("r_est 1e-10", 0.1, 201, 1.0, 0.1, 1.0, 1e-10, False),
    ("q_est 1e-6", 0.1, 201, 1.0, 0.1, 1e-6, 0.1, True),
    ("q_est 1e3", 0.1, 201, 1.0, 0.1, 1e3, 0.1, True),
    ("q_truth 1e-2, q_est 1e2", 0.1, 201, 1e-2, 0.1, 1e2, 0.1, True),
    ("q_truth 1e2, q_est 1e-2", 0.1, 201, 1e2, 0.1, 1e-2, 0.1, True),
    ("T 2000, dt 1.0", 1.0, 2000, 1.0, 0.1, 1.0, 0.1, True),
    ("T 2000, dt 1.0, r_est 1e-6", 1.0, 2000, 1.0, 0.1, 1.0, 1e-6, True),
]


def noise(oracle, rng, dt, T, B, q_truth, r_truth):
    _, Qd = oracle.discretize(q_truth, dt)
    L = np.linalg.cholesky(Qd)
    x0n = rng.standard_normal((2, B))
    w = np.ascontiguousarray(np.einsum("ij,tjb->tib", L, rng.standard_normal((T, 2, B))))
    v = np.sqrt(r_truth / dt) * rng.standard_normal((T, B))
    return x0n, w, v
