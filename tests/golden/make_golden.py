#!/usr/bin/env python
"""Generate tests/golden/kf_golden.npz -- the pins for the CPU oracle and the CUDA path.

The reference (arpg/kf_bayesopt) has no tests, fixtures or golden vectors and cannot be
compiled in this image (no Eigen3 / ROS), so these vectors come from an INDEPENDENT numpy /
scipy restatement of the reference arithmetic written directly from the reference sources
(it shares no code with oracle/kfbo_oracle.c):

  * discretisation: scipy.linalg.expm on the Van Loan matrix   (include/continuous2discrete.hpp:8-32,
                                                                 include/system_models.hpp:33-39,58)
  * control table u[i] = 2cos(0.75 j), j += dt                  (include/system_models.hpp:49-56)
  * simulator / Kalman filter / NEES / NIS / time-variance      (simulator.cpp:7-19, kalman_filter.cpp:12-38,
                                                                 trial.cpp:28-113)
  * thread averages, costs and the max over sample times        (trial.cpp:19,110-113, trial_node.cpp:70-119)

Run from the repo root:  python tests/golden/make_golden.py
Needs only numpy + scipy (no /root/reference access at run time).
"""
import os

import numpy as np
import scipy.linalg as sl

SEED = 20260101
Q_TRUTH, R_TRUTH = 1.0, 0.1   # config/vehicleParams.yaml:18-19
B, CORES = 16, 4
SAMPLE_TIMES = [(0.1, 201), (1.0, 201), (0.5, 211)]   # trial_node.cpp:105-114 (and the commented 0.5 variant)
CANDIDATES = [(1.0, 0.1), (0.1, 0.1), (5.0, 1.0), (2.5, 0.01)]


def discretize(q, dt):
    Fc = np.array([[0.0, 1.0], [0.0, 0.0]])
    Qc = np.array([[0.0, 0.0], [0.0, q]])
    A = np.zeros((4, 4))
    A[:2, :2] = -Fc * dt
    A[:2, 2:] = Qc * dt
    A[2:, 2:] = Fc.T * dt
    Bm = sl.expm(A)
    Fd = Bm[2:, 2:].T.copy()
    Qd = Fd @ Bm[:2, 2:]
    return Fd, Qd


def control_table(dt, n=2000):
    u, j = np.empty(n), 0.0
    for i in range(n):
        u[i] = 2 * np.cos(0.75 * j)
        j += dt
    return u


def run_trial(dt, T, q_est, r_est, x0n, w, v):
    """One trial with supplied post-transform noise; returns (row_nees, row_nis, var_nees, var_nis)."""
    Fd, _ = discretize(1.0, dt)
    _, Qe = discretize(q_est, dt)
    g = np.array([0.5 * dt * dt, dt])
    H = np.array([[1.0, 0.0]])
    Re = r_est / dt
    u = control_table(dt)
    x = np.zeros(2) + x0n
    xe, P = np.zeros(2), np.eye(2)
    nees, nis = np.empty(T), np.empty(T)
    for k in range(T):
        x = Fd @ x + g * u[k] + w[k]
        z = (H @ x)[0] + v[k]
        xp = Fd @ xe + g * u[k]
        Pp = Fd @ P @ Fd.T + Qe
        c = Pp @ H.T
        S = (H @ c)[0, 0] + Re
        K = c[:, 0] * (1.0 / S)
        nu = z - (H @ xp)[0]
        xe = xp + K * nu
        P = (np.eye(2) - np.outer(K, H[0])) @ Pp
        e = x - xe
        det = P[0, 0] * P[1, 1] - P[1, 0] * P[0, 1]
        Pinv = np.array([[P[1, 1], -P[0, 1]], [-P[1, 0], P[0, 0]]]) * (1.0 / det)
        nees[k] = e @ Pinv @ e
        nis[k] = nu * (1.0 / S) * nu
    row_nees, row_nis = 0.0, 0.0
    for k in range(T):
        row_nees += nees[k]
        row_nis += nis[k]
    an, ai = row_nees / T, row_nis / T
    vn = sum((nees[k] - an) ** 2 for k in range(T)) / (T - 1)
    vi = sum((nis[k] - ai) ** 2 for k in range(T)) / (T - 1)
    return row_nees, row_nis, vn, vi


def thread_cost(per_trial, T, nsimruns, cores, state_dof=2, obs_dof=1):
    """Thread averages (trial.cpp:19,110-113) -> mean over threads -> all four costs (trial_node.cpp:70-102)."""
    n = nsimruns // cores
    k_average = float(T * nsimruns // cores)
    a = np.zeros(4)
    for th in range(cores):
        cols = per_trial[:, th * n:(th + 1) * n]
        s = [0.0, 0.0, 0.0, 0.0]
        for i in range(n):
            for m in range(4):
                s[m] += cols[m, i]
        t_avg = [s[0] / k_average, s[1] / k_average, s[2] * cores / nsimruns, s[3] * cores / nsimruns]
        for m in range(4):
            a[m] += t_avg[m] / cores
    jn = abs(np.log(a[0] / state_dof))
    ji = abs(np.log(a[1] / obs_dof))
    cn = jn + abs(np.log(0.5 * a[2] / state_dof))
    ci = ji + abs(np.log(0.5 * a[3] / obs_dof))
    return a, np.array([jn, cn, ji, ci])   # order = cost_choice JNEES, CNEES, JNIS, CNIS


def riccati(dt, q, r, T):
    Fd, Qd = discretize(q, dt)
    H = np.array([[1.0, 0.0]])
    R = r / dt
    P = np.eye(2)
    out = np.empty((T, 7))
    for k in range(T):
        Pp = Fd @ P @ Fd.T + Qd
        c = Pp @ H.T
        S = (H @ c)[0, 0] + R
        K = c[:, 0] * (1.0 / S)
        P = (np.eye(2) - np.outer(K, H[0])) @ Pp
        out[k] = [S, K[0], K[1], P[0, 0], P[0, 1], P[1, 0], P[1, 1]]
    return out


def main():
    out = {}
    # --- discretisation known answers
    disc_in = [(1.0, 0.1), (1.0, 0.5), (1.0, 1.0), (0.1, 0.1), (5.0, 1.0), (2.5, 0.5), (0.37, 0.02)]
    out["disc_in"] = np.array(disc_in)
    out["disc_Fd"] = np.array([discretize(q, dt)[0] for q, dt in disc_in])
    out["disc_Qd"] = np.array([discretize(q, dt)[1] for q, dt in disc_in])
    # --- control table
    for dt in (0.1, 0.5, 1.0):
        out[f"u_{dt}"] = control_table(dt)
    # --- data-independent Riccati trace
    out["riccati_0.1"] = riccati(0.1, 1.0, 0.1, 201)
    out["riccati_1.0"] = riccati(1.0, 0.1, 0.01, 201)
    # --- supplied-noise trials
    rng = np.random.Generator(np.random.PCG64(SEED))
    for d, (dt, T) in enumerate(SAMPLE_TIMES):
        _, Qt = discretize(Q_TRUTH, dt)
        L = np.linalg.cholesky(Qt)
        sr = np.sqrt(R_TRUTH / dt)
        x0n = rng.standard_normal((2, B))
        w = np.einsum("ij,tjb->tib", L, rng.standard_normal((T, 2, B)))
        v = sr * rng.standard_normal((T, B))
        out[f"x0noise_{d}"], out[f"w_{d}"], out[f"v_{d}"] = x0n, w, v
        for c, (qe, re) in enumerate(CANDIDATES):
            pt = np.empty((4, B))
            for i in range(B):
                pt[:, i] = run_trial(dt, T, qe, re, x0n[:, i], w[:, :, i], v[:, i])
            out[f"per_trial_{d}_{c}"] = pt
            a, J = thread_cost(pt, T, B, CORES)
            out[f"avg_{d}_{c}"], out[f"J_{d}_{c}"] = a, J
    out["sample_times"] = np.array(SAMPLE_TIMES)
    out["candidates"] = np.array(CANDIDATES)
    out["meta"] = np.array([SEED, B, CORES, Q_TRUTH, R_TRUTH])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kf_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
