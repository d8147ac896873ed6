#!/usr/bin/env python
"""bench.py -- filter-steps/s and ms per J_NEES evaluation of the Monte Carlo cost evaluation.

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): robot1d J_NEES,
1,048,576 Monte Carlo trials x 1000 steps per sample time, two sample times (dt 0.1 / 0.5) per
evaluation as RunTrial::ProcessNoiseCallback always runs (trial_node.cpp:49-115), FP64, on-device
Philox noise.  A "step" of this benchmark is ONE full cost evaluation ("noise" message -> "cost").
At N > 1 the same 1M trials are sharded over the ranks (strong scaling) and the [C][D][4] cost sums
meet in one ncclAllReduce inside libkfbo.so.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line on rank 0.  --impl reference times the reference's threaded CPU path on the host cores:
the reference's own sources as compiled into oracle/_ref (cpu_baseline.kind "reference"), else the oracle's port.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TRIALS = 1 << 20
STEPS_PER_TRIAL = 1000
SAMPLE_TIMES = [(0.1, STEPS_PER_TRIAL), (0.5, STEPS_PER_TRIAL)]
# --workload cfg4 (BASELINE.json configs[3]): 64 x 64 candidates over the yaml bounds x 1024 trials x the README
# sample-time pair, candidates sharded over the ranks, one ncclAllReduce of the [4096][2][4] sums
CFG4_GRID, CFG4_TRIALS, CFG4_SAMPLE_TIMES = 64, 1024, [(0.1, 201), (0.5, 211)]
CFG4_BOUNDS = ((0.1, 5.0), (0.01, 1.0))   # config/bayesParams.yaml:18-19
Q_TRUTH, R_TRUTH = 1.0, 0.1          # config/vehicleParams.yaml:18-19
CANDIDATE = ([1.0], [0.1])           # launch/robot1d_test_trail.launch:4-5 (estimator == truth)
FLOPS_PER_STEP = 137.0               # reference-faithful dense 2x2 count, RNG excluded
BYTES_PER_STEP_SUPPLIED = 24.0       # 3 doubles of supplied noise per filter-step
BYTES_PER_TRIAL_SUPPLIED = 48.0      # 2 doubles x0 noise in + 4 doubles out
METRIC = "filter-steps/sec (trials x timesteps x sample times); ms_per_step = ms per J_NEES eval"
UNIT = "filter-steps/s"
WORKLOAD = ("cfg3: robot1d J_NEES, %d trials x %d steps x %d sample times (dt 0.1/0.5) per evaluation, "
            "FP64, on-device Philox4x32-10 noise" % (TRIALS, STEPS_PER_TRIAL, len(SAMPLE_TIMES)))


def profiled_traffic(kernel: str, workload: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` at this bench's size, from the committed
    `ncu --set full` capture (profiles/traffic.json, written by tools/ncu_traffic.py); None if there is no capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(f"{kernel}:{workload}", {}).get("dram_bytes_per_launch")
    except Exception:
        return None


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock / throttle reasons / power DURING the timed region (B200_PROFILING.md recipe).  Sampled through NVML
    (the library nvidia-smi reads) every 10 ms from a thread of this process, so that a timed region of a few tens of
    milliseconds (8 GPUs) still holds samples; the nvidia-smi loop of the recipe (100 ms) is the fallback."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.rows, self.proc, self.index, self.source, self._stop = [], None, index, None, False

    def _nvml_loop(self, nv, dev):
        bits = [nv.nvmlClocksThrottleReasonHwSlowdown, nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                nv.nvmlClocksThrottleReasonSwThermalSlowdown, nv.nvmlClocksThrottleReasonSwPowerCap]
        mx = str(nv.nvmlDeviceGetMaxClockInfo(dev, nv.NVML_CLOCK_SM))
        while not self._stop:
            try:
                why = nv.nvmlDeviceGetCurrentClocksThrottleReasons(dev)
                row = [str(nv.nvmlDeviceGetClockInfo(dev, nv.NVML_CLOCK_SM)), mx, str(nv.nvmlDeviceGetPowerUsage(dev) / 1e3)]
                row += ["Active" if why & b else "Not Active" for b in bits]
                self.rows.append((time.perf_counter(), row))
            except Exception:
                pass
            time.sleep(0.01)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            uuid = str(self.index)
            if uuid.startswith("GPU-"):
                try:
                    dev = nv.nvmlDeviceGetHandleByUUID(uuid)
                except TypeError:
                    dev = nv.nvmlDeviceGetHandleByUUID(uuid.encode())
            else:
                dev = nv.nvmlDeviceGetHandleByIndex(int(uuid))
            nv.nvmlDeviceGetClockInfo(dev, nv.NVML_CLOCK_SM)
            self.source = "nvml, 10 ms"
            self.t = threading.Thread(target=self._nvml_loop, args=(nv, dev), daemon=True)
            self.t.start()
            return
        except Exception:
            pass
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi -lms 100"
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self, t0: float, t1: float):
        if self.source is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi / NVML unavailable"]}
        time.sleep(0.15 if self.proc is not None else 0.02)
        self._stop = True
        if self.proc is not None:
            self.proc.terminate()
        window = "timed region"
        rows = [r for t, r in self.rows if t0 <= t <= t1]
        if not rows:       # a region shorter than the sampling period: the samples right around it
            rows, window = [r for t, r in self.rows if t0 - 0.15 <= t <= t1 + 0.15] or [r for _, r in self.rows], "around the timed region"
        sm = sorted(float(r[0]) for r in rows if r[0].replace(".", "").isdigit())
        reasons = sorted({n for r in rows for n, v in zip(self.NAMES, r[3:7]) if v.lower().startswith("active")})
        power = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": float(rows[0][1]) if rows and rows[0][1].replace(".", "").isdigit() else None,
                "power_w_max": max(power) if power else None, "samples": len(rows), "reasons": reasons,
                "source": self.source, "window": window}


def cpu_reference_run(steps: int, warmup: int, budget_s: float, kind: str = "auto"):
    """The reference's threaded CPU trial path (trial_node.cpp:54-75: CPUcoreNumber Trial objects, one std::thread
    per Trial::Run, getters averaged) on all host cores; each step a bounded sample of the cfg3 workload (same T, D,
    candidate).  kind "reference": the reference's OWN trial.cpp / simulator.cpp / kalman_filter.cpp / headers as
    compiled into oracle/_ref/libkfbo_ref.so (oracle/Makefile; Eigen3 is not installed, oracle/ref_shim stands in for
    it).  kind "port": the oracle's restatement (heap-free, one mt19937 per thread).  "auto": reference if built."""
    from oracle import oracle as o
    from oracle import ref
    cores = os.cpu_count() or 1
    use_ref = kind == "reference" or (kind == "auto" and ref.available())

    def one_eval(nsim):
        for dt, T in SAMPLE_TIMES:
            if use_ref:
                _, T_ref = ref.eval_threads(dt, dt * T, nsim, cores, Q_TRUTH, R_TRUTH, CANDIDATE[0][0], CANDIDATE[1][0], clock0=12345)
                assert T_ref == T, (T_ref, T)
            else:
                o.eval_mt(dt, T, Q_TRUTH, R_TRUTH, CANDIDATE[0][0], CANDIDATE[1][0], nsim, cores, seed=12345)

    cal = (2 if use_ref else 16) * cores
    t = time.perf_counter()
    one_eval(cal)
    rate = cal / (time.perf_counter() - t)                      # trials/s for a D=2 evaluation
    per_step = min(10.0, max(0.2, budget_s / max(1, steps + warmup)))
    nsim = max(cores, int(rate * per_step) // cores * cores)
    for _ in range(warmup):
        one_eval(nsim)
    t = time.perf_counter()
    for _ in range(steps):
        one_eval(nsim)
    el = time.perf_counter() - t
    work = float(nsim) * STEPS_PER_TRIAL * len(SAMPLE_TIMES) * steps
    model = ""
    try:
        model = next(l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name"))
    except Exception:
        pass
    how = ("the reference's own trial.cpp / simulator.cpp / kalman_filter.cpp / eigenmvn.h: Trial::Run on one std::thread per "
           "core sharing the reference's static mt19937 (trial_node.cpp:54-75), compiled by oracle/Makefile with a stand-in "
           "for Eigen3, which is not installed" if use_ref else
           "mt19937 + polar normals per thread; heap-free restatement (oracle/kfbo_oracle.c), faster than the Eigen-dynamic original")
    return {"value": work / el, "unit": UNIT, "cores": cores, "kind": "reference" if use_ref else "port",
            "sample": f"{steps} evaluations of {nsim} trials x {STEPS_PER_TRIAL} steps x {len(SAMPLE_TIMES)} sample times "
                      f"({el:.1f} s; {how})", "cpu_model": model}, el / steps * 1e3, nsim


def run_reference(args, rank: int):
    if rank != 0:
        return
    cb, ms, nsim = cpu_reference_run(args.steps, args.warmup, budget_s=120.0)
    if cb["kind"] == "reference":      # the oracle's restatement beside it, for scale (a few seconds)
        port, _, _ = cpu_reference_run(2, 1, budget_s=6.0, kind="port")
        cb["port"] = {k: port[k] for k in ("value", "unit", "cores", "kind", "sample")}
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample_trials_per_step": nsim, "note": "CPU path, bounded sample"},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def run_bo_loop(args):
    """BASELINE.json configs[4]: the BO loop (20 initial + 180 iterations = 200 cost evaluations of the default
    vehicleParams.yaml workload: 200 trials x 201 steps x 2 sample times) driving the GPU evaluator over the
    noise/cost topic contract -- the C++ program kf_bayesopt_b200/bin/kfbo_bo_loop (in-process topic bus, GP/EI
    stand-in for the absent bayesopt library).  CPU arm: the same 200 evaluations on the oracle's threaded path with
    the yaml's CPUcoreNumber = 10 threads, thread pool created per evaluation like trial_node.cpp:54-67."""
    exe = os.path.join(ROOT, "kf_bayesopt_b200", "bin", "kfbo_bo_loop")
    if not os.path.exists(exe):
        raise SystemExit("bench.py: build the host programs first (make -C kf_bayesopt_b200/host)")
    iters, init, B, T, D = 180, 20, 200, 201, 2
    out = subprocess.run([exe, "--iterations", str(iters), "--init", str(init), "--quiet"], capture_output=True, text=True, check=True)
    loop = json.loads(next(l for l in out.stdout.splitlines() if l.startswith("{")))
    # the same loop WITH the reference's fixed usleep(300 ms) + usleep(200 ms) per iteration (robot1d_bayesopt.cpp:91,96),
    # measured on a short run (20 evaluations = 10 s of sleep) rather than extrapolated
    outs = subprocess.run([exe, "--iterations", "10", "--init", "10", "--quiet", "--reference-sleeps"], capture_output=True, text=True, check=True)
    slept = json.loads(next(l for l in outs.stdout.splitlines() if l.startswith("{")))
    n = loop["iterations"]
    steps_total = float(n) * B * T * D
    line = {"metric": METRIC, "value": steps_total / loop["wall_s"], "unit": UNIT, "n_gpus": 1, "steps": n, "warmup": 0,
            "ms_per_step": loop["wall_s"] / n * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "cfg5: full Bayesian-optimisation loop, %d cost evaluations (20 initial + %d iterations) of "
                                   "%d trials x %d steps x %d sample times over the noise/cost topic contract" % (n, iters, B, T, D),
                       "optimiser": "GP (Matern-3/2 ARD) + EI stand-in for the absent bayesopt library; settings of bayesParams.yaml"},
            "bo_loop": loop,
            "bo_loop_with_reference_sleeps": {k: slept[k] for k in ("iterations", "wall_s", "evaluator_s", "evaluator_ms_per_iteration", "optimiser_s")},
            "e2e": {"value": steps_total / loop["wall_s"], "unit": UNIT, "h2d_bytes_per_step": 96 * D, "d2h_bytes_per_step": 32 * D,
                    "wall_s": loop["wall_s"], "evaluator_s": loop["evaluator_s"],
                    "wall_s_with_reference_sleeps": loop["wall_s"] + 0.5 * n,
                    "api": "kfbo_bo_loop: Robot1D::evaluateSample contract -> topic bus -> kfbo::RunTrial::ProcessNoiseCallback -> C ABI"},
            "gpu_launches": n}
    line["both_reference_nodes_unmodified"] = reference_nodes_loop()
    if not args.no_cpu_baseline:
        from oracle import oracle as o
        from oracle import ref
        cores = 10                                          # config/vehicleParams.yaml:1
        cands = [(1.0 + 0.01 * i, 0.1) for i in range(n)]
        if ref.available():
            # the reference's OWN node (trial_node.cpp compiled into oracle/_ref): the same number of "noise" messages
            # through RunTrial::ProcessNoiseCallback -- 10 Trial objects and std::threads per pass, two passes per message
            ref.set_params({"CPUcoreNumber": cores, "stateDOF": 2, "observationDOF": 1, "t_start": 0.0, "t_end": 20.1, "dt": 0.1,
                            "xi0": 0.0, "xidot0": 0.0, "Nsimruns": B, "pnoise": [Q_TRUTH], "onoise": [R_TRUTH],
                            "optimizationChoice": "all", "costChoice": "JNEES", "n_init_samples": init})   # config/*.yaml
            ref.node_run([([q], [r]) for q, r in cands[:8]], clock0=1)
            t = time.perf_counter()
            costs, _ = ref.node_run([([q], [r]) for q, r in cands], clock0=1)
            cpu_s = time.perf_counter() - t
            assert len(costs) == n
            kind, how = "reference", ("messages through the reference's own RunTrial::ProcessNoiseCallback (trial_node.cpp compiled into "
                                      "oracle/_ref with stand-ins for Eigen3 / ROS)")
        else:
            t = time.perf_counter()
            for q, r in cands:
                for dt in (0.1, 1.0):                       # trial_node.cpp:105-114
                    o.eval_mt(dt, T, Q_TRUTH, R_TRUTH, q, r, B, cores, seed=1)
            cpu_s = time.perf_counter() - t
            kind, how = "port", "evaluations by the oracle's threaded restatement"
        line["cpu_baseline"] = {"value": steps_total / cpu_s, "unit": UNIT, "cores": min(cores, os.cpu_count() or 1), "kind": kind,
                                "sample": f"{n} {how}: {B} trials x {T} steps x {D} sample times on {cores} threads "
                                          f"(yaml CPUcoreNumber), {cpu_s:.2f} s",
                                "evaluator_s": cpu_s,
                                "loop_wall_s_estimate": cpu_s + loop["optimiser_s"],
                                "loop_wall_s_estimate_with_reference_sleeps": cpu_s + loop["optimiser_s"] + 0.5 * n}
        line["per_iteration"] = {
            "evaluator_ms_gpu": loop["evaluator_ms_per_iteration"], "evaluator_ms_cpu": cpu_s / n * 1e3,
            "evaluator_speedup": (cpu_s / n * 1e3) / loop["evaluator_ms_per_iteration"],
            "wall_ms_gpu_no_sleeps": loop["wall_s"] / n * 1e3, "wall_ms_cpu_no_sleeps": (cpu_s + loop["optimiser_s"]) / n * 1e3,
            "wall_ms_gpu_with_sleeps_measured": slept["wall_s"] / slept["iterations"] * 1e3,
            "wall_ms_cpu_with_sleeps": (cpu_s + loop["optimiser_s"]) / n * 1e3 + 500.0,
            "note": "the loop is bound by the optimiser stand-in and, in the reference, by its 0.5 s of sleeps per iteration; the evaluator is what this repository replaces"}
    emit(line)


def reference_nodes_loop(n_init=20, n_iterations=180):
    """The same loop with BOTH of the reference's nodes unmodified in one process (tests/native/boloop: robot1d_bayesopt.cpp and
    trial_node.cpp built where they lie, one thread each, over the roscpp stand-in; trial_node.cpp on include/compat, so its
    Trial::Run calls -- CPUcoreNumber = 10 of them per sample time, on std::threads -- are launches on the GPU; the bayesopt LIBRARY
    is the GP / EI stand-in behind its interface).  The reference's sleeps are scaled to zero."""
    import ctypes as C
    import numpy as np
    lib = os.path.join(ROOT, "tests", "native", "boloop", "_boloop", "libkfbo_boloop.so")
    if not os.path.exists(lib):
        return {"unavailable": "tests/native/boloop is not built (needs the reference tree at build time)"}
    try:
        l = C.CDLL(lib)
        dp = C.POINTER(C.c_double)
        l.kfbo_boloop_param_clear.restype = None
        l.kfbo_boloop_param_set_int.argtypes = [C.c_char_p, C.c_longlong]
        l.kfbo_boloop_param_set_double.argtypes = [C.c_char_p, C.c_double]
        l.kfbo_boloop_param_set_string.argtypes = [C.c_char_p, C.c_char_p]
        l.kfbo_boloop_param_set_vector.argtypes = [C.c_char_p, dp, C.c_int]
        l.kfbo_boloop_run.restype = C.c_int
        l.kfbo_boloop_run.argtypes = [C.c_int, C.c_char_p, C.c_double, dp, dp, C.c_int, C.POINTER(C.c_int), dp, dp, C.c_char_p, C.c_int]
        # config/vehicleParams.yaml and config/bayesParams.yaml as `rosparam load` leaves them on the parameter server
        params = {"CPUcoreNumber": 10, "stateDOF": 2, "observationDOF": 1, "t_start": 0.0, "t_end": 20.1, "dt": 0.1, "xi0": 0.0, "xidot0": 0.0,
                  "Nsimruns": 200, "pnoise": [Q_TRUTH], "onoise": [R_TRUTH], "optimizationChoice": "all", "costChoice": "JNEES",
                  "surr_name": "sGaussianProcess", "kernel_name": "kMaternARD3", "crit_name": "cEI", "n_init_samples": n_init,
                  "n_iterations": n_iterations, "random_seed": 7, "verbose_level": 0, "noise": 1e-8, "force_jump": 10, "epsilon": 0.2,
                  "lower_bound": [0.1, 0.01], "upper_bound": [5.0, 1.0], "n_iteration_relearn": 10, "n_inner_evaluation": 500}
        l.kfbo_boloop_param_clear()
        for k, v in params.items():
            if isinstance(v, int):
                l.kfbo_boloop_param_set_int(k.encode(), v)
            elif isinstance(v, float):
                l.kfbo_boloop_param_set_double(k.encode(), v)
            elif isinstance(v, str):
                l.kfbo_boloop_param_set_string(k.encode(), v.encode())
            else:
                a = np.ascontiguousarray(v, dtype=np.float64)
                l.kfbo_boloop_param_set_vector(k.encode(), a.ctypes.data_as(dp), a.size)
        cap = n_init + n_iterations + 8
        noise, costs = np.zeros((cap, 4)), np.full(cap, np.nan)
        nc, wall, slept, log = C.c_int(0), C.c_double(0), C.c_double(0), C.create_string_buffer(1 << 12)
        n = l.kfbo_boloop_run(0, b"", 0.0, noise.ctypes.data_as(dp), costs.ctypes.data_as(dp), cap, C.byref(nc), C.byref(wall), C.byref(slept), log, len(log))
        costs = costs[:nc.value]
        return {"iterations": int(n), "costs_received": int(nc.value), "wall_s": wall.value, "ms_per_iteration": wall.value / max(n, 1) * 1e3,
                "best_cost": float(np.nanmin(costs)) if costs.size else None,
                "what": "robot1d_bayesopt.cpp (optimiser node) and trial_node.cpp (evaluator node, on include/compat: 2 x 10 Trial::Run "
                        "launches per message) compiled unmodified, one thread each, roscpp stand-in; the reference's 0.5 s of sleeps per "
                        "iteration scaled to 0; bayesopt library = the GP / EI stand-in"}
    except Exception as e:
        return {"unavailable": f"{type(e).__name__}: {e}"}


def leg_cfg5(ref_node_ms=None):
    """BASELINE.json configs[4] as a leg of the default line (N = 1): the 200-evaluation BO loop of kfbo_bo_loop over the
    noise/cost topic contract, without the reference's sleeps (the measured with-sleeps run and the reference node's own
    200 callbacks are `--workload cfg5`, profiles/bench_cfg5_r2*.json).  ref_node_ms: the reference node's callback time per
    message measured by the cfg1 leg of this same run."""
    exe = os.path.join(ROOT, "kf_bayesopt_b200", "bin", "kfbo_bo_loop")
    if not os.path.exists(exe):
        return {"unavailable": "kf_bayesopt_b200/bin/kfbo_bo_loop is not built (make -C kf_bayesopt_b200/host)"}
    try:
        out = subprocess.run([exe, "--iterations", "180", "--init", "20", "--quiet"], capture_output=True, text=True, check=True, timeout=120)
        loop = json.loads(next(l for l in out.stdout.splitlines() if l.startswith("{")))
    except Exception as e:                                   # a leg, not the headline: report, do not fail the line
        return {"unavailable": f"kfbo_bo_loop failed: {e}"}
    n = loop["iterations"]
    leg = {"workload": "cfg5: Bayesian-optimisation loop, %d cost evaluations (20 initial + 180 iterations) of 200 trials x 201 steps x 2 "
                       "sample times over the noise/cost topic contract; optimiser = GP/EI stand-in for the absent bayesopt library "
                       "(settings of bayesParams.yaml)" % n,
           "iterations": n, "wall_s": loop["wall_s"], "evaluator_s": loop["evaluator_s"],
           "evaluator_ms_per_iteration": loop["evaluator_ms_per_iteration"], "optimiser_s": loop["optimiser_s"],
           "wall_s_with_the_reference_sleeps": loop["wall_s"] + 0.5 * n, "best_cost": loop["best_cost"], "best_x": loop["best_x"],
           "note": "robot1d_bayesopt.cpp:91,96 sleeps 0.5 s per iteration: with them the loop is sleep-bound on either evaluator"}
    leg["both_reference_nodes_unmodified"] = reference_nodes_loop()
    if ref_node_ms:
        leg["reference_node_ms_per_iteration"] = ref_node_ms
        leg["evaluator_speedup"] = ref_node_ms / loop["evaluator_ms_per_iteration"]
        leg["wall_s_with_the_reference_node_estimate"] = loop["optimiser_s"] + ref_node_ms * 1e-3 * n
    return leg


def emit(line: dict) -> None:
    """The ONE JSON line, on the real stdout."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


# Everything else that writes to fd 1 (NCCL prints its version banner there, libraries may print warnings) goes
# to stderr, so stdout carries exactly one JSON line.
sys.stdout.flush()
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


# --------------------------------------------------------------------------------------------------------------------
# Legs of the default line besides the headline workload (every BASELINE.json config in a driver-run record)
# --------------------------------------------------------------------------------------------------------------------
DEFAULT_YAML = {"CPUcoreNumber": 10, "stateDOF": 2, "observationDOF": 1, "t_start": 0.0, "t_end": 20.1, "dt": 0.1, "xi0": 0.0,
                "xidot0": 0.0, "Nsimruns": 200, "pnoise": [Q_TRUTH], "onoise": [R_TRUTH], "optimizationChoice": "processNoise",
                "costChoice": "JNEES", "n_init_samples": 20}     # config/vehicleParams.yaml, config/bayesParams.yaml:10
CODE_TIMES = [(0.1, 201), (1.0, 201)]        # trial_node.cpp:105-114 (what the code runs)
README_TIMES = [(0.1, 201), (0.5, 211)]      # README.md:96 / BASELINE.json configs[1] ("0.1/0.5")


def timed_evals(torch, stream, ev, call, steps, warmup):
    """K evaluations timed with CUDA events on the launching stream -> (ms per evaluation, mean kernel ms, launches, last result)."""
    for _ in range(warmup):
        res = call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kms, launches = [], 0
    e0.record(stream)
    for _ in range(steps):
        res = call()
        ms, n = ev.last_kernel_ms()
        kms.append(ms)
        launches += n
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, float(sum(kms) / len(kms)), launches, res


def wall_evals(call, steps, warmup):
    for _ in range(warmup):
        call()
    t = time.perf_counter()
    for _ in range(steps):
        call()
    return (time.perf_counter() - t) / steps * 1e3


def oracle_slice_parity(kf, ev_factory, sample_times, q, r, seed, ntrials=300):
    """A <= 300-trial evaluation on the CUDA path against the CPU oracle on the SAME Philox streams (checker only):
    max relative error of the [D][4] cost sums."""
    import numpy as np
    from oracle import oracle as o
    with ev_factory(ntrials) as ev:
        ev.eval([q], [r])
        got = ev.last_sums(1)[0]
    want = np.stack([o.eval_philox(seed, k, dt, T, Q_TRUTH, R_TRUTH, q, r, 0, ntrials).sum(axis=1) for k, (dt, T) in enumerate(sample_times)])
    return float(np.max(np.abs(got - want) / np.abs(want)))


def leg_default_workload(kf, torch, stream, device, with_cpu):
    """BASELINE.json configs[0] and configs[1]: ONE "noise" -> "cost" evaluation of the default vehicleParams.yaml workload
    (200 trials x 201 steps x 2 sample times), as latency.  cfg1: optimizationChoice=processNoise (1-D candidate, onoise
    fixed at the yaml value), J_NEES, the code's sample-time pair, beside the REFERENCE NODE's own callback on the host
    cores (trial_node.cpp compiled into oracle/_ref, CPUcoreNumber = 10 threads).  cfg2: optimizationChoice=all (2-D
    candidate), the README pair 0.1/0.5, J_NEES and J_NIS out of the same pass."""
    import numpy as np
    out = {}
    seed = 0x4B46424F00002026
    for name, times, cand, choice in (("cfg1", CODE_TIMES, (0.7, R_TRUTH), "JNEES"), ("cfg2", README_TIMES, (0.7, 0.2), "JNEES")):
        rv = kf.ReadParaVehicle(nsimruns_=200, cpu_core_number_=10, pnoise_=[Q_TRUTH], onoise_=[R_TRUTH], cost_choice_=choice)
        with kf.CostEvaluator(rv, times, seed=seed, device=device) as ev:
            ev.set_cuda_stream(stream.cuda_stream)
            rt = kf.RunTrial(rv, evaluator=ev)
            msg = kf.Float64MultiArray.noise([cand[0]], [cand[1]])
            ms, kms, _, _ = timed_evals(torch, stream, ev, lambda: ev.eval([cand[0]], [cand[1]]), 200, 20)
            wall = wall_evals(lambda: rt.ProcessNoiseCallback(msg), 200, 20)
            res, phases = rt.last_result, ev.last_host_us()
            sums = ev.last_sums(1)[0]
            leg = {"workload": f"200 trials x 201 steps x 2 sample times {times}, candidate pnoise {cand[0]} onoise {cand[1]}",
                   "latency_ms": wall, "device_ms": ms, "kernel_ms": kms, "filter_steps_per_s": 200.0 * sum(T for _, T in times) / (wall * 1e-3),
                   "api": "RunTrial.ProcessNoiseCallback ('noise' -> 'cost') over the C ABI kfbo_eval, host buffers",
                   "host_phases_us": {k: round(v, 1) for k, v in phases.items()},
                   "J_NEES": [float(abs(np.log(a / 2.0))) for a in res.average_nees],
                   "J_NIS": [float(abs(np.log(a / 1.0))) for a in res.average_nis], "cost": float(res.max_cost), "index_max": int(res.index_max)}
        # the same workload's sums against the oracle on the same Philox streams (B = 200: the oracle needs milliseconds)
        leg["parity_vs_oracle_max_rel"] = oracle_slice_parity(
            kf, lambda n: kf.CostEvaluator(kf.ReadParaVehicle(nsimruns_=n, cpu_core_number_=10, pnoise_=[Q_TRUTH], onoise_=[R_TRUTH]), times,
                                           seed=seed, device=device), times, cand[0], cand[1], seed, ntrials=200)
        del sums
        out[name] = leg
    if with_cpu:
        from oracle import oracle as o
        from oracle import ref
        cand, n = (0.7, R_TRUTH), 20
        if ref.available():
            ref.set_params(dict(DEFAULT_YAML))
            msgs = [([cand[0]], [cand[1]])] * n
            ref.node_run(msgs[:4], clock0=1)
            t = time.perf_counter()
            costs, _ = ref.node_run(msgs, clock0=1)
            cpu_ms = (time.perf_counter() - t) / n * 1e3
            kind = "reference"
            how = (f"{n} 'noise' messages through the reference's own RunTrial::ProcessNoiseCallback (trial_node.cpp compiled into "
                   f"oracle/_ref with stand-ins for Eigen3 / ROS): 10 Trial objects + std::threads per pass, two passes")
        else:
            t = time.perf_counter()
            for _ in range(n):
                for dt, T in CODE_TIMES:
                    o.eval_mt(dt, T, Q_TRUTH, R_TRUTH, cand[0], cand[1], 200, 10, seed=1)
            cpu_ms = (time.perf_counter() - t) / n * 1e3
            kind, how = "port", f"{n} evaluations by the oracle's threaded restatement, 10 threads"
        out["cfg1"]["cpu_baseline"] = {"value": 200.0 * 402 / (cpu_ms * 1e-3), "unit": UNIT, "latency_ms": cpu_ms, "cores": min(10, os.cpu_count() or 1),
                                       "kind": kind, "sample": how}
    return out


def leg_cfg4(kf, torch, dist, stream, device, world, rank, steps=10, warmup=3):
    """BASELINE.json configs[3]: 4096 candidates x 1024 trials x 2 sample times, candidates sharded over the ranks, the
    [4096][2][4] sums gathered in one exchange.  Device-timed and end to end (host candidate arrays -> host costs)."""
    import numpy as np
    from kf_bayesopt_b200.dist import init_comm
    qq = np.repeat(np.geomspace(*CFG4_BOUNDS[0], CFG4_GRID), CFG4_GRID)
    rr = np.tile(np.geomspace(*CFG4_BOUNDS[1], CFG4_GRID), CFG4_GRID)
    rv = kf.ReadParaVehicle(nsimruns_=CFG4_TRIALS, cpu_core_number_=1, pnoise_=[Q_TRUTH], onoise_=[R_TRUTH])
    steps_per_eval = float(CFG4_TRIALS) * qq.size * sum(T for _, T in CFG4_SAMPLE_TIMES)
    with kf.CostEvaluator(rv, CFG4_SAMPLE_TIMES, seed=0x4B46424F00002026, device=device, max_candidates=qq.size) as ev:
        ev.set_cuda_stream(stream.cuda_stream)
        if world > 1:
            init_comm(ev, kf.SHARD_CANDIDATES)
            dist.barrier()
        ms, kms, launches, _ = timed_evals(torch, stream, ev, lambda: ev.eval_batch_costs(qq, rr), steps, warmup)
        if world > 1:
            dist.barrier()
        wall = wall_evals(lambda: ev.eval_batch_costs(qq, rr), steps, 1)
        t = torch.tensor([ms, wall, kms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, wall, kms = (float(x) for x in t.tolist())
        phases, path = ev.last_host_us(), ev.last_reduce_path()
    return {"workload": "cfg4: %d candidates (64 x 64 log grid over the yaml bounds) x %d trials x sample times %s" % (qq.size, CFG4_TRIALS, CFG4_SAMPLE_TIMES),
            "value": steps_per_eval / (ms * 1e-3), "unit": UNIT, "ms_per_sweep": ms, "kernel_ms": kms, "e2e_ms_per_sweep": wall,
            "e2e_value": steps_per_eval / (wall * 1e-3), "sharding": f"candidates over {world} rank(s); sums meet by: {path}",
            "api": "CostEvaluator.eval_batch_costs (host candidate arrays -> host cost arrays) over the C ABI",
            "host_phases_us": {k: round(v, 1) for k, v in phases.items()}, "gpu_launches_per_sweep": launches / steps}


def leg_one_process(kf, world, steps=20, warmup=3):
    """N > 1, rank 0 only, while the other ranks sit idle on the host: the SAME cfg3 evaluation and cfg4 sweep with ONE process
    driving all N GPUs (kfbo_group_*: one handle per device, one host thread per device, peer-memory exchange) -- the mode of
    the reference's node, which is one process.  Wall time per call, host buffers in and out."""
    import numpy as np
    out = {"mode": "one process drives all %d GPUs (kfbo_group_*, KFBO_OPT_GROUP_THREADS = 1); the other ranks of this run are idle" % world}
    qq = np.repeat(np.geomspace(*CFG4_BOUNDS[0], CFG4_GRID), CFG4_GRID)
    rr = np.tile(np.geomspace(*CFG4_BOUNDS[1], CFG4_GRID), CFG4_GRID)
    rv3 = kf.ReadParaVehicle(nsimruns_=TRIALS, cpu_core_number_=1, pnoise_=[Q_TRUTH], onoise_=[R_TRUTH])
    with kf.CostEvaluatorGroup(rv3, SAMPLE_TIMES, n_gpus=world, shard_mode=kf.SHARD_TRIALS, seed=0x4B46424F00002026) as gr:
        ms = wall_evals(lambda: gr.eval(*CANDIDATE), steps, warmup)
        res, path = gr.eval(*CANDIDATE), gr.last_reduce_path()
    steps3 = float(TRIALS) * sum(T for _, T in SAMPLE_TIMES)
    out["cfg3"] = {"ms_per_eval": ms, "value": steps3 / (ms * 1e-3), "unit": UNIT, "sums_meet_by": path,
                   "J_NEES": [float(c) for c in res.cost], "api": "CostEvaluatorGroup.eval over the C ABI kfbo_group_eval (wall clock)"}
    rv4 = kf.ReadParaVehicle(nsimruns_=CFG4_TRIALS, cpu_core_number_=1, pnoise_=[Q_TRUTH], onoise_=[R_TRUTH])
    with kf.CostEvaluatorGroup(rv4, CFG4_SAMPLE_TIMES, n_gpus=world, shard_mode=kf.SHARD_CANDIDATES, seed=0x4B46424F00002026,
                               max_candidates=qq.size) as gr:
        ms4 = wall_evals(lambda: gr.eval_batch_costs(qq, rr), max(steps // 2, 5), warmup)
        path4 = gr.last_reduce_path()
    steps4 = float(CFG4_TRIALS) * qq.size * sum(T for _, T in CFG4_SAMPLE_TIMES)
    out["cfg4"] = {"ms_per_sweep": ms4, "value": steps4 / (ms4 * 1e-3), "unit": UNIT, "sums_meet_by": path4,
                   "api": "CostEvaluatorGroup.eval_batch_costs over kfbo_group_eval_batch_costs (wall clock)"}
    return out


def dense_flops_per_step(N: int) -> int:
    """The reference-faithful dense flop count of one filter-step for an N-state, 1-observation model -- SURVEY.md 8(d)'s
    convention (every mul / add / sub / div of the N x N and 1 x 1 products counts 1, RNG excluded), which gives 137 for N = 2:
    simulator N(2N-1) + 2N + [N(2N-1) + 2N] + (2N-1) + 3; filter [N(2N-1) + 2N] + 2 N^2 (2N-1) + N^2 + N(2N-1) + 2N + (1+N) + 2N
    + 2N + 2N^2 + N^2 (2N-1); metrics N + inverse(N) + N(2N-1) + (2N-1) + 3 + 2 + 8, inverse by cofactors = 8 / 42 / 140."""
    inv = {2: 8, 3: 42, 4: 140}[N]
    sim = N * (2 * N - 1) + 2 * N + (N * (2 * N - 1) + 2 * N) + (2 * N - 1) + 3
    kf_ = (N * (2 * N - 1) + 2 * N) + 2 * N * N * (2 * N - 1) + N * N + N * (2 * N - 1) + 2 * N + (1 + N) + 2 * N + 2 * N + 2 * N * N + N * N * (2 * N - 1)
    met = N + inv + N * (2 * N - 1) + (2 * N - 1) + 3 + 2 + 8
    return sim + kf_ + met


# FP64 instructions the state-augmented kernels EXECUTE per filter-step, noise included (SASS of the step loops,
# profiles/sass_loops_r2q.txt: per thread and step pair / per lane and step pair x 4 lanes), beside the dense reference-style count
AUG_FP64_INSTRUCTIONS = {"thread": {3: 144, 4: 234}, "lanes": {3: 347, 4: 450}}


def leg_augmented(kf, torch, stream, device, fp64_peak):
    """SURVEY.md 8 row f4, north_star's "state-augmented variants": the same evaluation on a chain of N = 3 / 4 integrators
    (kfbo_aug.cuh) -- one trajectory per thread (the default) and, beside it, one covariance row per lane of a 4-lane group.
    The reference has no such variant; reported on its own, with its own flop count, never against the N = 2 figures."""
    import numpy as np
    from oracle import oracle as o
    from kf_bayesopt_b200 import _lib as L
    out = {}
    trials, T, times = 1 << 18, 1000, [(0.1, 1000), (0.5, 1000)]
    steps = float(trials) * T * len(times)
    for N in (3, 4):
        rv = kf.ReadParaVehicle(nsimruns_=trials, cpu_core_number_=1, pnoise_=[Q_TRUTH], onoise_=[R_TRUTH], state_dof_=N)
        flops = dense_flops_per_step(N)
        want = o.aug_eval_philox(N, 0x4B46424F00002026, 5, 0.1, T, Q_TRUTH, R_TRUTH, 0.7, 0.2, 0, 64)
        legs = {}
        with kf.CostEvaluator(rv, times, seed=0x4B46424F00002026, device=device, state_dim=N) as ev:
            ev.set_cuda_stream(stream.cuda_stream)
            for name, layout, kernel in (("thread", L.AUG_LAYOUT_THREAD, f"rollout_aug_thread<{N}>"), ("lanes", L.AUG_LAYOUT_LANES, f"rollout_aug<{N}>")):
                ev.set_option(L.OPT_AUG_LAYOUT, layout)
                ev.set_eval_counter(0)
                ms, kms, launches, res = timed_evals(torch, stream, ev, lambda: ev.eval(*CANDIDATE), 5, 3)
                small = ev.eval_philox_trials(0, 0.7, 0.2, 5, 0, 64)
                ach = flops * steps / (kms * 1e-3) / 1e12
                legs[name] = {"value": steps / (ms * 1e-3), "unit": UNIT, "ms_per_eval": ms, "kernel_ms": kms, "kernel": kernel,
                              "achieved_tflops": ach, "frac_of_fp64_peak": ach / fp64_peak if fp64_peak else None,
                              "fp64_instructions_per_filter_step": AUG_FP64_INSTRUCTIONS[name][N],
                              "J_NEES": [float(c) for c in res.cost],
                              "parity_vs_oracle_max_rel": float(np.max(np.abs(small - want) / np.abs(want)))}
        out[f"N{N}"] = {"state_dim": N, "workload": f"{trials} trials x {T} steps x {len(times)} sample times, on-device Philox noise",
                        "flops_per_filter_step": flops, **legs["thread"], "layout": "one trajectory per thread (KFBO_AUG_LAYOUT_THREAD, the default)",
                        "rows_across_lanes": legs["lanes"]}
    out["note"] = ("builder extension named by north_star, not in the reference (system_models.hpp:11-12 fixes N = 2); parity is against "
                   "oracle/kfbo_oracle_aug.c, which tests/test_oracle_aug.py pins with scipy / closed forms; flop counts by dense_flops_per_step "
                   "(the reference-style dense N x N count, mul/add 1, fma 2, structural zeros included: the per-thread kernel executes far fewer "
                   "instructions than that -- fp64_instructions_per_filter_step -- which is how frac_of_fp64_peak can exceed 1); the two layouts draw "
                   "the same Philox words and give every trajectory the same bits")
    return out


def parity_self_check(kf, torch, dist, device, world, rank):
    """N > 1, before anything is timed, on every rank: trial-sharded and candidate-sharded cost sums -- through the
    peer-memory exchange and through ncclAllReduce -- against THIS rank's own unsharded evaluation of the same Philox
    streams, plus a 300-trial sharded evaluation against the CPU oracle.  Returns the dict for the JSON line; the caller
    exits non-zero when "ok" is false."""
    import numpy as np
    from kf_bayesopt_b200 import _lib as L
    from kf_bayesopt_b200.dist import init_comm
    times, B, C, seed = README_TIMES, 4099, 13, 77                        # neither B nor C divides by the rank count
    rv = kf.ReadParaVehicle(nsimruns_=B, cpu_core_number_=1)
    q, r = np.linspace(0.1, 5.0, C), np.linspace(0.01, 1.0, C)
    with kf.CostEvaluator(rv, times, seed=seed, max_candidates=C, device=device) as ev1:
        want = ev1.eval_batch_raw(q, r)
        want_sums = ev1.last_sums(C).copy()
    out = {"ranks": world, "paths": [], "max_rel": 0.0, "bit_identical_candidates": True, "selection_identical": True}
    for mode, name in ((kf.SHARD_TRIALS, "trials"), (kf.SHARD_CANDIDATES, "candidates")):
        with kf.CostEvaluator(rv, times, seed=seed, max_candidates=C, device=device) as ev:
            peer = init_comm(ev, mode)
            for red in (["peer"] if peer else []) + ["nccl"]:
                ev.set_option(L.OPT_REDUCE, L.REDUCE_AUTO if red == "peer" else L.REDUCE_NCCL)
                ev.set_eval_counter(0)
                got = ev.eval_batch_raw(q, r)
                sums = ev.last_sums(C)
                rel = float(np.max(np.abs(sums - want_sums) / np.abs(want_sums)))
                rec = {"sharded": name, "met_by": ev.last_reduce_path(), "max_rel": rel}
                if mode == kf.SHARD_CANDIDATES:
                    rec["bit_identical"] = bool(np.array_equal(sums, want_sums) and np.array_equal(got["cost"], want["cost"]))
                    out["bit_identical_candidates"] &= rec["bit_identical"]
                out["selection_identical"] &= bool(np.array_equal(got["index_max"], want["index_max"]))
                out["max_rel"] = max(out["max_rel"], rel)
                out["paths"].append(rec)
    # a 300-trial evaluation, trials sharded over the ranks, against the oracle's restatement of the same streams
    def factory(n):
        ev = kf.CostEvaluator(kf.ReadParaVehicle(nsimruns_=n, cpu_core_number_=1), times, seed=seed, device=device)
        init_comm(ev, kf.SHARD_TRIALS)
        return ev
    out["oracle_max_rel"] = oracle_slice_parity(kf, factory, times, 0.7, 0.2, seed, ntrials=300)
    ok = out["max_rel"] <= 1e-12 and out["bit_identical_candidates"] and out["selection_identical"] and out["oracle_max_rel"] <= 1e-9
    flag = torch.tensor([1.0 if ok else 0.0, out["max_rel"], out["oracle_max_rel"]], dtype=torch.float64, device="cuda")
    worst = flag.clone()
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    out["ok"] = bool(flag[0].item() == 1.0)                               # every rank agrees
    out["max_rel"], out["oracle_max_rel"] = float(worst[1].item()), float(worst[2].item())
    out["tolerance"] = {"trials_sharded_vs_unsharded": 1e-12, "candidates_sharded": "bit-identical", "vs_oracle": 1e-9}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="kfbo", choices=["kfbo", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=["cfg3", "cfg4", "cfg5"],
                    help="cfg3 (default, the configuration the metric is quoted on; its line also carries the other configs as "
                         "`extra` legs), the cfg4 candidate sweep as the headline, or the cfg5 Bayesian-optimisation loop (N=1 only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-supplied", action="store_true", help="skip the supplied-noise (HBM roofline) leg")
    ap.add_argument("--no-reduced", action="store_true", help="accepted and ignored (the reduced form was removed in round 2)")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra legs (cfg1/cfg2 latency, cfg4 sweep, NCCL path) and the N>1 parity self-check")
    ap.add_argument("--reduce", default="auto", choices=["auto", "nccl"],
                    help="N > 1: how the cost sums of the ranks meet -- auto = the peer-memory exchange behind the rollout "
                         "kernel when the ranks can map each other's buffers, nccl = one ncclAllReduce")
    ap.add_argument("--peer-fuse", default="auto", choices=["auto", "never", "always"],
                    help="N > 1, peer-memory exchange: in the tail of the rollout kernel or as a kernel behind it (KFBO_OPT_PEER_FUSE)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "kfbo" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.workload == "cfg5":
        if rank == 0:
            run_bo_loop(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import kf_bayesopt_b200 as kf
    from kf_bayesopt_b200 import _lib as L
    from kf_bayesopt_b200.dist import init_comm

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the cost evaluator has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if world != args.gpus and rank == 0:
        print(f"bench.py: warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    line_extra = {}
    # ---- N > 1: multi-GPU correctness first, in the record the driver keeps (a failure ends the run non-zero)
    if world > 1 and not args.no_extra:
        pn = parity_self_check(kf, torch, dist, local_rank, world, rank)
        line_extra["parity_n"] = pn
        if not pn["ok"]:
            if rank == 0:
                emit({"metric": METRIC, "value": None, "unit": UNIT, "n_gpus": world, "parity_n": pn,
                      "error": "multi-GPU parity self-check failed; nothing was timed"})
            dist.barrier()
            dist.destroy_process_group()
            raise SystemExit(3)

    cfg4 = args.workload == "cfg4"
    if cfg4:
        trials, sample_times = CFG4_TRIALS, CFG4_SAMPLE_TIMES
        qq = np.repeat(np.geomspace(*CFG4_BOUNDS[0], CFG4_GRID), CFG4_GRID)
        rr = np.tile(np.geomspace(*CFG4_BOUNDS[1], CFG4_GRID), CFG4_GRID)
        n_cand, shard = qq.size, kf.SHARD_CANDIDATES
        workload = ("cfg4: cost-surface sweep, %d candidates x %d trials x %d sample times (dt 0.1 T 201 / dt 0.5 T 211), "
                    "FP64, on-device Philox4x32-10 noise" % (n_cand, trials, len(sample_times)))
    else:
        trials, sample_times, n_cand, shard, workload = TRIALS, SAMPLE_TIMES, 1, kf.SHARD_TRIALS, WORKLOAD
    rv = kf.ReadParaVehicle(nsimruns_=trials, cpu_core_number_=1, pnoise_=[Q_TRUTH], onoise_=[R_TRUTH])
    ev = kf.CostEvaluator(rv, sample_times, seed=0x4B46424F00002026, device=local_rank, max_candidates=n_cand)
    stream = torch.cuda.Stream()           # the stream every kernel of the evaluator is launched on
    ev.set_cuda_stream(stream.cuda_stream)
    if world > 1:
        init_comm(ev, shard, peer=(args.reduce == "auto"))
        ev.set_option(L.OPT_PEER_FUSE, {"never": L.PEER_FUSE_NEVER, "auto": L.PEER_FUSE_AUTO, "always": L.PEER_FUSE_ALWAYS}[args.peer_fuse])
    # tuning experiments only (tools/multi_round.sh): switch one of the launch-path features off
    for env, opt in (("KFBO_BENCH_GRAPH", L.OPT_CUDA_GRAPH), ("KFBO_BENCH_POLL", L.OPT_HOST_POLL)):
        if env in os.environ:
            ev.set_option(opt, int(os.environ[env]))
    steps_per_eval = float(trials) * n_cand * sum(T for _, T in sample_times)

    class _First:   # the first candidate of a batch, shaped like Result where the JSON line reads it
        def __init__(self, rec):
            self.cost, self.max_cost, self.index_max = rec["cost"][:len(sample_times)], float(rec["max_cost"]), int(rec["index_max"])

    def evaluate():
        if cfg4:
            return _First(ev.eval_batch_raw(qq, rr)[0])
        return ev.eval(*CANDIDATE)

    # ---- device-timed throughput: K evaluations, CUDA events on the launching stream, max over ranks
    for _ in range(args.warmup):
        res = evaluate()
    try:
        gpu_id = "GPU-" + str(torch.cuda.get_device_properties(local_rank).uuid)
    except Exception:
        gpu_id = str(local_rank)
    sampler = ClockSampler(gpu_id)
    sampler.start()
    time.sleep(0.25)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches_per_eval = ev.last_kernel_ms()[1]
    ev.kernel_ms_total(reset=True)           # the kernel times of the loop are accumulated inside the library and read once, after it
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(args.steps):
        res = evaluate()
    e1.record(stream)
    barrier()
    t1 = time.perf_counter()
    kms_total, kms_n = ev.kernel_ms_total()
    assert kms_n == args.steps, (kms_n, args.steps)
    kernel_ms, launches = [kms_total / kms_n], launches_per_eval * args.steps
    reduce_path = ev.last_reduce_path()
    host_us = ev.last_host_us()              # phases of the last timed evaluation on this rank (kfbo_last_host_us)
    clocks = sampler.stop(t0, t1)
    elapsed_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(elapsed_ms, op=dist.ReduceOp.MAX)
    elapsed_ms = float(elapsed_ms.item())
    value = steps_per_eval * args.steps / (elapsed_ms * 1e-3)

    # ---- end to end through the reference-facing interface: "noise" message in, "cost" message out
    # (host buffers; the per-step H2D is the candidate's model constants, the D2H the cost sums)
    rt = kf.RunTrial(rv, evaluator=ev)
    msg = kf.Float64MultiArray.noise(*CANDIDATE)
    e2e_call = (lambda: ev.eval_batch_costs(qq, rr)) if cfg4 else (lambda: rt.ProcessNoiseCallback(msg))
    for _ in range(2):
        e2e_call()
    barrier()
    w0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_call()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - w0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = steps_per_eval * args.steps / float(e2e_s.item())
    D = len(sample_times)
    h2d, d2h = ev.last_transfer_bytes()
    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "ms_per_eval": float(e2e_s.item()) / args.steps * 1e3,
           "api": ("CostEvaluator.eval_batch_costs (host candidate arrays -> host costs) over the C ABI kfbo_eval_batch_costs" if cfg4 else
                   "RunTrial.ProcessNoiseCallback (Float64MultiArray 'noise' -> Float64 'cost') over the C ABI kfbo_eval")}

    # ---- N > 1: the same evaluations with the sums meeting in ONE ncclAllReduce (north_star's wording) instead of the
    # peer-memory exchange -- device-timed like `value`
    if world > 1 and not args.no_extra and args.reduce == "auto":
        ev.set_option(L.OPT_REDUCE, L.REDUCE_NCCL)
        barrier()
        nms, _, _, _ = timed_evals(torch, stream, ev, evaluate, min(args.steps, 10), 3)
        nccl_path = ev.last_reduce_path()
        ev.set_option(L.OPT_REDUCE, L.REDUCE_AUTO)
        t = torch.tensor([nms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        line_extra["nccl_ms"] = float(t.item())
        line_extra["nccl"] = {"ms_per_step": float(t.item()), "value": steps_per_eval / (float(t.item()) * 1e-3), "unit": UNIT, "met_by": nccl_path,
                              "note": "same workload, KFBO_OPT_REDUCE = KFBO_REDUCE_NCCL; the default line's sums met by: " + reduce_path}

    # ---- roofline of the dominant kernel (rollout_philox): FP64 pipe
    k_ms = float(np.mean(kernel_ms))
    local_steps = steps_per_eval / world
    if rank == 0:
        fp64_peak, fp64_uniform = kf.measure_fp64_peak(local_rank, 400.0, both=True)
        achieved = FLOPS_PER_STEP * local_steps / (k_ms * 1e-3) / 1e12
        roofline = {"bound": "fp64", "kernel": "rollout_philox", "achieved": achieved, "peak": fp64_peak,
                    "unit": "TFLOP/s", "frac": achieved / fp64_peak, "traffic": profiled_traffic("rollout_philox", args.workload),
                    "peak_source": "measured live on this GPU: DFMA-chain microbenchmark x = fma(x, y, U), the 2.0-cycle operand form "
                                   "(kfbo_measure_fp64_peak); MEASURED_PEAKS.json has no FP64 figure",
                    "peak_uniform_operands": fp64_uniform,
                    "flops_per_filter_step": FLOPS_PER_STEP, "kernel_ms": k_ms,
                    "filter_steps_per_launch": local_steps,
                    "note": "tensor cores do not apply (2x2 FP64 recursion); RNG flops are excluded from 'achieved'; peak_uniform_operands "
                            "is the chain x = fma(x, U, U) round 1 divided by (2.2 cycles per instruction: not the pipe's peak)" +
                            ("; kernel_ms at N > 1 includes the in-kernel wait for the other ranks' sums when the exchange is fused"
                             if world > 1 else "")}
        line_extra["roofline"] = roofline

    # ---- supplied-noise leg: HBM roofline, inputs resident in HBM and larger than L2
    if rank == 0 and not args.no_supplied:
        Bs, Ts = 1 << 20, 1000                                    # 25 GB of noise >> 126 MB L2; 8192 CTAs = 7 waves
        g = torch.Generator(device="cuda").manual_seed(1)
        x0n = torch.randn((2, Bs), dtype=torch.float64, device="cuda", generator=g)
        w = torch.randn((Ts, 2, Bs), dtype=torch.float64, device="cuda", generator=g) * 0.1
        v = torch.randn((Ts, Bs), dtype=torch.float64, device="cuda", generator=g)
        rvs = kf.ReadParaVehicle(nsimruns_=Bs, cpu_core_number_=1)
        with kf.CostEvaluator(rvs, [(0.1, Ts)], device=local_rank) as evs:
            torch.cuda.synchronize()
            ms_s = []
            for i in range(3 + 5):
                evs.eval_supplied_dev(0, 1.0, 0.1, Bs, x0n.data_ptr(), w.data_ptr(), v.data_ptr())
                if i >= 3:
                    ms_s.append(evs.last_kernel_ms()[0])
        del x0n, w, v
        peaks, src = measured_peaks()
        byts = BYTES_PER_STEP_SUPPLIED * Bs * Ts + BYTES_PER_TRIAL_SUPPLIED * Bs
        ach = byts / (float(np.mean(ms_s)) * 1e-3) / 1e9
        line_extra["roofline_supplied"] = {
            "bound": "hbm", "kernel": "rollout_supplied", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": ach / peaks["hbm_gbs"], "traffic": profiled_traffic("rollout_supplied_bulk", "supplied"), "peak_source": src,
            "kernel_ms": float(np.mean(ms_s)),
            "filter_steps_per_s": Bs * Ts / (float(np.mean(ms_s)) * 1e-3),
            "workload": f"{Bs} trials x {Ts} steps, supplied noise resident in HBM ({byts / 1e9:.2f} GB > L2)"}

    # ---- the other BASELINE.json configs as legs of this line (cfg3 is the headline)
    if not args.no_extra and not cfg4:
        extra = {}
        if world > 1:
            dist.barrier()
        extra["cfg4"] = leg_cfg4(kf, torch, dist, stream, local_rank, world, rank)
        if rank == 0:
            extra.update(leg_default_workload(kf, torch, stream, local_rank, with_cpu=(world == 1 and not args.no_cpu_baseline)))
            extra["state_augmented"] = leg_augmented(kf, torch, stream, local_rank, line_extra.get("roofline", {}).get("peak"))
            if world == 1:
                extra["cfg5"] = leg_cfg5(extra.get("cfg1", {}).get("cpu_baseline", {}).get("latency_ms"))
            if world == 1:     # single GPU: the CUDA path against the oracle on a 300-trial slice of the headline workload's streams
                seed = 0x4B46424F00002026
                extra["parity_vs_oracle"] = {
                    "max_rel": oracle_slice_parity(kf, lambda n: kf.CostEvaluator(kf.ReadParaVehicle(nsimruns_=n, cpu_core_number_=1), sample_times,
                                                                                  seed=seed, device=local_rank), sample_times, 1.0, 0.1, seed),
                    "what": "cost sums of 300 trials x 1000 steps x 2 sample times, CUDA path vs CPU oracle on the same Philox streams", "tolerance": 1e-9}
        line_extra["extra"] = extra

    # ---- N > 1: the single-process mode (the reference's node is one process) on the same GPUs, rank 0 only.  The other
    # ranks wait on the HOST (a key of the rendezvous store, polled with sleeps -- no collective, no kernel on their GPUs),
    # for at most two minutes; a failure of this leg is reported in the line and does not fail the run.
    if world > 1 and not args.no_extra and not cfg4:
        torch.cuda.synchronize()
        dist.barrier()
        store, key = None, "kfbo_one_process_leg_done"
        try:
            store = dist.distributed_c10d._get_default_store()
        except Exception:
            store = None
        if store is not None:
            if rank == 0:
                try:
                    line_extra.setdefault("extra", {})["one_process"] = leg_one_process(kf, world)
                except Exception as e:
                    line_extra.setdefault("extra", {})["one_process"] = {"unavailable": f"{type(e).__name__}: {e}"}
                finally:
                    store.set(key, "1")
            else:
                deadline = time.perf_counter() + 120.0
                while time.perf_counter() < deadline:
                    try:
                        if store.check([key]):
                            break
                    except Exception:
                        break
                    time.sleep(0.02)

    # ---- CPU baseline (rank 0, N=1 only): the oracle's threaded path on a bounded sample
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not cfg4:
        cb, _, _ = cpu_reference_run(steps=3, warmup=1, budget_s=16.0)
        if cb["kind"] == "reference":
            port, _, _ = cpu_reference_run(2, 1, budget_s=6.0, kind="port")
            cb["port"] = {k: port[k] for k in ("value", "unit", "cores", "kind", "sample")}
        line_extra["cpu_baseline"] = cb

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload, "trials": trials, "candidates": n_cand,
                           "sample_times": sample_times,
                           "candidate": "64 x 64 log grid over q in [0.1,5], r in [0.01,1]" if cfg4 else {"pnoise": CANDIDATE[0], "onoise": CANDIDATE[1]},
                           "sharding": (f"{'candidates' if cfg4 else 'trials'} over {world} rank(s); the {n_cand * D * 4} doubles of "
                                        f"cost sums meet by: {reduce_path}") if world > 1 else "single GPU",
                           "l2": "n/a for the Philox path: no HBM-resident inputs (noise is generated on the fly); "
                                 "the supplied-noise leg streams 25 GB >> L2"},
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
                "host_phases_us": {k: round(v, 1) for k, v in host_us.items()},
                "result": {"J_NEES": [float(c) for c in res.cost], "cost": float(res.max_cost), "index_max": res.index_max}}
        line.update(line_extra)
        emit(line)
    ev.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
