"""CPU tests of the product's host side: libkfbo.so loads, exports exactly what include/kfbo.h
declares, and its host-only pieces (discretisation, control table, cost assembly, sharding) agree
with the oracle and the golden vectors.  No compute calls are made here (no GPU in this tier)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, rel_err


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "kfbo.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(kfbo_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(kfbo):
    from kf_bayesopt_b200 import _lib
    lib = _lib.load()
    declared = _declared_functions()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/kfbo.h but not exported"
    assert sorted(_lib.SIGNATURES) == declared, "ctypes table and header disagree"
    assert b"sm_100a" in lib.kfbo_version()


def test_python_constants_match_the_header_enums(kfbo):
    # every KFBO_OPT_* / KFBO_REDUCE_* / KFBO_PEER_FUSE_* / KFBO_AUG_LAYOUT_* / KFBO_SHARD_* value the ctypes layer carries is the header's
    from kf_bayesopt_b200 import _lib
    src = open(os.path.join(ROOT, "include", "kfbo.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    header = {m.group(1): int(m.group(2)) for m in re.finditer(r"\bKFBO_([A-Z0-9_]+)\s*=\s*(-?\d+)", src)}
    checked = 0
    for name, value in vars(_lib).items():
        if isinstance(value, int) and not isinstance(value, bool) and name.isupper() and name in header:
            assert header[name] == value, f"_lib.{name} = {value}, kfbo.h says KFBO_{name} = {header[name]}"
            checked += 1
    opts = [n for n in header if n.startswith("OPT_")]
    assert checked >= 15 and all(hasattr(_lib, n) for n in opts), [n for n in opts if not hasattr(_lib, n)]


def test_numpy_result_record_matches_the_struct(kfbo):
    import ctypes
    from kf_bayesopt_b200 import _lib, evaluator
    assert evaluator.RESULT_DTYPE.itemsize == ctypes.sizeof(_lib.KfboResult)
    for name, _ in _lib.KfboResult._fields_:
        assert evaluator.RESULT_DTYPE.fields[name][1] == getattr(_lib.KfboResult, name).offset


def test_struct_layout_matches_header(kfbo):
    from kf_bayesopt_b200 import _lib
    # sizes implied by include/kfbo.h with natural alignment
    assert C.sizeof(_lib.KfboParams) == 4 + 4 + 32 + 16 + 4 * 4 + 16 + 4 + 4 + 8 + 4 + 4
    assert C.sizeof(_lib.KfboResult) == 5 * 32 + 8 + 4 + 4
    p = _lib.KfboParams()
    _lib.load().kfbo_default_params(C.byref(p))
    # config/vehicleParams.yaml + trial_node.cpp:105-114
    assert (p.n_sample_times, p.nsimruns, p.cpu_core_number, p.state_dof, p.observation_dof) == (2, 200, 10, 2, 1)
    assert list(p.dt)[:2] == [0.1, 1.0] and list(p.max_iterations)[:2] == [201, 201]
    assert (p.pnoise_truth, p.onoise_truth, p.cost_choice) == (1.0, 0.1, kfbo.JNEES)


def test_host_discretisation_matches_oracle_and_golden(kfbo, oracle, golden):
    for (q, dt), Fd, Qd in zip(golden["disc_in"], golden["disc_Fd"], golden["disc_Qd"]):
        F, Q = kfbo.discretize(q, dt)
        assert np.array_equal(F, Fd)
        assert rel_err(Q, Qd) < 4 * np.finfo(float).eps
        Fo, Qo = oracle.discretize(q, dt)
        assert np.array_equal(F, Fo) and rel_err(Q, Qo) < 4 * np.finfo(float).eps


def test_host_control_table_matches_oracle(kfbo, oracle, golden):
    for dt in (0.1, 0.5, 1.0):
        assert np.array_equal(kfbo.control_table(dt), golden[f"u_{dt}"])
        assert np.array_equal(kfbo.control_table(dt), oracle.control_table(dt))


@pytest.mark.parametrize("choice", ["JNEES", "CNEES", "JNIS", "CNIS"])
def test_finalize_matches_oracle_cost_assembly(kfbo, oracle, golden, choice):
    B, cores = int(golden["meta"][1]), int(golden["meta"][2])
    rv = kfbo.ReadParaVehicle(nsimruns_=B, cpu_core_number_=cores, cost_choice_=choice)
    st = [(float(dt), int(T)) for dt, T in golden["sample_times"]]
    params = kfbo.make_params(rv, st)
    for c in range(len(golden["candidates"])):
        sums = np.stack([golden[f"per_trial_{d}_{c}"].sum(axis=1) for d in range(3)])
        res = kfbo.finalize(params, sums)
        costs = []
        for d, (dt, T) in enumerate(st):
            want = oracle.cost_from_per_trial(golden[f"per_trial_{d}_{c}"], T, B, cores, cost_choice=choice)
            got = [res.average_nees[d], res.average_nis[d], res.average_var_nees[d], res.average_var_nis[d]]
            assert rel_err(got, want[:4]) < 1e-13
            assert res.cost[d] == pytest.approx(want[4], rel=1e-11, abs=1e-13)
            costs.append(res.cost[d])
        m, i = oracle.max_cost(costs)
        assert res.index_max == i and res.max_cost == m     # selection is bit-identical given the costs


def test_finalize_integer_semantics_when_trials_do_not_divide(kfbo, oracle):
    # trial.cpp:19,28 with nsimruns % cores != 0: 5 trials/thread run, k_average = (T*23)/4
    rng = np.random.default_rng(1)
    T, nsim, cores = 201, 23, 4
    rv = kfbo.ReadParaVehicle(nsimruns_=nsim, cpu_core_number_=cores)
    params = kfbo.make_params(rv, [(0.1, T)])
    assert kfbo._lib.load().kfbo_effective_trials(C.byref(params)) == 20
    pt = rng.uniform(1, 3, size=(4, 20))
    res = kfbo.finalize(params, pt.sum(axis=1)[None])
    want = oracle.cost_from_per_trial(pt, T, nsim, cores)
    assert rel_err([res.average_nees[0], res.average_nis[0], res.average_var_nees[0], res.average_var_nis[0]], want[:4]) < 1e-13
    assert res.cost[0] == pytest.approx(want[4], rel=1e-12)


def test_finalize_tie_takes_first_index(kfbo):
    params = kfbo.make_params(kfbo.ReadParaVehicle(), [(0.1, 201), (1.0, 201)])
    s = np.array([[200 * 201 * 3.0, 1, 1, 1]] * 2)
    res = kfbo.finalize(params, s)
    assert res.cost[0] == res.cost[1] and res.index_max == 0


def test_shard_range_partitions(kfbo):
    for n, g in [(1 << 20, 8), (4096, 8), (10, 4), (3, 8), (0, 2)]:
        spans = [kfbo.shard_range(n, r, g) for r in range(g)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(g - 1))
        sizes = [e - b for b, e in spans]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(kfbo.KfboError):
        kfbo.shard_range(10, 4, 4)


def test_invalid_params_are_rejected_before_touching_the_device(kfbo):
    from kf_bayesopt_b200 import _lib
    lib = _lib.load()
    for mutate, code in [
        (lambda p: setattr(p, "n_sample_times", 5), -2),
        (lambda p: p.max_iterations.__setitem__(0, 2001), -2),     # control table limit (system_models.hpp:51)
        (lambda p: p.max_iterations.__setitem__(1, 1), -2),
        (lambda p: setattr(p, "cpu_core_number", 0), -1),
        (lambda p: setattr(p, "nsimruns", 5), -1),                 # fewer trials than threads: nothing to run
        (lambda p: setattr(p, "cost_choice", 7), -1),
        (lambda p: p.dt.__setitem__(0, 0.0), -1),
        (lambda p: setattr(p, "max_candidates", 0), -1),
    ]:
        p = _lib.KfboParams()
        lib.kfbo_default_params(C.byref(p))
        mutate(p)
        h = C.c_void_p()
        assert lib.kfbo_create(C.byref(p), C.byref(h)) == code
        assert h.value is None and lib.kfbo_last_error(None)
    assert lib.kfbo_create(None, None) == -1


def test_no_cpu_fallback(kfbo):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(kfbo.KfboError) as e:
        kfbo.CostEvaluator()
    assert e.value.code == -3 and "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "kf_bayesopt_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")) or f == "Makefile":
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower().replace("draw for draw by oracle/kfbo_oracle.c", ""), f


def test_read_para_vehicle_yaml_roundtrip(kfbo, tmp_path):
    y = tmp_path / "vehicleParams.yaml"
    y.write_text("CPUcoreNumber: 10\nstateDOF: 2\nobservationDOF: 1 #2\nt_start: 0.0\nt_end: 20.1\ndt: 0.1\n"
                 "xi0: 0.0\nxidot0: 0.0\nNsimruns: 200\npnoise: [1.0]\nonoise: [0.1]\n"
                 "optimizationChoice: all\ncostChoice: JNEES")
    rv = kfbo.ReadParaVehicle.from_yaml(str(y))
    assert rv.max_iterations_ == 201          # int(20.1/0.1) = int(201.00000000000003)
    assert (rv.nsimruns_, rv.cpu_core_number_, rv.pnoise_, rv.onoise_) == (200, 10, [1.0], [0.1])
    assert kfbo.sample_times_from(rv) == [(0.1, 201), (1.0, 201)]
    assert kfbo.sample_times_from(rv, kfbo.README_SECOND_SAMPLE_TIME) == [(0.1, 201), (0.5, 211)]
    assert kfbo.sample_times_from(rv, None) == [(0.1, 201)]


def test_header_is_plain_c_and_the_library_links_from_c(kfbo, tmp_path):
    # the boundary is a C ABI: include/kfbo.h compiles as C99 (-pedantic) and a C program links against libkfbo.so and
    # calls host-only entry points (no device needed: defaults, sharding arithmetic, the no-fallback error of kfbo_create)
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "abi.c"
    src.write_text(r'''
#include <stdio.h>
#include "kfbo.h"
int main(void)
{
    kfbo_params p;
    kfbo_handle *h = 0;
    int64_t b = -1, e = -1;
    kfbo_default_params(&p);
    if (p.n_sample_times != 2 || p.nsimruns != 200 || p.cpu_core_number != 10) return 1;
    if (kfbo_shard_range(10, 1, 4, &b, &e) != KFBO_OK || b != 3 || e != 6) return 2;
    if (kfbo_effective_trials(&p) != 200) return 3;
    printf("%s rc=%d\n", kfbo_version(), kfbo_create(&p, &h));
    if (h) kfbo_destroy(h);
    return 0;
}
''')
    exe = tmp_path / "abi"
    libdir = os.path.dirname(kfbo.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), str(src),
                           "-o", str(exe), "-L", libdir, "-lkfbo", f"-Wl,-rpath,{libdir}"])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "sm_100a" in out.stdout
    import torch
    if not torch.cuda.is_available():
        assert f"rc={-3}" in out.stdout                          # KFBO_ENODEVICE: no CPU fallback
