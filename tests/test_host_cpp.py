"""CPU checks of the C++ host layer (include/kfbo_trial.hpp, kf_bayesopt_b200/host/): yaml readers,
message builder, topic bus, GP/EI optimiser; and that the host programs build and fail loudly
without a GPU (no CPU fallback)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "kf_bayesopt_b200", "host")
BIN = os.path.join(ROOT, "kf_bayesopt_b200", "bin")

# Same keys, comments and layout as the reference's config/vehicleParams.yaml and config/bayesParams.yaml
VEHICLE_YAML = """CPUcoreNumber: 10

#state and observation
stateDOF: 2
observationDOF: 1 #2

#time
t_start: 0.0 #start of the simulation time
t_end: 20.1 #end of the simulation time
dt: 0.1

#initial state
xi0: 0.0
xidot0: 0.0

Nsimruns: 200

pnoise: [1.0]
onoise: [0.1]

#optimize "processNoise" or "observationNoise" , "all"
optimizationChoice: all

#JNIS or JNESS or CNISor CNEES
costChoice: JNEES"""

BAYES_YAML = """surr_name: sGaussianProcess #sStudentTProcessNIG sGaussianProcess
kernel_name: kMaternARD3 #normally use kMaternARD5
crit_name: cEI

n_init_samples: 20 #initial sample numbers
n_iterations: 100 #total iterations that need to be done
random_seed: -1 #random_seed
verbose_level: 1
noise: 0.00000001 #1e-7 seems it does't accept this way
force_jump: 10
epsilon: 0.2

lower_bound: [0.1, 0.01] #default for 3d process noise[0.000001, 0.000001, 0.000001]
upper_bound: [5.0, 1.0]  #[1.0, 1.0, 1.0]

n_iteration_relearn: 10 #Number of iterations between re-learning kernel parameters
n_inner_evaluation: 500
log_filename: bayesopt.log
"""


@pytest.fixture(scope="module")
def yamls(tmp_path_factory):
    d = tmp_path_factory.mktemp("cfg")
    v, b = d / "vehicleParams.yaml", d / "bayesParams.yaml"
    v.write_text(VEHICLE_YAML)
    b.write_text(BAYES_YAML)
    return str(v), str(b)


@pytest.fixture(scope="module")
def host_bins(kfbo):
    subprocess.check_call(["make", "-C", HOST, "-s"])
    return BIN


def test_host_layer_cpu_checks(tmp_path, yamls, kfbo):
    exe = tmp_path / "host_check"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "include"), "-I", HOST,
                           os.path.join(ROOT, "tests", "native", "host_check.cpp"), "-o", str(exe),
                           "-L", os.path.join(ROOT, "kf_bayesopt_b200"), "-lkfbo",
                           "-Wl,-rpath," + os.path.join(ROOT, "kf_bayesopt_b200")])
    out = subprocess.run([str(exe), *yamls], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    for name in ("read_para_vehicle", "read_para_bayes", "noise_message", "topic_bus", "gp_ei"):
        assert f"ok {name}" in out.stdout


def test_ros_stand_in_bus_mode_ping_pong(tmp_path):
    # the two-node loop of tests/native/boloop rests on it: two threads, publish -> the other node's callback -> answer
    exe = tmp_path / "ros_bus_check"
    subprocess.check_call(["g++", "-O2", "-std=c++14", "-pthread", "-I", os.path.join(ROOT, "oracle", "ref_shim"),
                           os.path.join(ROOT, "tests", "native", "ros_bus_check.cpp"), "-o", str(exe)])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "ok ros_bus" in out.stdout, out.stdout + out.stderr


def test_python_and_cpp_yaml_readers_agree(yamls, kfbo):
    rv = kfbo.ReadParaVehicle.from_yaml(yamls[0])
    assert (rv.cpu_core_number_, rv.nsimruns_, rv.max_iterations_, rv.cost_choice_) == (10, 200, 201, "JNEES")
    assert rv.pnoise_ == [1.0] and rv.onoise_ == [0.1] and rv.optimization_choice_ == "all"


@pytest.mark.skipif(shutil.which("nvidia-smi") is not None, reason="checks the no-GPU failure mode")
def test_host_programs_fail_loudly_without_a_gpu(host_bins, yamls):
    for cmd in (["kfbo_test_trial", "--vehicle", yamls[0]], ["kfbo_bo_loop", "--vehicle", yamls[0], "--bayes", yamls[1]]):
        out = subprocess.run([os.path.join(host_bins, cmd[0]), *cmd[1:]], capture_output=True, text=True)
        assert out.returncode != 0
        assert "no CPU fallback" in out.stderr
    out = subprocess.run([os.path.join(host_bins, "kfbo_robot1d_kf"), "--vehicle", yamls[0]], input="noise 1 1 1.0 0.1\n",
                         capture_output=True, text=True)
    assert out.returncode != 0 and "no CPU fallback" in out.stderr
