"""Host-side C++ facade (agri-fly_b200/host/agf_quadcopter.hpp): the reference's Components/Simulation object
API on top of the C ABI.  CPU: it compiles stand-alone and against the UNMODIFIED reference headers, links with
the library, and fails loudly without a GPU.  GPU: a Rappids_Simulator-style loop written against the object API
(examples/rappids_loop.cpp: Run(); timer.Advance(); SetCommandRadioMsg from a delay queue) reproduces the
oracle's trajectory bit for bit -- which also pins agf_batch_advance_clock + agf_batch_run(b, 0, 1) to the fused
agf_batch_run(b, dt, n) path and to the reference's Timer semantics."""
import os
import subprocess

import numpy as np
import pytest

from common import bit_equal, run_oracle
from conftest import ROOT

EX = os.path.join(ROOT, "examples", "rappids_loop.cpp")
BIN = os.path.join(ROOT, "examples", "rappids_loop")
INC = ["-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "agri-fly_b200", "host")]
REF = "/root/reference"


def build_example():
    lib_dir = os.path.join(ROOT, "agri-fly_b200")
    cmd = ["g++", "-std=c++14", "-O1", "-Wall", "-Werror"] + INC + [EX, "-L" + lib_dir, "-lagrifly_b200",
                                                                      "-Wl,-rpath," + lib_dir, "-o", BIN]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return BIN


def test_facade_compiles_and_links(agf):
    build_example()
    assert os.path.exists(BIN)


def test_facade_compiles_against_unmodified_reference_headers(agf):
    if not os.path.isdir(os.path.join(REF, "Common")):
        pytest.skip("reference tree not present on this machine")
    cmd = ["g++", "-std=c++14", "-fsyntax-only", "-DAGF_WITH_REFERENCE_HEADERS", "-I" + os.path.join(ROOT, "oracle", "shim"),
           "-I" + os.path.join(REF, "Common"), "-I" + os.path.join(REF, "Components")] + INC + [EX]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_facade_has_no_cpu_fallback(agf):
    from conftest import has_cuda
    if has_cuda():
        pytest.skip("a CUDA device is present")
    build_example()
    r = subprocess.run([BIN, "10"], capture_output=True, text=True)
    assert r.returncode != 0
    assert "agf_batch_create" in r.stderr or "CUDA" in r.stderr or "device" in r.stderr


@pytest.mark.gpu
def test_object_api_loop_matches_oracle(agf, checker_shared):
    build_example()
    sc = agf.scenarios.rates_scenario(agf.codec)
    r = subprocess.run([BIN, str(sc["nticks"])], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    rows = np.array([[float(x) for x in ln.split()] for ln in r.stdout.strip().splitlines()])
    assert len(rows) == 5
    ref, _ = run_oracle(checker_shared, agf, sc)
    for row in rows:
        k = int(row[0])
        assert bit_equal(row[1:18], ref[k, 0:17]), (k, row[1:18] - ref[k, 0:17])


FLEET = os.path.join(ROOT, "examples", "rappids_fleet.cpp")
FLEET_BIN = os.path.join(ROOT, "examples", "rappids_fleet")


def build_fleet():
    lib_dir = os.path.join(ROOT, "agri-fly_b200")
    r = subprocess.run(["g++", "-std=c++14", "-O1", "-Wall", "-Werror", INC[0], FLEET, "-L" + lib_dir, "-lagrifly_b200",
                        "-Wl,-rpath," + lib_dir, "-o", FLEET_BIN], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return FLEET_BIN


def test_fleet_example_compiles_against_the_c_abi_only(agf):
    build_fleet()
    assert os.path.exists(FLEET_BIN)


@pytest.mark.gpu
def test_fleet_example_writes_the_reference_csv(agf, orc_mod):
    """examples/rappids_fleet.cpp: a fleet flying the ROS node's flight stages with the mocap estimator, the whole offboard
    loop inside the kernel, vehicle 0 written as Rappids_Simulator's simulation.csv.  The rows equal what the oracle
    (reference controller + estimator in the same loop) gives for the same scenario, to the 6 digits the file carries."""
    from common import cfg_for
    build_fleet()
    r = subprocess.run([FLEET_BIN, "513", "6", "0"], capture_output=True, text=True, timeout=600)  # noise-free for the comparison
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    assert lines[0] + "\n" == agf.csv_header() and len(lines) == 601
    rows = np.array([[float(x) for x in ln.rstrip(",").split(",")] for ln in lines[1:]])
    assert rows.shape == (600, 40)
    if not orc_mod.available("port-shared"):
        pytest.skip("oracle port not built")
    O = orc_mod.Oracle("port-shared")
    sc = agf.scenarios.stages_scenario(3, nticks=3000)
    v = O.vehicle(agf.vehicle_cfg(None, 1, motor_time_const=0.015), uwb_comm_period=0.0)  # as the example builds it
    v.set_offboard_estimator(agf.offboard_estimator())
    ref = agf.offboard_ref(kind=1, start_us=500000, stop_us=3000000, desired_pos=(0.0, 0.0, 1.0), desired_yaw=0.0, traj_id=3)
    tr = v.run_offboard_ref(3000, agf.offboard_cfg(sc["quad_type"]), ref)
    want = tr[4::5]  # state after ticks 5, 10, ...
    np.testing.assert_allclose(rows[:, 1:7], want[:, 0:6], rtol=2e-5, atol=1e-9)     # position, velocity: 6 significant digits
    np.testing.assert_allclose(rows[:, 10:13], want[:, 10:13], rtol=2e-5, atol=1e-9)  # angular velocity
    assert rows[:, 3].max() > 0.9 and rows[-1, 3] < 0.05 and np.all(rows[:, 35] == 0)   # took off, landed, no panic
    assert np.max(np.abs(rows[:, 17:20] - rows[:, 1:4])) < 0.1                           # estimate follows the truth


MC = os.path.join(ROOT, "examples", "monte_carlo_fleet.cpp")
MC_BIN = os.path.join(ROOT, "examples", "monte_carlo_fleet")


def build_mc():
    lib_dir = os.path.join(ROOT, "agri-fly_b200")
    r = subprocess.run(["g++", "-std=c++14", "-O1", "-Wall", "-Werror", "-pthread", INC[0], MC, "-L" + lib_dir, "-lagrifly_b200",
                        "-Wl,-rpath," + lib_dir, "-o", MC_BIN], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return MC_BIN


def test_monte_carlo_fleet_example_compiles_against_the_c_abi_only(agf):
    """examples/monte_carlo_fleet.cpp: the multi-GPU host path in C++ -- one thread and one batch per GPU, the statistics
    read-out over NCCL through agf_batch_reduce_stats_nccl; the program links the C ABI only (no NCCL, no CUDA runtime)."""
    build_mc()
    needed = subprocess.check_output(["readelf", "-d", MC_BIN], text=True)
    assert "libnccl" not in needed and "libcudart" not in needed and "libagrifly_b200" in needed


@pytest.mark.gpu
def test_monte_carlo_fleet_sharded_equals_unsharded(agf):
    """The C3 population stepped by the C++ example: sharded over the GPUs of the node (one host thread per GPU, statistics
    all-gathered over NCCL inside the C ABI) against the same population in one batch.  Noise streams are keyed by the
    population-wide vehicle index, so every vehicle flies the same trajectory either way: the combined statistics agree
    (sums to rounding -- the order of the additions differs --, maxima and counts exactly).  With one visible GPU the
    sharded run degenerates to a single rank (the NCCL path is then covered by the 2-GPU bench runs under profiles/)."""
    import torch
    build_mc()
    g = min(2, torch.cuda.device_count())
    n = 8192

    def run(gpus, per):
        r = subprocess.run([MC_BIN, str(gpus), str(per), "5", "fp32"], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr + r.stdout
        line = [ln for ln in r.stdout.splitlines() if ln.startswith("STATS")][-1].split()
        return [float(x) for x in line[1:]], r.stdout

    whole, out = run(1, n)
    parts, out2 = run(g, n // g)
    print(out2)
    assert whole[0] == n and parts[0] == n
    assert whole[3] == parts[3] and whole[4] == parts[4] and whole[2] == parts[2]   # panics, non-finite, max error
    assert abs(whole[5] / parts[5] - 1) < 1e-12                                       # sum of squared errors
    assert whole[1] < 0.5 and whole[3] < 0.02 * n                                     # the population tracks its square
