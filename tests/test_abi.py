"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol the header declares,
and its host-side pieces (airframe tables, radio/telemetry codecs, shared libm) match the reference."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, has_cuda

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors.npz"))


def header_functions():
    names = set()
    for h in ("agrifly_b200.h", "agrifly_b200_rappids.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(agf_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_library_exports_every_declared_symbol(agf):
    names = header_functions()
    assert len(names) >= 35
    L = C.CDLL(agf.LIB_PATH)
    for n in names:
        assert hasattr(L, n), "missing export " + n
    assert set(names) == set(agf.abi.PROTOTYPES), set(names) ^ set(agf.abi.PROTOTYPES)


def test_struct_sizes_match_header(agf, tmp_path):
    """ctypes mirrors have the layout the C compiler gives the header's structs."""
    import subprocess
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "agrifly_b200.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(agf_logic_consts),sizeof(agf_vehicle_cfg),sizeof(agf_cmd_entry),sizeof(agf_batch_opts),'
                   'sizeof(agf_telemetry),sizeof(agf_offboard_cfg),sizeof(agf_offboard_target),sizeof(agf_offboard_ref),'
                   'sizeof(agf_offboard_estimator),sizeof(agf_csv_record),sizeof(agf_msg_telemetry),'
                   'sizeof(agf_msg_simulator_truth));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I" + os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    A = agf.abi
    assert sizes == [C.sizeof(A.LogicConsts), C.sizeof(A.VehicleCfg), C.sizeof(A.CmdEntry), C.sizeof(A.BatchOpts),
                     C.sizeof(A.Telemetry), C.sizeof(A.OffboardCfg), C.sizeof(A.OffboardTarget), C.sizeof(A.OffboardRef),
                     C.sizeof(A.OffboardEstimator), C.sizeof(A.CsvRecord), C.sizeof(A.MsgTelemetry), C.sizeof(A.MsgSimulatorTruth)]
    assert agf.lib().agf_field_size(0) == 24 and agf.lib().agf_field_size(15) == 324


@pytest.mark.parametrize("qt", [1, 2, 4, 5])
def test_airframe_tables_match_reference(agf, qt):
    lc = agf.abi.LogicConsts()
    assert agf.lib().agf_logic_consts_from_type(qt, C.byref(lc)) == 0
    assert np.array_equal(np.frombuffer(bytes(lc), np.uint8), GOLD["consts/%d" % qt])


def test_vehicle_id_map(agf):
    m = {3: 1, 4: 1, 10: 1, 2: 2, 17: 2, 13: 4, 19: 4, 1: 5, 26: 5, 0: 0, 8: 0, 255: 0}
    for vid, t in m.items():
        assert agf.quad_type_from_id(vid) == t


def test_cfg_widening_matches_rappids_main(agf):
    c = agf.vehicle_cfg(vehicle_id=1)
    assert c.mass == float(np.float32(0.142)) and c.inertia[4] == c.inertia[0] == float(np.float32(92.7e-6))
    assert c.prop_torque_from_speed_sqr == float(np.float32(0.00808) * np.float32(4.32e-8))
    assert c.motor_max_speed == float(np.float32(2000 - 999) / np.float32(0.14))
    assert c.motor_time_const == 0 and c.motor_inertia == 0


@pytest.mark.parametrize("kind", ["rates", "position", "acceleration"])
def test_radio_codec_matches_reference(agf, kind):
    vals = GOLD["codec/values"]
    for row, raw_ref, dec_ref in zip(vals, GOLD["codec/%s/raw" % kind], GOLD["codec/%s/decoded" % kind]):
        if kind == "rates":
            raw = agf.codec.encode_rates(3, row[0], row[1:4])
            used = 3 + 2 * 4
        elif kind == "position":
            raw = agf.codec.encode_position(1, row[0:3], row[3:6], row[6:9])
            used = 3 + 2 * 9
        else:
            raw = agf.codec.encode_acceleration(2, row[0:3], row[3])
            used = 3 + 2 * 4
        got = np.frombuffer(raw, np.uint8)
        # bytes the reference's Create*Command writes (the rest of its RawMessage is uninitialised)
        assert np.array_equal(np.delete(got[:used], 1), np.delete(raw_ref[:used], 1))
        t, f, fl = agf.codec.decode(raw_ref.tobytes())
        n = {"rates": 10, "position": 9, "acceleration": 4}[kind]
        assert np.array_equal(fl[:n], dec_ref[:n], equal_nan=True)


def test_radio_quantisation_known_answer(agf):
    """2.0 m encodes to 2.00012 m; saturation and NaN handling."""
    t, f, fl = agf.codec.decode(agf.codec.encode_position(0, (2.0, 25.0, -25.0), (np.nan, 0, 0), (0, 0, 0)))
    assert t == 3 and abs(fl[0] - 2.0001220703125) < 1e-7
    assert fl[1] == np.float32(20.0 * 32767 / 32768) and fl[2] == -20.0 and fl[3] == -10.0


def test_telemetry_decode_matches_reference(agf, orc_mod):
    for key in ("ref-glibc/full", "ref-glibc/rates"):
        for pk in ("tel1", "tel2"):
            t = agf.codec.decode_telemetry(GOLD["%s/%s" % (key, pk)].tobytes())
            assert t.type == (0 if pk == "tel1" else 1)
            if pk == "tel1":
                est = GOLD[key + "/full/kf_pos"]
                assert np.allclose(np.array(t.position), est, atol=30.0 / 32767 * 1.01)
            else:
                assert t.panic_reason == GOLD[key + "/full/first_panic_reason"]


def test_no_cpu_fallback(agf):
    if has_cuda():
        pytest.skip("a CUDA device is present")
    with pytest.raises(agf.AgfError) as e:
        agf.Batch(agf.vehicle_cfg(vehicle_id=1), 8)
    assert e.value.code == agf.abi.ENODEVICE


def test_product_does_not_reference_oracle():
    """Nothing under agri-fly_b200/ or include/ may import, include or link oracle/."""
    bad = []
    for base in ("agri-fly_b200", "include", "agrifly_b200"):
        for dp, dn, fn in os.walk(os.path.join(ROOT, base)):
            if "build" in dp or "__pycache__" in dp:
                continue
            for f in fn:
                if f.endswith((".so", ".o", ".sha", ".pyc")):
                    continue
                txt = open(os.path.join(dp, f), errors="ignore").read()
                for m in re.finditer(r'(#include\s*"[^"]*oracle[^"]*"|import\s+orc\b|from\s+orc\b|libagf_(port|ref|hostsim))', txt):
                    bad.append((os.path.join(dp, f), m.group(0)))
    assert not bad, bad


# ---- shared deterministic libm -----------------------------------------------------------------
def _math_lib(tmp_path_factory):
    d = tmp_path_factory.mktemp("agfmath")
    src = d / "m.c"
    src.write_text('#include "%s"\n' % os.path.join(ROOT, "agri-fly_b200", "csrc", "agf_math.h") +
                   "double w_sin(double x){return agf_sin(x);} double w_cos(double x){return agf_cos(x);}\n"
                   "double w_asin(double x){return agf_asin(x);} double w_acos(double x){return agf_acos(x);}\n"
                   "double w_atan2(double y,double x){return agf_atan2(y,x);}\n"
                   "float w_sinf(float x){return agf_sinf(x);} float w_acosf(float x){return agf_acosf(x);}\n")
    so = d / "m.so"
    import subprocess
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", str(src), "-o", str(so), "-lm"])
    L = C.CDLL(str(so))
    for n in ("w_sin", "w_cos", "w_asin", "w_acos"):
        getattr(L, n).restype = C.c_double
        getattr(L, n).argtypes = [C.c_double]
    L.w_atan2.restype = C.c_double
    L.w_atan2.argtypes = [C.c_double, C.c_double]
    L.w_sinf.restype = C.c_float
    L.w_sinf.argtypes = [C.c_float]
    L.w_acosf.restype = C.c_float
    L.w_acosf.argtypes = [C.c_float]
    return L


def test_shared_libm_accuracy(tmp_path_factory):
    L = _math_lib(tmp_path_factory)
    rng = np.random.default_rng(0)

    def ulps(f, ref, xs):
        got = np.array([f(float(x)) for x in xs])
        r = ref(xs)
        return np.nanmax(np.abs(got - r) / np.spacing(np.abs(r)))

    xs = np.concatenate([rng.uniform(-1, 1, 4000), rng.uniform(-100, 100, 4000), rng.uniform(-1e5, 1e5, 4000)])
    assert ulps(L.w_sin, np.sin, xs) <= 2 and ulps(L.w_cos, np.cos, xs) <= 2
    xa = np.concatenate([rng.uniform(-1, 1, 8000), 1 - 10 ** rng.uniform(-12, 0, 1000), -1 + 10 ** rng.uniform(-12, 0, 1000)])
    assert ulps(L.w_asin, np.arcsin, xa) <= 2 and ulps(L.w_acos, np.arccos, xa) <= 2
    y, x = rng.normal(size=4000) * 10 ** rng.uniform(-3, 3, 4000), rng.normal(size=4000) * 10 ** rng.uniform(-3, 3, 4000)
    got = np.array([L.w_atan2(float(a), float(b)) for a, b in zip(y, x)])
    assert np.max(np.abs(got - np.arctan2(y, x)) / np.spacing(np.abs(np.arctan2(y, x)))) <= 2
    assert L.w_sin(0.0) == 0 and L.w_cos(0.0) == 1 and L.w_acos(1.0) == 0 and np.isnan(L.w_acos(1.5))
    assert L.w_atan2(0.0, -1.0) == np.pi and L.w_atan2(1.0, 0.0) == np.pi / 2
    # float entry points: correctly rounded from the binary64 evaluation
    xf = rng.uniform(-3, 3, 2000).astype(np.float32)
    got = np.array([L.w_sinf(float(v)) for v in xf], np.float32)
    assert np.max(np.abs(got.astype(np.float64) - np.sin(xf.astype(np.float64)))) < 6e-8


def test_nccl_is_bound_at_run_time_not_link_time(agf):
    """The statistics collective lives in the C ABI (agf_batch_reduce_stats_nccl); libnccl.so.2 is resolved with dlopen on
    first use, so the library loads on hosts without NCCL and has no link-time dependency on it."""
    import subprocess
    from agrifly_b200 import LIB_PATH
    needed = subprocess.check_output(["readelf", "-d", LIB_PATH], text=True)
    assert "libnccl" not in needed
    L = agf.lib()
    v = C.c_int(0)
    rc = L.agf_nccl_version(C.byref(v))
    if rc == agf.abi.ENCCL:
        pytest.skip("no libnccl.so.2 on this host: " + L.agf_last_error_string().decode())
    assert rc == 0 and v.value >= 20400
    uid = (C.c_uint8 * agf.abi.NCCL_UNIQUE_ID_BYTES)()
    assert L.agf_nccl_get_unique_id(uid) == 0 and any(bytes(uid))
    # argument checking happens before any device work
    comm = C.c_void_p()
    assert L.agf_nccl_comm_init_rank(uid, 2, 5, 0, C.byref(comm)) == agf.abi.EINVAL
    assert L.agf_batch_reduce_stats_nccl(None, None, None, None) == agf.abi.EINVAL


def test_planner_dispatch_entry_points_check_their_arguments_before_any_device_work(agf):
    """agf_rappids_set_dispatch / agf_rappids_get_plan_work (the work-ordered dispatch of the planning pass): argument errors are
    reported without a device."""
    L = agf.lib()
    assert L.agf_rappids_set_dispatch(None, 1) == agf.abi.EINVAL
    assert L.agf_rappids_get_plan_work(None, None, 0, 0) == agf.abi.EINVAL
    assert L.agf_rappids_plan(None) == agf.abi.EINVAL
    assert b"null" in L.agf_last_error_string()
