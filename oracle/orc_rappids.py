"""oracle/orc_rappids.py -- TEST INFRASTRUCTURE: ctypes driver for the CPU oracles of the RAPPIDS
planner path (oracle/rappids_api.h).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs import this.  Flavours: ref-glibc /
ref-shared (unmodified reference sources, oracle/_ref/libagf_rappids_ref_*.so) and port-glibc /
port-shared (independent restatement, oracle/libagf_rappids_port_*.so).
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PYR = 17

# TrajectoryTestResult bits (DepthImagePlanner.hpp:38-44)
LOW_COST, DYN_FEASIBLE, VEL_ADMISSIBLE, COLLISION_FREE = 1, 2, 4, 8


class Cfg(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("depth_scale", C.c_double),
                ("focal_length", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
                ("true_radius", C.c_double), ("planning_radius", C.c_double),
                ("min_checking_dist", C.c_double), ("min_thrust", C.c_double), ("max_thrust", C.c_double),
                ("max_angvel", C.c_double), ("min_section_time", C.c_double), ("max_velocity", C.c_double),
                ("max_pyramids", C.c_int32), ("cost_kind", C.c_int32), ("cost_vec", C.c_double * 3)]


class Sampler(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("min_x", "max_x", "min_y", "max_y", "min_depth", "max_depth",
                                          "min_time", "max_time")]


class Out(C.Structure):
    _fields_ = [("found", C.c_int32), ("best_index", C.c_int32), ("n_generated", C.c_int32),
                ("n_cost_checks", C.c_int32), ("n_collision_checks", C.c_int32),
                ("n_velocity_checks", C.c_int32), ("n_collision_free", C.c_int32), ("n_pyramids", C.c_int32),
                ("best_cost", C.c_double), ("best_coeffs", C.c_double * 18), ("best_tf", C.c_double)]

    def as_dict(self):
        d = {n: getattr(self, n) for n, _ in self._fields_ if n != "best_coeffs"}
        d["best_coeffs"] = np.array(self.best_coeffs).reshape(6, 3)
        return d


def default_cfg(width=320, height=240, cost_kind=0, cost_vec=(0.0, 0.0, 1.0), max_pyramids=0, **kw):
    """The configuration Rappids_Simulator uses (Simulator/Rappids_Simulator/main.cpp:121-122,167-169,360,
    484-489) with the DepthImagePlanner constructor defaults (DepthImagePlanner.cpp:44-52)."""
    c = Cfg()
    c.width, c.height = width, height
    c.depth_scale = 10.0 / 256.0
    c.focal_length = width / 2.0
    c.cx, c.cy = width / 2.0, height / 2.0
    c.true_radius, c.planning_radius, c.min_checking_dist = 0.116, 0.174, 0.5
    c.min_thrust, c.max_thrust, c.max_angvel, c.min_section_time, c.max_velocity = 5.0, 30.0, 20.0, 0.02, 5.0
    c.max_pyramids = max_pyramids
    c.cost_kind = cost_kind
    c.cost_vec[:] = cost_vec
    for k, v in kw.items():
        setattr(c, k, v)
    return c


PATHS = {
    "ref-glibc": os.path.join(HERE, "_ref", "libagf_rappids_ref_glibc.so"),
    "ref-shared": os.path.join(HERE, "_ref", "libagf_rappids_ref_shared.so"),
    "port-glibc": os.path.join(HERE, "libagf_rappids_port_glibc.so"),
    "port-shared": os.path.join(HERE, "libagf_rappids_port_shared.so"),
}


def available(flavour):
    return os.path.exists(PATHS[flavour])


_dp = C.POINTER(C.c_double)


def _d(a):
    return a.ctypes.data_as(_dp) if a is not None else None


class Planner:
    def __init__(self, flavour):
        self.flavour = flavour
        self.lib = L = C.CDLL(PATHS[flavour])
        L.orc_rappids_flavour.restype = C.c_char_p
        assert L.orc_rappids_flavour().decode() == flavour, (L.orc_rappids_flavour(), flavour)
        L.orc_rappids_plan.restype = C.c_int
        L.orc_rappids_plan.argtypes = [C.POINTER(Cfg), C.c_void_p, _dp, _dp, _dp, C.c_int32, _dp, C.c_int32,
                                       C.POINTER(Sampler), C.POINTER(Out), C.c_void_p, _dp, _dp, C.c_int32]
        L.orc_rappids_plan_many.restype = C.c_int
        L.orc_rappids_plan_many.argtypes = [C.POINTER(Cfg), C.c_int32, C.c_void_p, _dp, _dp, _dp, C.c_int32, _dp,
                                            C.POINTER(Out), C.c_void_p, C.c_int32]
        L.orc_rappids_solve_cubic.restype = C.c_int
        L.orc_rappids_solve_cubic.argtypes = [C.c_double] * 3 + [_dp]
        L.orc_rappids_solve_quartic.restype = C.c_int
        L.orc_rappids_solve_quartic.argtypes = [C.c_double] * 4 + [_dp]
        L.orc_rappids_primitive.restype = C.c_int
        L.orc_rappids_primitive.argtypes = [_dp] * 4 + [C.c_double] * 6 + [_dp, C.POINTER(C.c_int32),
                                                                            C.POINTER(C.c_int32)]

    def pixels_read(self):
        """port only: pixels InflatePyramid read (reference scan order) in all plans since the last call"""
        self.lib.orc_rappids_pixels_read.restype = C.c_uint64
        return int(self.lib.orc_rappids_pixels_read())

    def ground_truth(self, cfg, image, vel0, acc0, grav, candidates):
        """DepthImagePlanner::IsCollisionFreeGroundTruth for each candidate [n][4] -> bool[n]."""
        L = self.lib
        L.orc_rappids_ground_truth.restype = C.c_int
        L.orc_rappids_ground_truth.argtypes = [C.POINTER(Cfg), C.c_void_p, _dp, _dp, _dp, C.c_int32, _dp, C.c_void_p]
        image = np.ascontiguousarray(image, dtype=np.uint16)
        v, a, g = (np.ascontiguousarray(x, dtype=np.float64) for x in (vel0, acc0, grav))
        candidates = np.ascontiguousarray(candidates, dtype=np.float64).reshape(-1, 4)
        out = np.zeros(len(candidates), dtype=np.uint8)
        rc = L.orc_rappids_ground_truth(C.byref(cfg), image.ctypes.data, _d(v), _d(a), _d(g), len(candidates), _d(candidates),
                                        out.ctypes.data)
        assert rc == 0
        return out.astype(bool)

    def plan(self, cfg, image, vel0, acc0, grav, n=None, candidates=None, seed=0, sampler=None, max_pyr=256):
        """-> dict(out fields..., results[n] u8, candidates[n,4], pyramids[n_pyr,17])."""
        image = np.ascontiguousarray(image, dtype=np.uint16)
        assert image.shape == (cfg.height, cfg.width)
        v, a, g = (np.ascontiguousarray(x, dtype=np.float64) for x in (vel0, acc0, grav))
        if candidates is not None:
            candidates = np.ascontiguousarray(candidates, dtype=np.float64)
            n = candidates.shape[0]
        res = np.zeros(n, dtype=np.uint8)
        cout = np.zeros((n, 4))
        pyr = np.full((max_pyr, PYR), np.nan)
        out = Out()
        rc = self.lib.orc_rappids_plan(C.byref(cfg), image.ctypes.data, _d(v), _d(a), _d(g), n, _d(candidates),
                                       seed, C.byref(sampler) if sampler is not None else None, C.byref(out),
                                       res.ctypes.data, _d(cout), _d(pyr), max_pyr)
        assert rc == 0
        d = out.as_dict()
        d["results"] = res
        d["candidates"] = cout
        d["pyramids"] = pyr[:min(out.n_pyramids, max_pyr)]
        return d

    def plan_many(self, cfg, images, vel0, acc0, grav, candidates, threads=1, want_results=True):
        images = np.ascontiguousarray(images, dtype=np.uint16)
        n = images.shape[0]
        v, a, g = (np.ascontiguousarray(x, dtype=np.float64) for x in (vel0, acc0, grav))
        candidates = np.ascontiguousarray(candidates, dtype=np.float64)
        k = candidates.shape[1]
        outs = (Out * n)()
        res = np.zeros((n, k), dtype=np.uint8) if want_results else None
        rc = self.lib.orc_rappids_plan_many(C.byref(cfg), n, images.ctypes.data, _d(v), _d(a), _d(g), k,
                                            _d(candidates), outs, res.ctypes.data if want_results else None, threads)
        assert rc == 0
        return outs, res

    def solve_cubic(self, a, b, c):
        r = np.zeros(3)
        n = self.lib.orc_rappids_solve_cubic(a, b, c, _d(r))
        return n, r

    def solve_quartic(self, a, b, c, d):
        r = np.zeros(4)
        n = self.lib.orc_rappids_solve_quartic(a, b, c, d, _d(r))
        return n, r[:n]

    def primitive(self, vel0, acc0, grav, goal, T, fmin=5.0, fmax=30.0, wmax=20.0, min_section=0.02, vmax=5.0):
        v, a, g, q = (np.ascontiguousarray(x, dtype=np.float64) for x in (vel0, acc0, grav, goal))
        abg = np.zeros(9)
        ir, vr = C.c_int32(), C.c_int32()
        self.lib.orc_rappids_primitive(_d(v), _d(a), _d(g), _d(q), T, fmin, fmax, wmax, min_section, vmax, _d(abg),
                                       C.byref(ir), C.byref(vr))
        return abg.reshape(3, 3), ir.value, vr.value
