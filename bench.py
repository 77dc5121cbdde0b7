#!/usr/bin/env python
"""bench.py -- vehicle-steps/s of the batched simulation step on N B200s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]                      # own arm (CUDA path), headline = C3 FP32
    python bench.py --precision fp64 | --math parity | --hk off              # first-class arms of the same workload
    python bench.py --config c4 [--c4-mode rates|full]                       # C4: parameter sweep + per-tick logging (HBM)
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]     # reference CPU arm (rank 0 only)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (config.workload).  C3 (default) = BASELINE config 3's per-GPU shard: 131 072 vehicles per GPU (1 M over 8
GPUs), full onboard loop (IMU synthesis with noise -> 9-state EKF with 8-anchor UWB ranging -> position / attitude /
rate controllers -> mixer) WITH the reference's per-Run() housekeeping (battery / temperature filters, UpdateWarnings,
rate monitors: `--hk on`, what Quadcopter_T::Run does every tick), 4-waypoint square from an in-kernel command schedule.
One "step" = one launch of the step kernel = `ticks_per_step` simulation ticks (2 ms each; default 15 000 = three 10 s
horizons, so that the driver's 20 timed steps last about 2 s) for every vehicle of the shard, followed by the
Monte-Carlo statistics read-out (one small kernel, and ONE all-gather of 16 doubles per rank over the library's own NCCL
communicator when N > 1 -- agf_batch_reduce_stats_comm).  Vehicles shard by contiguous index range; there is no
communication inside the step (weak scaling).

value   = vehicle-steps/s over all GPUs, state resident in HBM, CUDA events on the launching stream, max over ranks.
e2e     = the same through the public C ABI with HOST buffers: the population's 6-DOF state lives in pinned host memory
          between steps; every step copies it in (agf_batch_set_state), runs the ticks, and copies the state (C4: and
          the last logged record) and the statistics back out.
roofline  C3 is ALU-bound (nothing is a contraction; state stays in registers).  FLOP per vehicle-step, three ways:
            frac               SURVEY.md 8(d)'s contract figure (2 900 full mode: the sparsity-aware hand count)
            frac_instrumented  hardware-counted on the literal restatement of the reference's algorithm (parity kernel, HK on)
            frac_executed      hardware-counted on the very kernel that was timed (same template arguments, HK as timed)
          each x vehicle-steps/s of the step kernel alone / pipe peak (148 SM x 128 (FP32) or 64 (FP64) lanes x 2 x
          sm_max_mhz of MEASURED_PEAKS.json).  Kernels with an FP64 plant report both pipes (`pipes`) and take the
          busier one as `frac`.  Counts: profiles/flops.json (ncu csv files under profiles/r*/).
          C4 is reported against the measured HBM copy bandwidth: 68 B logged per vehicle-step.
"""
import argparse
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LOG_BYTES_FP32 = 68.0
SURVEY_FLOP = dict(full=2900.0, rates=1180.0)  # the contract figure for `frac`


def load_flops():
    """profiles/flops.json: hardware-counted FLOP per vehicle-step per kernel, {"<math>_<prec>_<mode>[_hk]": {"fp32": a, "fp64": b}}"""
    with open(os.path.join(ROOT, "profiles", "flops.json")) as f:
        return json.load(f)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c3", choices=["c3", "c4"])
    ap.add_argument("--c4-mode", default="rates", choices=["rates", "full"])
    ap.add_argument("--vehicles-per-gpu", type=int, default=0, help="default: 131072 (c3), 2097152 (c4)")
    ap.add_argument("--ticks-per-step", type=int, default=0, help="default: 15000 (c3), 1500 (c4)")
    ap.add_argument("--precision", default="fp32", choices=["fp32", "fp64"])
    ap.add_argument("--math", default="fast", choices=["fast", "parity"])
    ap.add_argument("--hk", default="on", choices=["on", "off"], help="the reference's per-Run() housekeeping (always on with --math parity)")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary measurements")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU work for the cpu_baseline sample")
    return ap.parse_args()


def peaks():
    p = dict(hbm_gbs=6650.0, sm_max_mhz=1965.0, source="fallback")
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            j = json.load(f)
        p.update(hbm_gbs=float(j["hbm_gbs"]), sm_max_mhz=float(j.get("sm_max_mhz", 1965.0)), source="measured")
    except Exception:
        pass
    p["fp32_tflops"] = 148 * 128 * 2 * p["sm_max_mhz"] * 1e6 / 1e12
    p["fp64_tflops"] = 148 * 64 * 2 * p["sm_max_mhz"] * 1e6 / 1e12
    return p


class ClockSampler:
    """SM clock and clock-event reasons sampled through NVML every 10 ms DURING the timed region (a thread of this
    process; nvidia-smi -lms buffers its output and loses short regions)."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.stop_flag = False
        self.th = None
        self.h = None
        self.nv = None

    def start(self):
        import threading
        try:
            import pynvml as nv
            nv.nvmlInit()
            uuid = None
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            except Exception:
                pass
            self.h = None
            if uuid:
                for cand in ("GPU-" + uuid, ("GPU-" + uuid).encode()):
                    try:
                        self.h = nv.nvmlDeviceGetHandleByUUID(cand)
                        break
                    except Exception:
                        continue
            if self.h is None:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES")
                idx = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
                self.h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.nv = nv
        except Exception:
            self.nv = None
            return

        def loop():
            nv = self.nv
            while not self.stop_flag:
                try:
                    sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                    mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
                    pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    self.rows.append((sm, mx, pw, rs))
                except Exception:
                    pass
                time.sleep(0.01)

        self.th = threading.Thread(target=loop, daemon=True)
        self.th.start()

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        if self.nv is None or self.th is None:
            return out
        self.stop_flag = True
        self.th.join(timeout=2)
        nv, rows = self.nv, self.rows
        if not rows:
            return out
        out["sm_mhz"] = statistics.median([r[0] for r in rows])
        out["sm_min_mhz"] = min(r[0] for r in rows)
        out["sm_max_mhz"] = max(r[1] for r in rows)
        out["power_w_max"] = max(r[2] for r in rows)
        out["power_w_median"] = statistics.median([r[2] for r in rows])
        names = (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap),
                 ("hw_power_brake_slowdown", nv.nvmlClocksEventReasonHwPowerBrakeSlowdown))
        for nm, bit in names:
            if any(r[3] & bit for r in rows):
                out["reasons"].append(nm)
        out["samples"] = len(rows)
        return out


def traffic_per_launch(key):
    """dram__bytes_read.sum + dram__bytes_write.sum of one step-kernel launch of a configuration, from the committed ncu
    captures (profiles/r*/traffic.json: {key: {"bytes": B, "vehicles": n, "ticks": S}}); None when it was not captured."""
    for rnd in ("r2", "r1"):
        try:
            with open(os.path.join(ROOT, "profiles", rnd, "traffic.json")) as f:
                t = json.load(f)
            if key in t:
                return t[key]
        except Exception:
            pass
    return None


def workload(agf, n, first, precision, ticks_total, seed=7, device=0, stream=None, hk=True, uwb=True, noise=True, math="fast",
             sweep=False):
    """The C3 shard (uwb=True: full onboard loop, waypoint square) or its rates-mode counterpart; sweep=True gives every
    vehicle its own mass / inertia / motor constants (C4)."""
    import numpy as np
    s = agf.scenarios
    cfg = agf.vehicle_cfg(vehicle_id=1, motor_time_const=0.015)
    prec = agf.abi.PREC_FP32 if precision == "fp32" else agf.abi.PREC_FP64
    cfgs = agf.sweep_cfgs(cfg, n, seed=9 + first) if sweep else cfg
    b = agf.Batch(cfgs, n, precision=prec, math=agf.abi.MATH_FAST if math == "fast" else agf.abi.MATH_PARITY, device=device,
                  uwb_comm_period=0.004 if uwb else 0.0, sigma_gyro=0.1 if noise else 0.0, sigma_acc=0.2 if noise else 0.0,
                  seed=seed, first_global_index=first, stream=stream, telemetry_warnings=hk)
    if uwb:
        for i, p in s.ANCHORS_8:
            b.add_anchor(i, p)
    init = s.monte_carlo_initial_states(first + n, seed=1234, yaw_max=np.pi / 3)[first:]
    b.set_state13(init)
    if uwb:
        b.set_schedule(s.waypoint_square_schedule(agf.codec, nticks=ticks_total))
    else:
        raw = agf.codec.encode_rates(0, 9.81 * 1.02, (0.0, 0.0, 0.0))
        b.set_schedule([(k, raw, -1) for k in range(0, ticks_total, 10)])
    return b, init


def flop_key(math, precision, uwb, hk):
    return "%s_%s_%s%s" % (math, precision, "full" if uwb else "rates", "_hk" if (hk and math == "fast") else "")


def alu_roofline(pk, flops, math, precision, uwb, hk, kernel_sps, kms, kl):
    """The three FLOP-per-step figures x kernel rate against the pipe peaks (module docstring)."""
    mode = "full" if uwb else "rates"
    ex = flops.get(flop_key(math, precision, uwb, hk)) or flops.get(flop_key(math, precision, uwb, False))
    lit = flops["parity_fp64_" + mode]
    survey = SURVEY_FLOP[mode]
    p32, p64 = pk["fp32_tflops"], pk["fp64_tflops"]
    r = dict(kernel_vehicle_steps_per_s=kernel_sps)
    if precision == "fp32":
        # everything on the FP32 pipe
        r.update(bound="fp32_alu", peak=p32, unit="TFLOP/s", achieved=kernel_sps * survey / 1e12, frac=kernel_sps * survey / 1e12 / p32,
                 frac_instrumented=kernel_sps * (lit["fp32"] + lit["fp64"]) / 1e12 / p32,
                 frac_executed=None if ex is None else kernel_sps * (ex["fp32"] + ex["fp64"]) / 1e12 / p32)
    else:
        # FP64 plant + FP32 onboard logic: two pipes; the contract figure is split like the literal count (613 of
        # 1 180 rates-mode FLOP are FP64; full mode adds FP32 work only)
        s64 = 613.0
        s32 = survey - s64
        f64, f32 = kernel_sps * s64 / 1e12 / p64, kernel_sps * s32 / 1e12 / p32
        r.update(bound="fp64_alu" if f64 >= f32 else "fp32_alu", unit="TFLOP/s",
                 peak=p64 if f64 >= f32 else p32, achieved=kernel_sps * (s64 if f64 >= f32 else s32) / 1e12, frac=max(f64, f32),
                 pipes=dict(contract=dict(fp64_flop=s64, fp32_flop=s32, fp64_frac=f64, fp32_frac=f32),
                            instrumented=dict(fp64_flop=lit["fp64"], fp32_flop=lit["fp32"],
                                              fp64_frac=kernel_sps * lit["fp64"] / 1e12 / p64, fp32_frac=kernel_sps * lit["fp32"] / 1e12 / p32)))
        r["frac_instrumented"] = max(r["pipes"]["instrumented"]["fp64_frac"], r["pipes"]["instrumented"]["fp32_frac"])
        if ex is not None:
            r["pipes"]["executed"] = dict(fp64_flop=ex["fp64"], fp32_flop=ex["fp32"], fp64_frac=kernel_sps * ex["fp64"] / 1e12 / p64,
                                          fp32_frac=kernel_sps * ex["fp32"] / 1e12 / p32)
            r["frac_executed"] = max(r["pipes"]["executed"]["fp64_frac"], r["pipes"]["executed"]["fp32_frac"])
        else:
            r["frac_executed"] = None
    r["flop_per_vehicle_step"] = dict(contract_survey_8d=survey, instrumented_literal=lit, executed_by_timed_kernel=ex,
                                      source="profiles/flops.json")
    r["note"] = ("ALU-bound kernel (no contraction, state in registers; HBM is touched at launch boundaries only). frac = SURVEY 8(d) "
                 "contract FLOP x vehicle-steps/s of the step kernel alone (CUDA events inside the library on the launching stream, %d "
                 "launches, %.3f ms each) / pipe peak (148 SM x lanes x 2 x %.0f MHz, %s sm_max_mhz); frac_instrumented uses the "
                 "hardware-counted literal algorithm, frac_executed the hardware-counted FLOP of the timed kernel itself"
                 % (kl, kms / max(kl, 1), pk["sm_max_mhz"], pk["source"]))
    return r


def kernel_rate(b, n, S, launches=2):
    """vehicle-steps/s of the step kernel alone: one warm launch, then `launches` timed by the library's CUDA events"""
    b.run(S)
    b.sync()
    b.step_kernel_time()
    for _ in range(launches):
        b.run(S)
    ms, nl = b.step_kernel_time()
    return n * S * nl / (ms * 1e-3), ms, nl


def cpu_baseline(args, n_threads=None, seconds=12.0):
    """The reference's CPU implementation of the same workload on the host cores: oracle/_ref (the unmodified
    reference sources) when it was built, else the oracle port.  Bounded sample.  No product library is loaded here:
    the vehicle configuration comes from the reference's own QuadcopterConstants inside the harness."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import orc
    import numpy as np
    from importlib import import_module
    scen = import_module("agrifly_b200.scenarios")
    kind = "reference" if orc.available("ref-glibc") else "port"
    O = orc.Oracle("ref-glibc" if kind == "reference" else "port-glibc")
    cores = n_threads or os.cpu_count() or 1
    cfg = orc.reference_vehicle_cfg(O, vehicle_id=1, motor_time_const=0.015)
    anchors = np.array([[i, *p] for i, p in scen.ANCHORS_8], np.float32)
    nticks = 1000
    sched = scen.waypoint_square_schedule(orc.RefCodec(O), nticks=nticks)
    # calibrate: ~4e5 vehicle-steps/s/core for the full mode at -O3
    n = int(seconds * cores * 4e5 / nticks)
    n = max(cores * 4, min(n, cores * 8192))
    init = scen.monte_carlo_initial_states(n, seed=1234, yaw_max=np.pi / 3)
    _, secs = O.run_population(cfg, n, init13=init, anchors=anchors, nticks=nticks, sched=sched, threads=cores,
                               uwb_comm_period=0.004, sigma_acc=0.2 if kind == "reference" else 0.0,
                               sigma_gyro=0.1 if kind == "reference" else 0.0)
    return dict(value=n * nticks / secs, unit="vehicle-steps/s", cores=cores, kind=kind,
                sample="%d vehicles x %d ticks (the GPU arm's C3 workload: full onboard mode, EKF + 8-anchor UWB, IMU noise, waypoint "
                       "square; per-vehicle-step metric, so the smaller population and shorter horizon do not change it), %s, %d threads, "
                       "%.1f s. Caveat: the oracle build multiplies the 9x9 EKF matrices through a naive sequential-k Eigen shim "
                       "(oracle/shim/Eigen), not Eigen 3.3.7's vectorised kernels, which the reference would use when built with its own "
                       "CMake -- a real build may be somewhat faster on the EKF products" %
                       (n, nticks, "oracle/_ref (unmodified reference sources, g++ -O3, no FMA)" if kind == "reference"
                        else "oracle port (-O3, no FMA)", cores, secs)), secs


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.time()
    vals = []
    cb = None
    for i in range(args.warmup + args.steps):
        cb, secs = cpu_baseline(args, seconds=max(2.0, min(args.cpu_seconds, 60.0 / max(1, args.warmup + args.steps))))
        if i >= args.warmup:
            vals.append((cb["value"], secs))
        if time.time() - t0 > 240:
            break
    v = sum(x for x, _ in vals) / max(1, len(vals))
    ms = 1e3 * sum(s for _, s in vals) / max(1, len(vals))
    cb["value"] = v
    line = dict(metric="vehicle-steps/s", value=v, unit="vehicle-steps/s", n_gpus=args.gpus, steps=len(vals),
                warmup=args.warmup, ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f64 plant + f32 onboard logic (reference mixed precision)", data="synthetic", impl="reference",
                config=dict(workload="C3 shard on the host CPU (a bounded sample of the GPU arm's workload; the metric is per "
                                     "vehicle-step, so the ratio to the GPU arm stands although population and horizon are smaller): "
                                     + cb["sample"]),
                cpu_baseline=cb,
                e2e=dict(value=v, unit="vehicle-steps/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line))


def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    import numpy as np
    import torch
    import agrifly_b200 as agf
    from agrifly_b200 import sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the simulation step has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pk = peaks()
    flops = load_flops()
    c4 = args.config == "c4"
    K, W = args.steps, args.warmup
    S = args.ticks_per_step or (1500 if c4 else 15000)
    n = args.vehicles_per_gpu or ((1 << 21) if c4 else 131072)
    hk = (args.hk == "on") or args.math == "parity"
    uwb = (not c4) or args.c4_mode == "full"
    n_total = n * world
    first = rank * n
    stream = torch.cuda.Stream()
    ticks_total = (W + K) * S * 2 + 16
    b, init = workload(agf, n, first, args.precision, ticks_total, device=local, stream=stream.cuda_stream, hk=hk, uwb=uwb,
                       math=args.math, sweep=c4)
    LOG_RING = 32
    if c4:
        b.enable_log(1, LOG_RING)
    stats = torch.zeros(16, dtype=torch.float64, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    comm = sharding.StatsComm(b, dist, local) if dist is not None else None  # the library's own NCCL communicator

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step():
        b.run(S)
        if comm is not None:
            comm.reduce_device(stats.data_ptr())  # stats kernel + ONE all-gather + combine, on the batch's stream
        else:
            b.stats_device(stats.data_ptr())

    with torch.cuda.stream(stream):
        for _ in range(W):
            one_step()
            flush.zero_()
        barrier()
        b.step_kernel_time()
        l0 = b.launch_count
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t_wall0 = time.time()
        e0.record(stream)
        for _ in range(K):
            one_step()
            flush.zero_()  # > L2 (126 MB) written between timed iterations
        e1.record(stream)
        barrier()
        t_wall = time.time() - t_wall0
        clocks = sampler.stop() if rank == 0 else None
        ms = e0.elapsed_time(e1)
        kms, kl = b.step_kernel_time()
        launches = b.launch_count - l0
        final_stats = stats.cpu().numpy().copy()

        # ---- e2e: public API with host buffers -------------------------------------------------
        # the population's 6-DOF state lives in pinned host memory between steps: read back at the end of a step, handed in
        # again at the start of the next (a host-owned state, as with the reference's objects; teleporting the vehicles to
        # some fixed state every step instead would throw the onboard estimators off and change the workload)
        pin13 = torch.empty((n, 13), dtype=torch.float64).pin_memory()
        agf._check(b.L.agf_batch_get_state(b.h, pin13.data_ptr(), 0, n))
        if c4:
            rec_out = torch.empty((n, 17), dtype=torch.float64).pin_memory()
        ke = max(2, min(K, 10))

        def e2e_step():
            agf._check(b.L.agf_batch_set_state(b.h, pin13.data_ptr(), 0, n))  # H2D from pinned host memory, one copy
            b.run(S)
            if c4:  # D2H: the newest logged record of every vehicle
                agf._check(b.L.agf_batch_read_log(b.h, b.log_count - 1, rec_out.data_ptr(), 0, n))
            agf._check(b.L.agf_batch_get_state(b.h, pin13.data_ptr(), 0, n))  # D2H: the state, into the same pinned buffer
            return comm.reduce_host() if comm is not None else b.stats()  # D2H statistics vector

        e2e_step()
        barrier()
        t0 = time.time()
        for _ in range(ke):
            e2e_step()
        barrier()
        t_e2e = time.time() - t0
    t = torch.tensor([ms, t_e2e * 1e3, kms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms, kms = [float(x) for x in t.cpu()]
    value = n_total * S * K / (ms * 1e-3)
    e2e_value = n_total * S * ke / (e2e_ms * 1e-3)
    kernel_sps = n * S * kl / (kms * 1e-3) if kms > 0 else 0.0  # per GPU, step kernel alone
    mode_txt = ("full onboard loop (IMU noise, 9-state EKF + 8-anchor UWB, position/attitude/rate control, mixer), 4-waypoint square"
                if uwb else "rates mode (IMU noise, complementary filter, rate control, mixer), hover thrust command")
    if c4:
        wl = ("C4 shard: parameter sweep -- every vehicle its own mass, inertia, kF, ktau, motor time constant -- with the 17-float state "
              "record of every vehicle logged to HBM EVERY tick (ring of %d records), %s" % (LOG_RING, mode_txt))
        gbs = kernel_sps * LOG_BYTES_FP32 * (1.0 if args.precision == "fp32" else 2.0) / 1e9
        tr = traffic_per_launch("c4_%s_%s" % (args.precision, args.c4_mode))
        roof = dict(bound="hbm", achieved=gbs, peak=pk["hbm_gbs"], unit="GB/s", frac=gbs / pk["hbm_gbs"],
                    traffic=None if tr is None else tr["bytes"] * (n * S) / float(tr["vehicles"] * tr["ticks"]),
                    traffic_over_algorithmic=None if tr is None else tr["bytes"] / (tr["vehicles"] * tr["ticks"] * LOG_BYTES_FP32),
                    kernel_vehicle_steps_per_s=kernel_sps,
                    alu=alu_roofline(pk, flops, args.math, args.precision, uwb, hk, kernel_sps, kms, kl),
                    note="algorithmic bytes = 68 B logged per vehicle-step (17 floats) x vehicle-steps of one launch / its CUDA-event "
                         "duration; peak = measured copy bandwidth (%s); traffic = ncu dram__bytes of a captured launch scaled to this "
                         "launch's vehicle-steps (profiles/r2/traffic.json); `alu` = the same kernel against the FP32 pipe" % pk["source"])
    else:
        wl = "C3 shard: waypoint tracking, " + mode_txt
        roof = alu_roofline(pk, flops, args.math, args.precision, uwb, hk, kernel_sps, kms, kl)
        tr = traffic_per_launch("c3_%s_%s%s" % (args.precision, args.math, "_hk" if hk else "")) or \
            traffic_per_launch("c3_%s_%s" % (args.precision, args.math))
        # the state round trip of one launch: independent of the tick count, proportional to the vehicles
        roof["traffic"] = None if tr is None else tr["bytes"] * n / float(tr["vehicles"])
    line = dict(
        metric="vehicle-steps/s", value=value, unit="vehicle-steps/s", n_gpus=world, steps=K, warmup=W,
        ms_per_step=ms / K, higher_is_better=True, scaling="weak", vs_baseline=None,
        dtype="f32" if args.precision == "fp32" else "f64 plant + f32 onboard logic (the reference's mixed precision)", data="synthetic",
        config=dict(workload=wl, vehicles_per_gpu=n, vehicles_total=n_total, ticks_per_step=S, dt_us=2000, precision=args.precision,
                    math="fast (FMA, CUDA libm)" if args.math == "fast" else "parity (no FMA, shared libm: bit-identical to the oracle)",
                    housekeeping="on (battery/temperature filters, UpdateWarnings, rate monitors, propeller calibration: what the "
                                 "reference's Run() does)" if hk else "off (telemetry_warnings=0)",
                    parallelism="vehicles sharded by index range, no collective in the step; one all-gather of a 16-double statistics "
                                "vector per step on the library's own NCCL communicator" if world > 1 else "single GPU",
                    l2_flush="256 MiB memset between timed steps (inside the timed region); state stays in registers "
                             "during a step, HBM is touched only at launch boundaries" + (" and by the trajectory log" if c4 else "")),
        e2e=dict(value=e2e_value, unit="vehicle-steps/s", h2d_bytes_per_step=int(n * 13 * 8),
                 d2h_bytes_per_step=int(n * (13 + (17 if c4 else 0)) * 8 + 128), steps=ke,
                 note="per GPU bytes; every step: the population's 6-DOF state in from pinned host memory (agf_batch_set_state), the "
                      "ticks, %sthe state (agf_batch_get_state) and the statistics vector back out" % ("the newest log record, " if c4 else "")),
        gpu_launches=int(launches), roofline=roof, wall_ms=t_wall * 1e3, stats=sharding.summarize_stats(final_stats),
    )
    if rank == 0 and clocks is not None:
        line["clocks"] = dict(sm_mhz=clocks["sm_mhz"], sm_max_mhz=clocks["sm_max_mhz"], reasons=clocks["reasons"],
                              samples=clocks["samples"], power_w_max=clocks.get("power_w_max"), power_w_median=clocks.get("power_w_median"),
                              sm_min_mhz=clocks.get("sm_min_mhz"))
        if clocks["sm_mhz"] and "frac" in roof and not c4:
            roof["frac_at_observed_clock"] = roof["frac"] * pk["sm_max_mhz"] / clocks["sm_mhz"]
    if comm is not None:
        comm.close()
    b.close()

    # ---- secondary measurements (rank 0, N == 1) ---------------------------------------------------
    if rank == 0 and world == 1 and not args.no_extras:
        line["extra"] = extras(args, agf, pk, flops, n if not c4 else 131072, local, stream)
        cb, _ = cpu_baseline(args, seconds=args.cpu_seconds)
        line["cpu_baseline"] = cb
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def extras(args, agf, pk, flops, n, local, stream):
    """Kernel-only rates of the other variants, each with its own roofline (short runs: 2 launches of 500 ticks)."""
    import torch
    out = {}
    S = 500
    with torch.cuda.stream(stream):
        for name, prec, math, uwb, hk in (("fp32_full_hk_off", "fp32", "fast", True, False), ("fp32_full_hk_on", "fp32", "fast", True, True),
                                          ("fp64_full_hk_on", "fp64", "fast", True, True), ("fp64_full_parity", "fp64", "parity", True, True),
                                          ("fp32_rates_hk_on", "fp32", "fast", False, True), ("fp32_rates_hk_off", "fp32", "fast", False, False),
                                          ("fp64_rates_hk_on", "fp64", "fast", False, True), ("fp64_rates_parity", "fp64", "parity", False, True)):
            try:
                nb = n if prec == "fp32" else n // 2
                bb, _ = workload(agf, nb, 0, prec, 4 * S + 16, device=local, stream=stream.cuda_stream, uwb=uwb, hk=hk, math=math)
                sps, kms, kl = kernel_rate(bb, nb, S)
                r = alu_roofline(pk, flops, math, prec, uwb, hk, sps, kms, kl)
                r.pop("note", None)
                r.pop("flop_per_vehicle_step", None)
                out[name] = dict(vehicle_steps_per_s=sps, vehicles=nb, ticks_per_launch=S, roofline=r)
                bb.close()
            except Exception as ex:  # a secondary number must not take the headline down
                out[name] = dict(error=str(ex))
        if args.config != "c4":
            # C4 in short: parameter sweep + per-tick logging (the full line: bench.py --config c4)
            for mode in ("rates", "full"):
                try:
                    nl = 1 << 21
                    bl, _ = workload(agf, nl, 0, "fp32", 3 * 128 + 16, device=local, stream=stream.cuda_stream, uwb=(mode == "full"), hk=True,
                                     sweep=True)
                    bl.enable_log(1, 32)
                    sps, kms, kl = kernel_rate(bl, nl, 128)
                    gbs = sps * LOG_BYTES_FP32 / 1e9
                    out["c4_" + mode] = dict(vehicle_steps_per_s=sps, vehicles=nl, ticks_per_launch=128,
                                             roofline=dict(bound="hbm", achieved=gbs, peak=pk["hbm_gbs"], unit="GB/s", frac=gbs / pk["hbm_gbs"]))
                    bl.close()
                except Exception as ex:
                    out["c4_" + mode] = dict(error=str(ex))
        # the in-kernel offboard loop: Rappids_Simulator's closed loop, FP32, with the true state and with the
        # mocap estimator feeding the controller
        for key, with_est in (("fp32_offboard_loop_truth", False), ("fp32_offboard_loop_mocap", True)):
            try:
                bo = agf.Batch(agf.vehicle_cfg(vehicle_id=1, motor_time_const=0.015), n, precision=agf.abi.PREC_FP32,
                               math=agf.abi.MATH_FAST, device=local, stream=stream.cuda_stream, telemetry_warnings=False)
                bo.set_offboard_loop(agf.offboard_cfg(5), [(0, (0.0, 0.0, 2.0)), (3000000, (1.0, -0.5, 2.5))])
                if with_est:
                    bo.set_offboard_estimator(agf.offboard_estimator())
                sps, _, _ = kernel_rate(bo, n, S)
                out[key] = dict(vehicle_steps_per_s=sps, vehicles=n,
                                note="rates mode + offboard position controller at 100 Hz in the kernel" +
                                     (" + MocapStateEstimator at 200 Hz" if with_est else ", true state"))
                bo.close()
            except Exception as ex:
                out[key] = dict(error=str(ex))
    # C5: batched RAPPIDS planner (K6)
    try:
        out["rappids_c5"] = rappids_extra(agf, pk)
    except Exception as ex:
        out["rappids_c5"] = dict(error=str(ex))
    return out


def rappids_extra(agf, pk):
    nr, kr = 65536, 512  # BASELINE config 5: 64K vehicles
    pop = agf.scenarios.rappids_population(nr, seed=2024)
    with agf.Rappids(agf.rappids_cfg(math=agf.abi.MATH_FAST), nr, kr) as pl:
        pl.render_scenes(pop["row_bg"], pop["boxes"])
        pl.set_states(pop["vel0"], pop["acc0"], pop["grav"])
        pl.sample_candidates(kr, seed=7)
        pl.plan()
        pl.sync()
        pl.plan_kernel_time()
        for _ in range(3):
            pl.plan()
        pl.sync()
        pms, pcnt = pl.plan_kernel_time()
        # the same launch with vehicles handed out in index order instead of by their previous plan's work (the first plan of a
        # handle, or scenes that share nothing with the last frame)
        pl.set_dispatch(False)
        for _ in range(2):
            pl.plan()
        pl.sync()
        pms_index, _ = pl.plan_kernel_time()
        pl.set_dispatch(True)
        st = pl.stats()
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import orc_rappids
        # algorithmic bytes per plan: the depth pixels (2 B) InflatePyramid reads in the reference's scan order, COUNTED by the
        # CPU restatement on a sample of this very workload (oracle/port: orc_rappids_pixels_read)
        nsamp = 256
        port = orc_rappids.Planner("port-glibc")
        port.pixels_read()
        port.plan_many(orc_rappids.default_cfg(), pl.get_images(0, nsamp), pop["vel0"][:nsamp], pop["acc0"][:nsamp], pop["grav"][:nsamp],
                       pl.get_candidates(0, nsamp), threads=os.cpu_count() or 1, want_results=False)
        bytes_per_plan = 2.0 * port.pixels_read() / nsamp
        gbs = nr * bytes_per_plan / (pms * 1e-3) / 1e9
        r = dict(plans_per_s=nr / (pms * 1e-3), candidates_per_s=nr * kr / (pms * 1e-3), vehicles=nr, candidates=kr, ms_per_launch=pms,
                 found_fraction=st["found"] / nr, plans_per_s_index_order=nr / (pms_index * 1e-3), ms_per_launch_index_order=pms_index,
                 dispatch="candidate pass (one thread per candidate) + planning pass (one warp per vehicle), vehicles handed out by "
                          "descending device time of their previous plan; *_index_order: handed out in index order",
                 roofline=dict(bound="hbm", achieved=gbs, peak=pk["hbm_gbs"], unit="GB/s", frac=gbs / pk["hbm_gbs"],
                               algorithmic_bytes_per_plan=bytes_per_plan,
                               note="pixel scans of InflatePyramid: %.3f MB of depth pixels per plan in the reference's scan order, counted "
                                    "by the CPU restatement on %d plans of this workload; the kernel is latency-bound, not bandwidth-bound"
                                    % (bytes_per_plan / 1e6, nsamp)))
        # the reference planner on the host cores, same images / states / candidates (bounded sample)
        fl = "ref-glibc" if orc_rappids.available("ref-glibc") else "port-glibc"
        ns = min(nr, 128 * (os.cpu_count() or 1))
        imgs = pl.get_images(0, ns)
        cands = pl.get_candidates(0, ns)
        P = orc_rappids.Planner(fl)
        reps, t0 = 0, time.time()
        while reps < 40 and time.time() - t0 < 2.0:
            P.plan_many(orc_rappids.default_cfg(), imgs, pop["vel0"][:ns], pop["acc0"][:ns], pop["grav"][:ns], cands,
                        threads=os.cpu_count() or 1, want_results=False)
            reps += 1
        dt = time.time() - t0
        r["cpu_baseline"] = dict(value=ns * reps / dt, unit="plans/s", cores=os.cpu_count() or 1,
                                 kind="reference" if fl.startswith("ref") else "port",
                                 sample="%d x %d plans x %d candidates, %.2f s" % (reps, ns, kr, dt))
    return r


if __name__ == "__main__":
    main()
