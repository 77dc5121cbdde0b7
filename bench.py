#!/usr/bin/env python
"""bench.py -- vehicle-steps/s of the batched simulation step on N B200s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]                # own arm (CUDA path)
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]   # reference CPU arm (rank 0 only)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (config.workload): BASELINE config 3's per-GPU shard -- 131072 vehicles per GPU (1M over 8 GPUs),
FP32 plant, full onboard loop (IMU synthesis with noise -> 9-state EKF with 8-anchor UWB ranging -> position /
attitude / rate controllers -> mixer), 4-waypoint square from an in-kernel command schedule.  One "step" =
one launch of the step kernel = `ticks_per_step` simulation ticks (2 ms each) for every vehicle of the shard,
followed by the Monte-Carlo statistics reduction (one small kernel + all-reduce of 16 doubles over NCCL when
N > 1).  Vehicles shard by contiguous index range; there is no communication inside the step (weak scaling).

value   = vehicle-steps/s over all GPUs, state resident in HBM, timed with CUDA events on the launching stream,
          max over ranks.
e2e     = same metric through the public C ABI with HOST buffers: every step copies the population's 6-DOF
          state in from pinned host memory (agf_batch_set_state), runs the ticks, and copies positions and the
          statistics vector back (agf_batch_get_field / agf_batch_reduce_stats).
roofline: the step is ALU-bound (nothing is a contraction; state stays in registers): achieved = vehicle-steps/s of
          the step kernel alone x algorithmic FLOP per vehicle-step (2489 full mode, counted with hardware counters on
          the literal restatement of the reference's algorithm; SURVEY.md 8d's hand count 2900 and the 1378 FLOP the
          fast kernel really executes are reported beside it) against the FP32 pipe peak 148 SM x 128 lanes x 2 x
          sm_max_mhz (MEASURED_PEAKS.json).  The HBM-bound logging configuration (C4) is reported under
          "roofline_logging" against the measured copy bandwidth.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# Algorithmic FLOP per vehicle-step (+,-,x = 1, FMA = 2), three ways (DESIGN.md "Work per vehicle-step"):
#   *_INSTR  counted with hardware counters (ncu smsp__sass_thread_inst_executed_op_{f,d}{add,mul,fma}) on the parity
#            kernel, i.e. the reference's algorithm restated literally minus its structural zeros, in flight, noise on
#            (profiles/r1/flops_parity_*.csv).  This is the figure the roofline uses.
#   *_SURVEY the hand count of SURVEY.md 8d (2900 "sparsity-aware" full mode / 1180 rates mode)
#   *_EXEC   what the fast FP32 kernel actually executes per step (symmetric packed EKF, closed forms)
FLOP_FULL_INSTR, FLOP_FULL_SURVEY, FLOP_FULL_EXEC = 2489.0, 2900.0, 1378.0
FLOP_RATES_INSTR, FLOP_RATES_SURVEY, FLOP_RATES_EXEC = 1264.0, 1180.0, 699.0
FLOP_FULL, FLOP_RATES = FLOP_FULL_INSTR, FLOP_RATES_INSTR
LOG_BYTES_FP32 = 68.0


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--vehicles-per-gpu", type=int, default=131072)
    ap.add_argument("--ticks-per-step", type=int, default=500)
    ap.add_argument("--precision", default="fp32", choices=["fp32", "fp64"])
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary measurements (fp64, logging, cpu)")
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
        names = (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap),
                 ("hw_power_brake_slowdown", nv.nvmlClocksEventReasonHwPowerBrakeSlowdown))
        for nm, bit in names:
            if any(r[3] & bit for r in rows):
                out["reasons"].append(nm)
        out["samples"] = len(rows)
        return out


def traffic_per_launch(args, n, ticks):
    """dram__bytes_read.sum + dram__bytes_write.sum of one step-kernel launch of the bench configuration, from the
    committed ncu capture (profiles/r1/traffic.json); None when the configuration was not captured."""
    try:
        with open(os.path.join(ROOT, "profiles", "r1", "traffic.json")) as f:
            t = json.load(f)
        key = "%s_uwb_%d_%d" % (args.precision, n, ticks)
        return t.get(key)
    except Exception:
        return None


def workload(agf, n, first, precision, ticks_total, seed=7, device=0, stream=None, hk=False, uwb=True, noise=True):
    import numpy as np
    s = agf.scenarios
    cfg = agf.vehicle_cfg(vehicle_id=1, motor_time_const=0.015)
    prec = agf.abi.PREC_FP32 if precision == "fp32" else agf.abi.PREC_FP64
    b = agf.Batch(cfg, n, precision=prec, math=agf.abi.MATH_FAST, device=device, uwb_comm_period=0.004 if uwb else 0.0,
                  sigma_gyro=0.1 if noise else 0.0, sigma_acc=0.2 if noise else 0.0, seed=seed, first_global_index=first,
                  stream=stream, telemetry_warnings=hk)
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


def cpu_baseline(args, n_threads=None, seconds=12.0):
    """The reference's CPU implementation of the same workload on the host cores: oracle/_ref (the unmodified
    reference sources) when it was built, else the oracle port.  Bounded sample."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import agrifly_b200 as agf
    import orc
    import numpy as np
    kind = "reference" if orc.available("ref-glibc") else "port"
    O = orc.Oracle("ref-glibc" if kind == "reference" else "port-glibc")
    cores = n_threads or os.cpu_count() or 1
    s = agf.scenarios
    cfg = agf.vehicle_cfg(vehicle_id=1, motor_time_const=0.015)
    anchors = np.array([[i, *p] for i, p in s.ANCHORS_8], np.float32)
    nticks = 1000
    sched = s.waypoint_square_schedule(agf.codec, nticks=nticks)
    # calibrate: ~4e5 vehicle-steps/s/core for the full mode at -O3
    n = int(seconds * cores * 4e5 / nticks)
    n = max(cores * 4, min(n, cores * 8192))
    init = s.monte_carlo_initial_states(n, seed=1234, yaw_max=np.pi / 3)
    _, secs = O.run_population(cfg, n, init13=init, anchors=anchors, nticks=nticks, sched=sched, threads=cores,
                               uwb_comm_period=0.004, sigma_acc=0.2 if kind == "reference" else 0.0,
                               sigma_gyro=0.1 if kind == "reference" else 0.0)
    return dict(value=n * nticks / secs, unit="vehicle-steps/s", cores=cores, kind=kind,
                sample="%d vehicles x %d ticks, full onboard mode (EKF+UWB), %s, %d threads, %.1f s" %
                       (n, nticks, "oracle/_ref (unmodified reference sources, -O3, no FMA)" if kind == "reference"
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
                config=dict(workload="C3 shard on the host CPU: " + cb["sample"]), cpu_baseline=cb,
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
    S, K, W = args.ticks_per_step, args.steps, args.warmup
    n = args.vehicles_per_gpu
    n_total = n * world
    first = rank * n
    stream = torch.cuda.Stream()
    ticks_total = (W + K) * S * 2 + 16
    b, init = workload(agf, n, first, args.precision, ticks_total, device=local, stream=stream.cuda_stream)
    stats = torch.zeros(16, dtype=torch.float64, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step():
        b.run(S)
        b.stats_device(stats.data_ptr())
        if dist is not None:
            sharding.combine_stats(stats, dist)

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
        pin13 = torch.from_numpy(np.ascontiguousarray(init[:, 0:13])).pin_memory()  # the population's 6-DOF state, pinned host memory
        pos_out = torch.empty((n, 3), dtype=torch.float64).pin_memory()
        ke = max(2, min(K, 20))

        def e2e_step():
            agf._check(b.L.agf_batch_set_state(b.h, pin13.data_ptr(), 0, n))  # H2D from pinned host memory, one copy
            b.run(S)
            agf._check(b.L.agf_batch_get_field(b.h, 0, pos_out.data_ptr(), 0, n))  # D2H positions
            return b.stats()  # D2H statistics vector

        e2e_step()
        barrier()
        t0 = time.time()
        for _ in range(ke):
            st_e2e = e2e_step()
        barrier()
        t_e2e = time.time() - t0
    t = torch.tensor([ms, t_e2e * 1e3, kms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms, kms = [float(x) for x in t.cpu()]
    value = n_total * S * K / (ms * 1e-3)
    e2e_value = n_total * S * ke / (e2e_ms * 1e-3)
    flop = FLOP_FULL
    kernel_steps_per_s = n * S * kl / (kms * 1e-3) if kms > 0 else 0.0  # per GPU, step kernel alone
    peak = pk["fp32_tflops"] if args.precision == "fp32" else pk["fp64_tflops"]
    achieved = kernel_steps_per_s * flop / 1e12
    line = dict(
        metric="vehicle-steps/s", value=value, unit="vehicle-steps/s", n_gpus=world, steps=K, warmup=W,
        ms_per_step=ms / K, higher_is_better=True, scaling="weak", vs_baseline=None,
        dtype="f32" if args.precision == "fp32" else "f64 plant + f32 onboard logic", data="synthetic",
        config=dict(workload="C3 shard: waypoint tracking, full onboard loop (IMU noise, 9-state EKF + 8-anchor UWB, "
                             "position/attitude/rate control, mixer), 4-waypoint square",
                    vehicles_per_gpu=n, vehicles_total=n_total, ticks_per_step=S, dt_us=2000, precision=args.precision,
                    math="fast (FMA, CUDA libm)", parallelism="vehicles sharded by index range, no collective in the step; "
                    "all-reduce of a 16-double statistics vector per step",
                    l2_flush="256 MiB memset between timed steps (inside the timed region); state stays in registers "
                             "during a step, HBM is touched only at launch boundaries"),
        e2e=dict(value=e2e_value, unit="vehicle-steps/s", h2d_bytes_per_step=int(n * 13 * 8), d2h_bytes_per_step=int(n * 3 * 8 + 128),
                 steps=ke, note="per GPU bytes; state set from pinned host memory, positions + statistics read back"),
        gpu_launches=int(launches),
        roofline=dict(bound="fp32_alu" if args.precision == "fp32" else "fp64_alu", achieved=achieved, peak=peak, unit="TFLOP/s",
                      frac=achieved / peak, traffic=traffic_per_launch(args, n, S),
                      flop_per_vehicle_step=dict(used=flop, instrumented_literal=FLOP_FULL_INSTR, survey_hand_count=FLOP_FULL_SURVEY,
                                                 executed_by_fast_kernel=FLOP_FULL_EXEC),
                      frac_survey_count=kernel_steps_per_s * FLOP_FULL_SURVEY / 1e12 / peak,
                      frac_executed=kernel_steps_per_s * FLOP_FULL_EXEC / 1e12 / peak,
                      kernel_vehicle_steps_per_s=kernel_steps_per_s,
                      note="ALU-bound kernel (no contraction, state in registers; HBM is touched at launch boundaries only: "
                           "traffic = DRAM bytes of one launch from profiles/, a few %% of what the HBM could move in that time). "
                           "achieved = %.0f FLOP per vehicle-step (hardware-counted on the literal restatement of the reference's "
                           "algorithm) x vehicle-steps/s of the step kernel alone (CUDA events inside the library on the launching "
                           "stream, %d launches, %.3f ms each); peak = FP32 pipe 148 SM x 128 lanes x 2 x %.0f MHz (%s sm_max_mhz); "
                           "frac_executed counts only the FLOP the fast kernel really issues"
                           % (flop, kl, kms / max(kl, 1), pk["sm_max_mhz"], pk["source"])),
        wall_ms=t_wall * 1e3, stats=sharding.summarize_stats(final_stats),
    )
    if rank == 0 and clocks is not None:
        line["clocks"] = dict(sm_mhz=clocks["sm_mhz"], sm_max_mhz=clocks["sm_max_mhz"], reasons=clocks["reasons"],
                              samples=clocks["samples"], power_w_max=clocks.get("power_w_max"))
        if clocks["sm_mhz"]:
            line["roofline"]["frac_at_observed_clock"] = achieved / (peak * clocks["sm_mhz"] / pk["sm_max_mhz"])
    b.close()

    # ---- secondary measurements (rank 0, N == 1): FP64 mode, HBM-bound logging, CPU baseline --------
    if rank == 0 and world == 1 and not args.no_extras:
        extras = {}
        with torch.cuda.stream(stream):
            for name, prec, uwb in (("fp64_full", "fp64", True), ("fp32_rates", "fp32", False), ("fp64_rates", "fp64", False)):
                nb = n if prec == "fp32" else n // 2
                bb, _ = workload(agf, nb, 0, prec, 4 * S + 16, device=local, stream=stream.cuda_stream, uwb=uwb)
                bb.run(S)
                bb.sync()
                bb.step_kernel_time()
                bb.run(S)
                bb.run(S)
                k2, l2 = bb.step_kernel_time()
                sps = nb * S * l2 / (k2 * 1e-3)
                fl = FLOP_FULL if uwb else FLOP_RATES
                pkk = pk["fp32_tflops"] if prec == "fp32" else pk["fp64_tflops"]
                extras[name] = dict(vehicle_steps_per_s=sps, vehicles=nb, flop_per_step=fl, achieved_tflops=sps * fl / 1e12,
                                    frac_of_alu_peak=sps * fl / 1e12 / pkk)
                bb.close()
            # C4-style logging: every tick, 17 floats per vehicle -> HBM
            nl = 1 << 21
            bl, _ = workload(agf, nl, 0, "fp32", 3 * 64 + 16, device=local, stream=stream.cuda_stream, uwb=False)
            bl.enable_log(1, 32)
            bl.run(64)
            bl.sync()
            bl.step_kernel_time()
            bl.run(64)
            bl.run(64)
            k3, l3 = bl.step_kernel_time()
            sps = nl * 64 * l3 / (k3 * 1e-3)
            gbs = sps * LOG_BYTES_FP32 / 1e9
            line["roofline_logging"] = dict(bound="hbm", achieved=gbs, peak=pk["hbm_gbs"], unit="GB/s", frac=gbs / pk["hbm_gbs"],
                                            traffic=None, vehicle_steps_per_s=sps,
                                            note="rates mode, %d vehicles, 17 floats logged per vehicle-step (68 B), ring of 32 records" % nl)
            bl.close()
            # the in-kernel offboard loop: Rappids_Simulator's closed loop, FP32, with the true state and with the
            # mocap estimator feeding the controller
            for key, with_est in (("fp32_offboard_loop_truth", False), ("fp32_offboard_loop_mocap", True)):
                try:
                    bo = agf.Batch(agf.vehicle_cfg(vehicle_id=1, motor_time_const=0.015), n, precision=agf.abi.PREC_FP32,
                                   math=agf.abi.MATH_FAST, device=local, stream=stream.cuda_stream, telemetry_warnings=False)
                    bo.set_offboard_loop(agf.offboard_cfg(5), [(0, (0.0, 0.0, 2.0)), (3000000, (1.0, -0.5, 2.5))])
                    if with_est:
                        bo.set_offboard_estimator(agf.offboard_estimator())
                    bo.run(S)
                    bo.sync()
                    bo.step_kernel_time()
                    bo.run(S)
                    bo.run(S)
                    k4, l4 = bo.step_kernel_time()
                    extras[key] = dict(vehicle_steps_per_s=n * S * l4 / (k4 * 1e-3), vehicles=n,
                                       note="rates mode + offboard position controller at 100 Hz in the kernel" +
                                            (" + MocapStateEstimator at 200 Hz" if with_est else ", true state"))
                    bo.close()
                except Exception as ex:  # a secondary number must not take the headline down
                    extras[key] = dict(error=str(ex))
        # C5: batched RAPPIDS planner (K6), HBM-latency-bound pixel scans
        try:
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
                st = pl.stats()
                bytes_per_plan = 1.06e6  # algorithmic: bytes of the pixels the reference scans per plan of this workload (ncu DRAM reads of the all-pixels build, profiles/r1/rappids_plan_fast_summary_v0.txt)
                gbs = nr * bytes_per_plan / (pms * 1e-3) / 1e9
                extras["rappids_c5"] = dict(plans_per_s=nr / (pms * 1e-3), candidates_per_s=nr * kr / (pms * 1e-3), vehicles=nr,
                                            candidates=kr, ms_per_launch=pms, found_fraction=st["found"] / nr,
                                            roofline=dict(bound="hbm", achieved=gbs, peak=pk["hbm_gbs"], unit="GB/s", frac=gbs / pk["hbm_gbs"],
                                                          note="pixel scans of InflatePyramid: %.2f MB of pixels per plan in the reference algorithm; latency-bound" % (bytes_per_plan / 1e6)))
                # the reference planner on the host cores, same images / states / candidates (bounded sample)
                sys.path.insert(0, os.path.join(ROOT, "oracle"))
                import orc_rappids
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
                extras["rappids_c5"]["cpu_baseline"] = dict(value=ns * reps / dt, unit="plans/s", cores=os.cpu_count() or 1,
                                                            kind="reference" if fl.startswith("ref") else "port",
                                                            sample="%d x %d plans x %d candidates, %.2f s" % (reps, ns, kr, dt))
        except Exception as ex:
            extras["rappids_c5"] = dict(error=str(ex))
        line["extra"] = extras
        cb, _ = cpu_baseline(args, seconds=args.cpu_seconds)
        line["cpu_baseline"] = cb
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
