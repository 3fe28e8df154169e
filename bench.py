#!/usr/bin/env python
"""bench.py -- decoded depth frames/s of the fused structured-light kernel on B200.

    python bench.py --gpus N --steps K --warmup W            (N == 1)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...                     (the CPU path, host cores)

Workload: BASELINE.json configs[1] -- 1920x1200 camera, 9 Gray pairs (8 bits +
the complementary half-period bit) + 4-step phase shift, projector width 2560.
A "step" is one pass of the hot path (one fused kernel launch) over a batch of
`--batch` independent frame sets that are already resident in HBM; the batch is
far larger than L2 (126 MB), so no input byte is served from cache.  Frame sets
are sharded across ranks with no data-path collective (scaling: weak -- every
rank processes its own `--batch` frame sets per step).

`value`  = frame sets decoded per second, whole job, inputs resident in HBM.
`e2e`    = the same metric through the public host call
           (capi.Reconstructor.reconstruct_into -> slc_reconstruct_host) with
           pinned HOST buffers: H2D upload of every stack and D2H download of
           every XYZ map and mask inside the timed region, pipelined over
           stream slots.
`roofline` = algorithmic bytes (39 B/px: 22 u8 planes read + float4 XYZ + u8
           mask written) per launch / mean launch duration (CUDA events on the
           launching stream), against the measured HBM copy bandwidth.
`cpu_baseline` = the reference's own compiled sources (oracle/_ref/dynaframe_ref, kind
           "reference"; one single-threaded process per host thread) timed on this box's
           host cores on a bounded sample, rank 0, N == 1, with the oracle port's 1-thread
           and OpenMP numbers beside it; kind "port" where the reference cannot run the
           geometry (N != 4, modulation mask) or its binary is absent.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from structured_light_calculation_b200 import synth  # noqa: E402
from structured_light_calculation_b200.calibration import load_calibration  # noqa: E402
from structured_light_calculation_b200.configs import CONFIGS  # noqa: E402
from structured_light_calculation_b200 import distributed as D  # noqa: E402

METRIC = "decoded_depth_frames_per_sec"
UNIT = "frames/s"
FALLBACK_HBM_GBS = 6650.0   # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback"


def ncu_traffic_per_stack(cfg_name):
    """DRAM bytes per frame set from the committed ncu --set full capture, if any."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            return float(json.load(f)[cfg_name]["dram_bytes_per_stack"])
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "50", "-f", self.path], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            with open(self.path) as f:
                for line in f:
                    parts = [p.strip() for p in line.split(",")]
                    if len(parts) < 9:
                        continue
                    try:
                        sm.append(float(parts[1]))
                        mx.append(float(parts[2]))
                    except ValueError:
                        continue
                    for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                         parts[5:9]):
                        if val.lower().startswith("active"):
                            reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def host_threads() -> int:
    """Host threads the CPU path may use: the cores this process may run on (torchrun
    exports OMP_NUM_THREADS=1, which the oracle's explicit num_threads() overrides)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def build_inputs(cfg, pool: int):
    base = load_calibration(os.path.join(ROOT, "tests", "golden", "Result.yml"))
    cal = synth.synthetic_calibration(cfg, base)
    scene = synth.make_scene(cfg, cal)
    stacks = [synth.render_stack(cfg, scene, noise_sigma=1.0, seed=1234 + i) for i in range(pool)]
    return cal, scene, stacks


def reference_processes(cfg, cal, stack, procs: int, reps: int, skip: int):
    """`procs` concurrent copies of oracle/_ref/dynaframe_ref -- the reference's OWN path sources
    compiled in place (oracle/Makefile) -- each timing `reps` x (FillFirstProjectorU + FillCoordinate(0))
    on this stack.  The reference is single-threaded, so one process per host thread is all the
    parallelism it can use.  Returns per-process lists of seconds per repetition (first `skip` dropped:
    the first repetition also reads the .bmp files)."""
    from oracle import ref_runner as R
    ws = R.Workspace(cfg, cal, stack)
    try:
        env = ws.env(1)
        ps = [subprocess.Popen([R.BINARY, "time", str(reps + skip)], env=env, stdout=subprocess.PIPE,
                               stderr=subprocess.DEVNULL, text=True) for _ in range(procs)]
        out = []
        for p in ps:
            txt, _ = p.communicate()
            if p.returncode != 0:
                raise RuntimeError(f"dynaframe_ref exited {p.returncode}")
            line = [ln for ln in txt.splitlines() if ln.startswith('{"seconds"')][-1]
            out.append(json.loads(line)["seconds"][skip:])
        return out
    finally:
        ws.close()


def reference_available(cfg) -> bool:
    from oracle import ref_runner as R
    return cfg.phase_steps == 4 and cfg.modulation_min == 0 and R.available()   # the reference reads exactly 4 phase images


def run_reference(args, cfg):
    """--impl reference: the reference's own CPU implementation of the path on the host cores --
    oracle/_ref (the reference's sources compiled in place) when it is there and the geometry is one
    the reference can run, else the oracle port with OpenMP."""
    rank, _, world = D.env_rank_world()
    if rank != 0:
        return 0
    threads = host_threads()
    cal, _, stacks = build_inputs(cfg, 1)
    per_step = args.ref_stacks_per_step
    t0 = time.perf_counter()
    if reference_available(cfg):
        runs = reference_processes(cfg, cal, stacks[0], threads, args.steps * per_step, max(1, args.warmup))
        value = sum(len(r) / sum(r) for r in runs)                  # frame sets/s summed over the concurrent processes
        total = max(sum(r) for r in runs)
        kind = "reference"
        note = ("the reference's own sources (CDecodeGray/CDecodePhase/CCalculation .cpp) compiled in place against "
                "a minimal OpenCV stand-in (oracle/ref_shim); it is single-threaded, so one process per host thread; "
                "FillFirstProjectorU + FillCoordinate(0) per frame set, images served from memory after the first read")
        sample = (f"{threads} concurrent processes x {args.steps * per_step} frame sets of {cfg.width}x{cfg.height}; "
                  f"single process {len(runs[0]) / sum(runs[0]):.2f} frame sets/s while all run")
    else:
        from oracle import sl_oracle as O
        ocfg = O.make_config(cfg.width, cfg.height, cfg.projector_width, cfg.gray_digits, cfg.phase_steps,
                             cfg.fov_min, cfg.fov_max, cfg.modulation_min, threads)
        ocal = O.make_calib(cal.cam, cal.pro, cal.R, cal.T)
        O.time_reconstruct(ocfg, ocal, stacks[0], max(1, args.warmup))          # warm-up
        secs = O.time_reconstruct(ocfg, ocal, stacks[0], args.steps * per_step)
        total = float(secs.sum())
        value = args.steps * per_step / total
        kind = "port"
        note = ("CPU oracle port of the reference loops with OpenMP rows (the reference itself reads exactly four "
                "phase images and has no modulation mask, or its binary is not built); hot loops only, no I/O")
        sample = f"{args.steps * per_step} x one {cfg.width}x{cfg.height} stack, OpenMP {threads} threads"
    wall = time.perf_counter() - t0
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(cfg), "frame_sets_per_step": per_step, "note": note},
        "mpix_per_s": value * cfg.pixels / 1e6,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": wall,
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_name(cfg):
    which = {"config1": "configs[0]", "config2": "configs[1]", "config3": "configs[2]", "config4": "configs[3]", "config5": "configs[4]",
             "reference_default": "none: the reference's own StaticParameters"}.get(cfg.name, cfg.name)
    extra = ", modulation mask" if cfg.modulation_min > 0 else ""
    return (f"{cfg.width}x{cfg.height} stack, {cfg.gray_digits} Gray pairs (LSB = the complementary half-period bit) + "
            f"{cfg.phase_steps}-step phase shift{extra}, projector width {cfg.projector_width} (BASELINE {which})")


def cpu_baseline(cfg, cal, stack):
    """The CPU path on this box's host cores, bounded to some tens of seconds: the reference's own
    compiled sources (one process per host thread -- it is single-threaded) when available, and the
    oracle port (1 thread / OpenMP) beside it."""
    from oracle import sl_oracle as O
    ocal = O.make_calib(cal.cam, cal.pro, cal.R, cal.T)
    threads = host_threads()

    def timed(nthreads, reps):
        ocfg = O.make_config(cfg.width, cfg.height, cfg.projector_width, cfg.gray_digits, cfg.phase_steps,
                             cfg.fov_min, cfg.fov_max, cfg.modulation_min, nthreads)
        O.time_reconstruct(ocfg, ocal, stack, 1)
        return float(np.median(O.time_reconstruct(ocfg, ocal, stack, reps)))

    t1 = timed(1, 8)
    tn = timed(threads, 30) if threads > 1 else t1
    out = {
        "value": 1.0 / tn, "unit": UNIT, "cores": threads, "kind": "port",
        "sample": f"one {cfg.width}x{cfg.height} stack: median of 30 reps on {threads} threads (OpenMP rows), "
                  f"8 reps on 1 thread; hot loops only",
        "port_single_thread_value": 1.0 / t1, "port_single_thread_ms_per_frame": 1e3 * t1,
        "port_all_cores_value": 1.0 / tn, "port_all_cores_ms_per_frame": 1e3 * tn,
    }
    if reference_available(cfg):
        solo = reference_processes(cfg, cal, stack, 1, 6, 1)[0]
        runs = reference_processes(cfg, cal, stack, threads, 6, 1)
        agg = sum(len(r) / sum(r) for r in runs)
        out.update({
            "value": agg, "kind": "reference",
            "sample": f"the reference's own compiled sources (oracle/_ref): {threads} concurrent single-threaded "
                      f"processes x 6 frame sets of {cfg.width}x{cfg.height} (FillFirstProjectorU + FillCoordinate(0), "
                      f"images in memory); one process alone: 6 frame sets",
            "reference_single_process_value": len(solo) / sum(solo),
            "reference_single_process_ms_per_frame": 1e3 * sum(solo) / len(solo),
        })
    return out


def run_dynamic(args):
    """--path dynamic: the reference's CalculateOther mode -- a sequence of
    single stripe images tracked frame to frame (StripRegression + FillOtherDeltaProU +
    FillCoordinate), at the reference's own geometry (1280x1024, 100 frames, window 21).
    A step is `--batch // 64` sequences; every sequence is two kernel launches."""
    import torch
    from structured_light_calculation_b200 import capi
    from oracle import sl_oracle as O   # U0 for the synthetic sequence + CPU baseline only

    rank, local_rank, world = D.env_rank_world()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        D.init_process_group("nccl")
    cfg = CONFIGS["reference_default"]
    F, window, S = args.dyna_frames, 21, max(1, args.batch // 64)
    base = load_calibration(os.path.join(ROOT, "tests", "golden", "Result.yml"))
    cal = synth.synthetic_calibration(cfg, base)
    scene = synth.make_scene(cfg, cal)
    stack = synth.render_stack(cfg, scene, noise_sigma=1.0, seed=77)
    pool = synth.render_dyna_frames(cfg, cal, 8, stripe_period=20.0, z_step=0.3, noise_sigma=1.5)
    # 8 rendered positions visited back and forth (0..7,6..1,0..): the plane oscillates
    order = [k if k < 8 else 14 - k for k in (f % 14 for f in range(F))]
    frames = np.stack([pool[k] for k in order])
    rec = capi.Reconstructor(cfg, device=local_rank, max_batch=1, num_slots=1)
    rec.set_calibration(cal)
    u0 = rec.reconstruct(stack, parity=True)["proj_u"][0]
    npx = cfg.pixels
    d_frames = torch.empty((S, F, cfg.height, cfg.width), dtype=torch.uint8, device=dev)
    for i in range(S):
        d_frames[i].copy_(torch.from_numpy(np.roll(frames, i, axis=0)))
    d_u0 = torch.from_numpy(u0).to(dev)
    d_xyzw = torch.empty((S, F - 1, cfg.height, cfg.width, 4), dtype=torch.float32, device=dev)
    d_mask = torch.empty((S, F - 1, cfg.height, cfg.width), dtype=torch.uint8, device=dev)
    d_dz = torch.empty((S, F - 1, cfg.height, cfg.width), dtype=torch.float32, device=dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)

    def step():
        for i in range(S):
            rec.dyna_track_device(d_frames[i].data_ptr(), F, d_u0.data_ptr(), d_xyzw[i].data_ptr(),
                                  d_mask[i].data_ptr(), d_dz[i].data_ptr(), window, stream.cuda_stream)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    launches0 = rec.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    D.barrier()
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    D.barrier()
    ms = D.max_over_ranks(e0.elapsed_time(e1), dev)
    launches = int(D.sum_over_ranks(rec.launch_count() - launches0, dev))
    value = world * S * (F - 1) * args.steps / (ms * 1e-3)

    # end to end: host frames in, host maps out (one blocking call per sequence), pinned host buffers
    E = args.dyna_e2e_frames
    h_frames = capi.PinnedArray((E, cfg.height, cfg.width), np.uint8)
    h_frames.array[...] = frames[:E]
    h_u0 = capi.PinnedArray((cfg.height, cfg.width), np.float64)
    h_u0.array[...] = u0
    h_xyzw = capi.PinnedArray((E - 1, cfg.height, cfg.width, 4), np.float32)
    h_mask = capi.PinnedArray((E - 1, cfg.height, cfg.width), np.uint8)
    h_dz = capi.PinnedArray((E - 1, cfg.height, cfg.width), np.float32)
    rec.dyna_track_into(h_frames, E, h_u0, h_xyzw, h_mask, h_dz, window)
    D.barrier()
    t0 = time.perf_counter()
    e2e_reps = 5
    for _ in range(e2e_reps):
        rec.dyna_track_into(h_frames, E, h_u0, h_xyzw, h_mask, h_dz, window)
    e2e_s = D.max_over_ranks(time.perf_counter() - t0, dev)
    e2e_value = world * e2e_reps * (E - 1) / e2e_s
    out = {"xyzw": h_xyzw.array, "mask": h_mask.array}

    if rank == 0:
        ocfg = O.make_config(cfg.width, cfg.height, cfg.projector_width, cfg.gray_digits, cfg.phase_steps)
        ocal = O.make_calib(cal.cam, cal.pro, cal.R, cal.T)
        first = O.reconstruct(ocfg, ocal, stack)
        t0 = time.perf_counter()
        want = O.dyna_sequence(ocfg, ocal, first["proj_u"], first["z"], frames[:4], window)
        cpu_s = (time.perf_counter() - t0) / 3
        tol = 1e-5 * (cfg.fov_max - cfg.fov_min)
        checked = all(np.array_equal(out["mask"][f], want[f]["mask"]) and
                      np.array_equal(out["xyzw"][f, ..., 3], want[f]["proj_u"].astype(np.float32)) and
                      np.abs(out["xyzw"][f, ..., 2] - want[f]["z"]).max() <= tol for f in range(3))
        peak, peak_kind = hbm_peak()
        alg = 22 * npx * (F - 1) * S          # 1 B image in; float4 XYZ + u8 mask + f32 deltaZ out (intermediates excluded)
        achieved = alg / (ms * 1e-3 / args.steps) / 1e9
        line = {
            "metric": "dynamic_frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"dynamic sequence: {F} stripe images {cfg.width}x{cfg.height}, window {window} "
                                   f"(reference CalculateOther), {S} sequence(s) per step"},
            "mpix_per_s": value * npx / 1e6,
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": args.dyna_e2e_frames * npx,
                    "d2h_bytes_per_step": (args.dyna_e2e_frames - 1) * npx * 21,
                    "api": "capi.Reconstructor.dyna_track_into -> slc_dyna_track_host, pinned host buffers"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_kind": f"of {peak_kind}",
                         "kernel": "strip_regression21_kernel + dyna_fused_kernel (per sequence)",
                         "algorithmic_bytes_per_step": alg},
            "cpu_baseline": {"value": 1.0 / cpu_s, "unit": "frames/s", "cores": 1, "kind": "port",
                             "sample": "3 dynamic frames, oracle port, 1 thread"},
            "checked_against_oracle": bool(checked),
        }
        print(json.dumps(line), flush=True)
    rec.close()
    return 0


def run_pointcloud(args):
    """--path pointcloud: CCalculation::Result -- the text cloud of one
    1920x1200 frame (BASELINE configs[1] geometry) formatted on the device from the f64
    ProjectorU plane.  A step is `--pc-frames` frames, two kernel launches each."""
    import torch
    from structured_light_calculation_b200 import capi
    from oracle import sl_oracle as O   # CPU baseline + byte check only

    rank, local_rank, world = D.env_rank_world()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        D.init_process_group("nccl")
    cfg = CONFIGS[args.config]
    cal, scene, stacks = build_inputs(cfg, 1)
    rec = capi.Reconstructor(cfg, device=local_rank, max_batch=1, num_slots=1)
    rec.set_calibration(cal)
    first = rec.reconstruct(stacks[0], parity=True)
    u_host = first["proj_u"][0]
    npx = cfg.pixels
    F = args.pc_frames
    d_u = torch.from_numpy(u_host).to(dev)
    cap = 43 * npx + 16
    d_text = torch.empty((cap,), dtype=torch.uint8, device=dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    nbytes = npts = 0

    def step():
        nonlocal nbytes, npts
        for _ in range(F):
            nbytes, npts = rec.pointcloud_text_device(d_u.data_ptr(), d_text.data_ptr(), cap, 0, stream.cuda_stream)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    launches0 = rec.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    D.barrier()
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    D.barrier()
    ms = D.max_over_ranks(e0.elapsed_time(e1), dev)
    launches = int(D.sum_over_ranks(rec.launch_count() - launches0, dev))
    value = world * F * args.steps / (ms * 1e-3)

    # the binary companion (float3 of the valid pixels of the first-frame map, reference order), same timing
    d_xyzw = torch.from_numpy(first["xyzw"][0]).to(dev)
    d_msk = torch.from_numpy(first["mask"][0]).to(dev)
    d_xyz = torch.empty((npx, 3), dtype=torch.float32, device=dev)
    n_compact = 0
    for _ in range(3):
        n_compact = rec.pointcloud_compact_device(d_xyzw.data_ptr(), d_msk.data_ptr(), d_xyz.data_ptr(), npx,
                                                  capi.SLC_ORDER_REFERENCE, stream.cuda_stream)
    torch.cuda.synchronize()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record(stream)
    for _ in range(F * args.steps):
        rec.pointcloud_compact_device(d_xyzw.data_ptr(), d_msk.data_ptr(), d_xyz.data_ptr(), npx,
                                      capi.SLC_ORDER_REFERENCE, stream.cuda_stream)
    c1.record(stream)
    torch.cuda.synchronize()
    compact_ms = c0.elapsed_time(c1) / (F * args.steps)
    xyz_host = d_xyz[:n_compact].cpu().numpy()
    m = first["mask"][0].T.astype(bool)                         # reference order: u outer, v inner
    compact_ok = bool(n_compact == int(m.sum()) and
                      np.array_equal(xyz_host, np.transpose(first["xyzw"][0][..., :3], (1, 0, 2))[m]))

    # end to end: host f64 plane in, host text out
    e2e_reps = 10
    h_u = capi.PinnedArray((cfg.height, cfg.width), np.float64)
    h_u.array[...] = u_host
    h_text = capi.PinnedArray((cap,), np.uint8)
    rec.pointcloud_text_into(h_u, h_text)
    t0 = time.perf_counter()
    for _ in range(e2e_reps):
        nb, _n = rec.pointcloud_text_into(h_u, h_text)
    e2e_s = D.max_over_ranks(time.perf_counter() - t0, dev)
    e2e_value = world * e2e_reps / e2e_s
    text = h_text.array[:nb].tobytes()

    if rank == 0:
        ocfg = O.make_config(cfg.width, cfg.height, cfg.projector_width, cfg.gray_digits, cfg.phase_steps,
                             cfg.fov_min, cfg.fov_max, cfg.modulation_min)
        want = O.reconstruct(ocfg, O.make_calib(cal.cam, cal.pro, cal.R, cal.T), stacks[0])
        t0 = time.perf_counter()
        wtext, wn = O.result_text(ocfg, want["x"], want["y"], want["z"])
        cpu_s = time.perf_counter() - t0
        peak, peak_kind = hbm_peak()
        alg = (8 * npx + nbytes) * F            # f64 ProjectorU read once + the text written once
        achieved = alg / (ms * 1e-3 / args.steps) / 1e9
        line = {
            "metric": "pointcloud_text_frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"Result() text cloud of a {cfg.width}x{cfg.height} frame ({npts} points, "
                                   f"{nbytes} bytes), {F} frame(s) per step"},
            "binary_cloud": {"frames_per_s": 1e3 / compact_ms, "points": n_compact, "bytes": 12 * n_compact,
                             "gb_per_s": (17 * npx + 12 * n_compact) / (compact_ms * 1e-3) / 1e9,
                             "api": "slc_pointcloud_compact_device (float3 of the valid pixels, reference order)",
                             "checked": compact_ok},
            "points_per_s": value * npts, "text_gb_per_s": value * nbytes / 1e9,
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": 8 * npx, "d2h_bytes_per_step": nbytes,
                    "api": "capi.Reconstructor.pointcloud_text_into -> slc_pointcloud_text_host, pinned host buffers"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_kind": f"of {peak_kind}",
                         "kernel": "pc_emit_kernel<0,false> + pc_emit_kernel<0,true> (per frame)",
                         "algorithmic_bytes_per_step": alg,
                         "note": "pass 1 (x, y, z and their exact digits, f64) is FP64/conversion bound, pass 2 (characters) integer-issue / shared-store bound; neither is HBM bound"},
            "cpu_baseline": {"value": 1.0 / cpu_s, "unit": "frames/s", "cores": 1, "kind": "port",
                             "sample": "one frame, oracle Result() restatement (snprintf %g), 1 thread, no file I/O"},
            "checked_against_oracle": bool(text == wtext and wn == npts),
        }
        print(json.dumps(line), flush=True)
    rec.close()
    return 0


def run_ingest(args):
    """--path ingest: CSensor::LoadDatas -- the 2G+N .bmp files of one frame set
    (reference file layout, 8-bit gray palette, tmpfs) read, uploaded and unpacked on the device
    straight into the plane-major stack.  value = frame sets/s of the batched unpack (one
    slc_bmp_unpack_batch_device call per step, 64 files per launch) on raw pixel arrays already resident in HBM; e2e = files ->
    device stack through slc_load_bmp_planes."""
    import shutil
    import torch
    from structured_light_calculation_b200 import capi
    from oracle.bmp_oracle import decode_bmp_gray      # CPU baseline + check only

    rank, local_rank, world = D.env_rank_world()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        D.init_process_group("nccl")
    cfg = CONFIGS[args.config]
    cal, scene, stacks = build_inputs(cfg, 1)
    planes = stacks[0]
    tmp = tempfile.mkdtemp(prefix="slc_ingest_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        paths = synth.write_reference_layout(os.path.join(tmp, "group"), cfg, planes)
        rec = capi.Reconstructor(cfg, device=local_rank, max_batch=1, num_slots=1)
        rec.set_calibration(cal)
        npx, P = cfg.pixels, cfg.planes
        d_stack = torch.empty((P, cfg.height, cfg.width), dtype=torch.uint8, device=dev)
        files = [open(p, "rb").read() for p in paths]
        infos = [capi.bmp_parse(f) for f in files]
        d_raw = [torch.frombuffer(bytearray(f[i.pixel_offset:]), dtype=torch.uint8).to(dev) for f, i in zip(files, infos)]
        stream = torch.cuda.Stream(device=dev)
        torch.cuda.set_stream(stream)
        R = args.ingest_reps

        # R distinct copies of the set and R output stacks: 2 x R x 50.7 MB, well beyond the 126 MB L2
        raw_sets = [[t.clone() for t in d_raw] for _ in range(R)]
        raw_ptrs = [[t.data_ptr() for t in rs] for rs in raw_sets]
        d_stacks = torch.empty((R, P, cfg.height, cfg.width), dtype=torch.uint8, device=dev)

        all_ptrs = [q for rp in raw_ptrs for q in rp]
        all_infos = infos * R

        def step():
            # one call for the R frame sets of the step (R x P files, 64 files per launch)
            rec.bmp_unpack_batch_device(all_ptrs, all_infos, d_stacks.data_ptr(), stream.cuda_stream)

        for _ in range(args.warmup):
            step()
        torch.cuda.synchronize()
        launches0 = rec.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        D.barrier()
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(args.steps):
            step()
        e1.record(stream)
        torch.cuda.synchronize()
        D.barrier()
        ms = D.max_over_ranks(e0.elapsed_time(e1), dev)
        launches = int(D.sum_over_ranks(rec.launch_count() - launches0, dev))
        value = world * R * args.steps / (ms * 1e-3)
        ok = all(bool(np.array_equal(d_stacks[r].cpu().numpy(), planes)) for r in (0, R - 1))

        reps = 5
        rec.load_bmp_planes(paths, d_stack.data_ptr())
        t0 = time.perf_counter()
        for _ in range(reps):
            rec.load_bmp_planes(paths, d_stack.data_ptr())
        e2e_s = D.max_over_ranks(time.perf_counter() - t0, dev)
        ok = ok and bool(np.array_equal(d_stack.cpu().numpy(), planes))
        if rank == 0:
            t0 = time.perf_counter()
            for f in files[:6]:
                decode_bmp_gray(f)
            cpu_s = (time.perf_counter() - t0) * P / 6
            peak, peak_kind = hbm_peak()
            alg = 2 * npx * P * R                     # every pixel byte read once and written once
            achieved = alg / (ms * 1e-3 / args.steps) / 1e9
            line = {
                "metric": "ingest_frame_sets_per_sec", "value": value, "unit": "frame sets/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": f"{P} x {cfg.width}x{cfg.height} 8-bit .bmp (reference file layout) -> plane-major "
                                       f"device stack, {R} frame set(s) per step"},
                "e2e": {"value": world * reps / e2e_s, "unit": "frame sets/s", "h2d_bytes_per_step": sum(len(f) for f in files),
                        "d2h_bytes_per_step": 0, "api": "capi.Reconstructor.load_bmp_planes -> slc_load_bmp_planes (tmpfs files)"},
                "gpu_launches": launches,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": None, "peak_kind": f"of {peak_kind}", "kernel": "bmp_unpack_batch_kernel (64 files per launch)",
                             "algorithmic_bytes_per_step": alg,
                             "note": f"{R} distinct frame sets per step ({2 * R * npx * P / 1e6:.0f} MB of traffic, beyond L2)"},
                "cpu_baseline": {"value": 1.0 / cpu_s, "unit": "frame sets/s", "cores": 1, "kind": "port",
                                 "sample": "6 files through the numpy restatement of imread's BMP decoder, scaled to one frame set"},
                "checked_against_oracle": ok,
            }
            print(json.dumps(line), flush=True)
        rec.close()
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return 0


def run_app(args):
    """--path app: the whole reference program (main.cpp:42-45: Init, CalculateFirst, CalculateOther with
    its file reads and text clouds) against the same program on this library
    (examples/dynaframe_main.cpp -> include/dynaframe_b200.hpp), on the same input files, at the
    reference's own geometry.  The two sets of text clouds are compared byte for byte."""
    from oracle import ref_runner as R          # runs the reference binary: the baseline of this mode
    rank, _, world = D.env_rank_world()
    if rank != 0:
        return 0
    exe = os.path.join(ROOT, "structured_light_calculation_b200", "bin", "dynaframe_main")
    if not os.path.exists(exe):
        raise SystemExit("bench.py: build the library first (python -c 'import __graft_entry__ as g; g.build()')")
    if not R.available():
        print(json.dumps({"metric": "app_frames_per_sec", "unavailable": "oracle/_ref/dynaframe_ref not built"}))
        return 0
    cfg = CONFIGS["reference_default"]
    n = args.app_frames
    base = load_calibration(os.path.join(ROOT, "tests", "golden", "Result.yml"))
    cal = synth.synthetic_calibration(cfg, base)
    scene = synth.make_scene(cfg, cal)
    stack = synth.render_stack(cfg, scene, noise_sigma=1.0, seed=77)
    pool = synth.render_dyna_frames(cfg, cal, min(n, 8), stripe_period=20.0, z_step=0.3, noise_sigma=1.5)
    frames = np.stack([pool[k if k < len(pool) else 2 * len(pool) - 2 - k]
                       for k in (f % max(1, 2 * len(pool) - 2) for f in range(n))])
    tmp = tempfile.mkdtemp(prefix="slc_app_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        ws = R.Workspace(cfg, cal, stack, frames, root=tmp)
        out = os.path.join(tmp, "ours")
        os.makedirs(out)
        cmd = [exe, ws.data, str(cfg.width), str(cfg.height), str(cfg.projector_width), str(cfg.gray_digits),
               str(cfg.phase_steps), str(n), out]
        subprocess.run(cmd, cwd=ws.cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)      # warm-up (driver, page cache)
        t0 = time.perf_counter()
        res = subprocess.run(cmd, cwd=ws.cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        ours_s = time.perf_counter() - t0
        if res.returncode != 0:
            raise SystemExit(f"dynaframe_main failed ({res.returncode}): {res.stdout[-500:]} {res.stderr[-500:]}")
        phases = json.loads([ln for ln in res.stdout.splitlines() if ln.startswith("{")][-1])
        ref = ws.run_app()
        ref_s = ref["seconds"]
        same, nbytes = True, 0
        for f in range(n):
            name = "iFrame.txt" if f == 0 else f"cFrame{f}.txt"
            with open(os.path.join(out, name), "rb") as fh:
                mine = fh.read()
            same = same and (mine == ref["clouds"][f])
            nbytes += len(mine)
        line = {
            "metric": "app_frames_per_sec", "value": n / ours_s, "unit": "frames/s", "n_gpus": 1, "steps": 1, "warmup": 1,
            "ms_per_step": 1e3 * ours_s, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"whole program: {cfg.planes} pattern .bmp + {n} dynaCam .bmp of {cfg.width}x{cfg.height} in, "
                                   f"{n} text clouds ({nbytes} bytes) out; process start to exit, files on tmpfs"},
            "phases_s": phases,
            "e2e": {"value": n / ours_s, "unit": "frames/s", "h2d_bytes_per_step": (cfg.planes + n) * cfg.pixels,
                    "d2h_bytes_per_step": nbytes, "api": "examples/dynaframe_main.cpp (CCalculation::Init / CalculateFirst / "
                                                         "CalculateOther / Result)"},
            "cpu_baseline": {"value": n / ref_s, "unit": "frames/s", "cores": 1, "kind": "reference",
                             "sample": f"oracle/_ref/dynaframe_ref full: the reference's own Init + CalculateFirst + CalculateOther "
                                       f"on the same files ({ref_s:.2f} s, process start to exit)"},
            "speedup_vs_reference": ref_s / ours_s,
            "clouds_byte_identical": bool(same),
        }
        print(json.dumps(line), flush=True)
    finally:
        import shutil
        shutil.rmtree(tmp, ignore_errors=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="config2", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=256, help="frame sets per step (device-resident batch)")
    ap.add_argument("--pool", type=int, default=4, help="distinct rendered stacks tiled into the batch")
    ap.add_argument("--e2e-stacks", type=int, default=24, help="frame sets per end-to-end step")
    ap.add_argument("--e2e-chunk", type=int, default=2, help="frame sets per upload/launch/download chunk")
    ap.add_argument("--e2e-slots", type=int, default=4)
    ap.add_argument("--pxt", type=int, default=0, help="tuning: pixels per thread (4/8/16)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-numa-bind", action="store_true")
    ap.add_argument("--path", default="first", choices=["first", "dynamic", "pointcloud", "ingest", "app"],
                    help="first = the headline first-frame path; dynamic = CalculateOther sequences; "
                         "pointcloud = Result() text formatting; ingest = .bmp files -> device stack")
    ap.add_argument("--ingest-reps", type=int, default=8)
    ap.add_argument("--app-frames", type=int, default=16, help="--path app: first frame + this many - 1 dynamic frames")
    ap.add_argument("--pc-frames", type=int, default=4)
    ap.add_argument("--dyna-frames", type=int, default=100)
    ap.add_argument("--dyna-e2e-frames", type=int, default=24)
    ap.add_argument("--ref-stacks-per-step", type=int, default=2)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)
    cfg = CONFIGS[args.config]

    if args.impl == "reference":
        return run_reference(args, cfg)
    if args.path == "dynamic":
        return run_dynamic(args)
    if args.path == "pointcloud":
        return run_pointcloud(args)
    if args.path == "ingest":
        return run_ingest(args)
    if args.path == "app":
        return run_app(args)

    import torch
    from structured_light_calculation_b200 import capi

    rank, local_rank, world = D.env_rank_world()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback "
                         "(use --impl reference for the CPU path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    orig_affinity = os.sched_getaffinity(0)
    numa = D.bind_to_gpu_numa_node(local_rank) if not args.no_numa_bind else {"bound": False}
    if world > 1:
        D.init_process_group("nccl")
    if args.pxt:
        capi.load_library().slc_tune_pixels_per_thread(args.pxt)

    F = args.batch
    cal, scene, stacks = build_inputs(cfg, args.pool)
    rec = capi.Reconstructor(cfg, device=local_rank, max_batch=args.e2e_chunk, num_slots=args.e2e_slots)
    rec.set_calibration(cal)
    info = rec.info()

    # ---- device-resident batch (torch owns the memory; the kernel is ours) ----
    d_in = torch.empty((F, cfg.planes, cfg.height, cfg.width), dtype=torch.uint8, device=dev)
    d_xyzw = torch.empty((F, cfg.height, cfg.width, 4), dtype=torch.float32, device=dev)
    d_mask = torch.empty((F, cfg.height, cfg.width), dtype=torch.uint8, device=dev)
    pool_dev = [torch.from_numpy(s).to(dev) for s in stacks]
    for i in range(F):
        d_in[i].copy_(pool_dev[i % len(pool_dev)])
    del pool_dev
    torch.cuda.synchronize()
    # a dedicated (non-default) torch stream: the kernel is launched on it and the
    # torch.cuda.Events that time it are recorded on it
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)

    def step():
        rec.reconstruct_device(d_in.data_ptr(), F, d_xyzw.data_ptr(), d_mask.data_ptr(), None, stream.cuda_stream)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = rec.launch_count()
    events = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    D.barrier()
    torch.cuda.synchronize()
    events[0].record(stream)
    for i in range(args.steps):
        step()
        events[i + 1].record(stream)
    torch.cuda.synchronize()
    D.barrier()
    total_ms = events[0].elapsed_time(events[-1])
    per_launch_ms = [events[i].elapsed_time(events[i + 1]) for i in range(args.steps)]
    launches = rec.launch_count() - launches0
    total_ms_max = D.max_over_ranks(total_ms, dev)
    launches_all = int(D.sum_over_ranks(launches, dev))
    value = world * F * args.steps / (total_ms_max * 1e-3)

    # ---- end to end through the public host call, pinned host buffers ----
    E = args.e2e_stacks
    h_in = capi.PinnedArray((E, cfg.planes, cfg.height, cfg.width), np.uint8)
    h_xyzw = capi.PinnedArray((E, cfg.height, cfg.width, 4), np.float32)
    h_mask = capi.PinnedArray((E, cfg.height, cfg.width), np.uint8)
    for i in range(E):
        h_in.array[i] = stacks[i % len(stacks)]
    e2e_steps = args.steps
    for _ in range(2):
        rec.reconstruct_into(h_in, E, h_xyzw, h_mask)
    D.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        rec.reconstruct_into(h_in, E, h_xyzw, h_mask)     # blocking: returns with results in host memory
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    D.barrier()
    e2e_s_max = D.max_over_ranks(e2e_s, dev)
    e2e_value = world * E * e2e_steps / e2e_s_max
    clocks = sampler.stop() if rank == 0 else None

    os.sched_setaffinity(0, orig_affinity)   # the CPU baseline below may use every core again
    # ---- spot check of what was just computed (not timed) ----
    checked = None
    cpu = None
    if rank == 0:
        from oracle import sl_oracle as O   # checker + CPU baseline only
        ocfg = O.make_config(cfg.width, cfg.height, cfg.projector_width, cfg.gray_digits, cfg.phase_steps,
                             cfg.fov_min, cfg.fov_max, cfg.modulation_min, host_threads())
        want = O.reconstruct(ocfg, O.make_calib(cal.cam, cal.pro, cal.R, cal.T), stacks[0])
        z_dev = d_xyzw[0, :, :, 2].cpu().numpy()
        m_dev = d_mask[0].cpu().numpy()
        tol = 1e-5 * (cfg.fov_max - cfg.fov_min)
        checked = bool(np.array_equal(m_dev, want["mask"]) and np.abs(z_dev - want["z"]).max() <= tol
                       and np.array_equal(h_mask.array[0], want["mask"])
                       and np.abs(h_xyzw.array[0, :, :, 2] - want["z"]).max() <= tol)
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_baseline(cfg, cal, stacks[0])

    if rank == 0:
        peak, peak_kind = hbm_peak()
        mean_launch_ms = statistics.fmean(per_launch_ms)
        alg_bytes = cfg.algorithmic_bytes_per_pixel * cfg.pixels * F
        achieved = alg_bytes / (mean_launch_ms * 1e-3) / 1e9
        traffic_ps = ncu_traffic_per_stack(args.config)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(cfg), "frame_sets_per_step_per_gpu": F,
                       "planes": cfg.planes, "bytes_per_pixel_algorithmic": cfg.algorithmic_bytes_per_pixel,
                       "l2_policy": f"inputs larger than L2: {F * cfg.stack_bytes / 1e9:.1f} GB read + "
                                    f"{F * cfg.pixels * 17 / 1e9:.1f} GB written per step",
                       "parallelism": f"frame sets sharded over {world} GPU(s), no collective",
                       "distinct_stacks_in_pool": len(stacks)},
            "mpix_per_s": value * cfg.pixels / 1e6,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": E * cfg.stack_bytes,
                    "d2h_bytes_per_step": E * cfg.pixels * 17, "frame_sets_per_step_per_gpu": E,
                    "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s_max / e2e_steps,
                    "api": "capi.Reconstructor.reconstruct_into -> slc_reconstruct_host, pinned host buffers, "
                           f"{args.e2e_slots} stream slots x {args.e2e_chunk} frame sets"},
            "gpu_launches": launches_all,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": (traffic_ps * F if traffic_ps else None),
                         "peak_kind": f"of {peak_kind}", "kernel": "slc::reconstruct_vec_kernel",
                         "algorithmic_bytes_per_launch": alg_bytes, "mean_launch_ms": mean_launch_ms,
                         "min_launch_ms": min(per_launch_ms), "max_launch_ms": max(per_launch_ms)},
            "cpu_baseline": cpu,
            "clocks": clocks,
            "kernel": {"variant": info.kernel_variant, "regs": info.kernel_regs, "block": info.kernel_block,
                       "smem": info.kernel_smem},
            "checked_against_oracle": checked,
            "numa": numa,
        }
        print(json.dumps(line), flush=True)
    rec.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
