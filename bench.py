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
    """DRAM bytes per frame set: dram__bytes_read.sum + dram__bytes_write.sum of one full-size launch in
    the committed ncu launch list of this same command (profiles/traffic.json), divided by its frame sets."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            e = json.load(f)[cfg_name]
        if "dram_bytes_per_launch" in e:
            return float(e["dram_bytes_per_launch"]) / float(e["launch_frame_sets"])
        return float(e["dram_bytes_per_stack"])
    except Exception:
        return None


def link_ceiling(n_gpus: int):
    """Probed host-link ceiling for GPUs 0..N-1 together (profiles/hostlink_ceiling.json, written by
    profiles/hostlink_probe.py on this pool's boxes): {h2d_gbs, d2h_gbs, bidir_gbs} or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "hostlink_ceiling.json")) as f:
            return json.load(f)["per_n"].get(str(n_gpus))
    except Exception:
        return None


def bench_config(cfg, args, world):
    """`config` of the JSON line: the workload and how it is run -- the SAME object in the b200 arm and the
    reference arm (both are launched with the same flags), so the driver's same_config holds."""
    return {"workload": workload_name(cfg), "planes": cfg.planes,
            "bytes_per_pixel_algorithmic": cfg.algorithmic_bytes_per_pixel,
            "frame_sets_per_step_per_gpu": args.batch,
            "l2_policy": f"inputs larger than L2: {args.batch * cfg.stack_bytes / 1e9:.1f} GB read + "
                         f"{args.batch * cfg.pixels * 17 / 1e9:.1f} GB written per step on the GPU arm (the CPU arm "
                         f"times a bounded sample of the same frame sets, see cpu_baseline.sample)",
            "parallelism": f"frame sets sharded over {world} GPU(s), no collective",
            "distinct_stacks_in_pool": args.pool}


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
        "config": bench_config(cfg, args, world),
        "mpix_per_s": value * cfg.pixels / 1e6,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample,
                         "frame_sets_per_step": per_step, "note": note},
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


def run_dynamic(args, emit=True):
    """--path dynamic: the reference's CalculateOther mode -- a sequence of
    single stripe images tracked frame to frame (StripRegression + FillOtherDeltaProU +
    FillCoordinate), at the reference's own geometry (1280x1024, 100 frames, window 21).
    A step is `--batch // 64` sequences; every sequence is two kernel launches."""
    import torch
    from structured_light_calculation_b200 import capi
    from oracle import sl_oracle as O   # U0 for the synthetic sequence + CPU baseline only

    rank, local_rank, world = D.env_rank_world()
    line = None
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
    # the same with the depth-only result (z + one bit per pixel instead of 21 B/px back over the link)
    bufs_d, res_d = capi.alloc_result(cfg, E - 1, capi.SLC_RESULT_DEPTH, pinned=True)
    rec.dyna_track_into_ex(h_frames, E, h_u0, res_d, window)
    D.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_reps):
        rec.dyna_track_into_ex(h_frames, E, h_u0, res_d, window)
    e2e_depth_s = D.max_over_ranks(time.perf_counter() - t0, dev)
    e2e_depth_value = world * e2e_reps * (E - 1) / e2e_depth_s
    depth_ok = bool(np.array_equal(bufs_d["depth"].array, h_xyzw.array[..., 2]) and
                    np.array_equal(capi.unpack_mask_bits(bufs_d["mask_bits"], E - 1, npx), h_mask.array.reshape(E - 1, npx)))

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
            "e2e_compact": {"depth": {"value": e2e_depth_value, "unit": "frames/s", "h2d_bytes_per_step": args.dyna_e2e_frames * npx,
                                      "d2h_bytes_per_step": (args.dyna_e2e_frames - 1) * (npx * 4 + capi.bits_bytes(npx)),
                                      "api": "slc_dyna_track_host_ex SLC_RESULT_DEPTH (z + bit mask per frame)",
                                      "checked_bit_equal_to_full_map": depth_ok}},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_kind": f"of {peak_kind}",
                         "kernel": "strip_regression21_kernel + dyna_fused_kernel (per sequence)",
                         "algorithmic_bytes_per_step": alg},
            "cpu_baseline": {"value": 1.0 / cpu_s, "unit": "frames/s", "cores": 1, "kind": "port",
                             "sample": "3 dynamic frames, oracle port, 1 thread"},
            "checked_against_oracle": bool(checked),
        }
        if emit:
            print(json.dumps(line), flush=True)
    rec.close()
    return line if rank == 0 else None


def run_pointcloud(args, emit=True):
    """--path pointcloud: CCalculation::Result -- the text cloud of one
    1920x1200 frame (BASELINE configs[1] geometry) formatted on the device from the f64
    ProjectorU plane.  A step is `--pc-frames` frames, one kernel launch each."""
    import torch
    from structured_light_calculation_b200 import capi
    from oracle import sl_oracle as O   # CPU baseline + byte check only

    rank, local_rank, world = D.env_rank_world()
    line = None
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

    # the binary companion: float3 of the valid pixels in Result()'s order, B maps per asynchronous launch
    # (slc_compact_points_device: one chained-scan launch, counts stay on the device); B distinct copies of the
    # map so that nothing is served from L2 (B x 39 MB in, B x 26 MB out)
    B = 16
    d_xyzw = torch.from_numpy(first["xyzw"][0]).to(dev).unsqueeze(0).repeat(B, 1, 1, 1)
    d_msk = torch.from_numpy(first["mask"][0]).to(dev).unsqueeze(0).repeat(B, 1, 1)
    d_xyz = torch.empty((B, npx, 3), dtype=torch.float32, device=dev)
    d_cnt = torch.zeros(B, dtype=torch.int64, device=dev)
    l1 = rec.launch_count()
    for _ in range(3):
        rec.compact_points_device(d_xyzw.data_ptr(), d_msk.data_ptr(), B, d_xyz.data_ptr(), npx, d_cnt.data_ptr(),
                                  capi.SLC_ORDER_REFERENCE, None, stream.cuda_stream)
    torch.cuda.synchronize()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record(stream)
    creps = max(4, args.steps)
    for _ in range(creps):
        rec.compact_points_device(d_xyzw.data_ptr(), d_msk.data_ptr(), B, d_xyz.data_ptr(), npx, d_cnt.data_ptr(),
                                  capi.SLC_ORDER_REFERENCE, None, stream.cuda_stream)
    c1.record(stream)
    torch.cuda.synchronize()
    compact_ms = c0.elapsed_time(c1) / (creps * B)
    # the same lists in the maps' memory order (the other tile shape of slc_compact.cu)
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rec.compact_points_device(d_xyzw.data_ptr(), d_msk.data_ptr(), B, d_xyz.data_ptr(), npx, d_cnt.data_ptr(),
                              capi.SLC_ORDER_ROW_MAJOR, None, stream.cuda_stream)
    r0.record(stream)
    for _ in range(creps):
        rec.compact_points_device(d_xyzw.data_ptr(), d_msk.data_ptr(), B, d_xyz.data_ptr(), npx, d_cnt.data_ptr(),
                                  capi.SLC_ORDER_ROW_MAJOR, None, stream.cuda_stream)
    r1.record(stream)
    torch.cuda.synchronize()
    compact_rm_ms = r0.elapsed_time(r1) / (creps * B)
    rm_ok = bool(np.array_equal(d_xyz[0, :int(d_cnt[0].item())].cpu().numpy(),
                                first["xyzw"][0][..., :3][first["mask"][0].astype(bool)]))
    rec.compact_points_device(d_xyzw.data_ptr(), d_msk.data_ptr(), B, d_xyz.data_ptr(), npx, d_cnt.data_ptr(),
                              capi.SLC_ORDER_REFERENCE, None, stream.cuda_stream)      # the checks below read this order
    torch.cuda.synchronize()
    compact_launches = (rec.launch_count() - l1) // (2 * creps + 5)
    n_compact = int(d_cnt[B - 1].item())
    xyz_host = d_xyz[B - 1, :n_compact].cpu().numpy()
    m = first["mask"][0].T.astype(bool)                         # reference order: u outer, v inner
    compact_ok = bool(n_compact == int(m.sum()) and bool((d_cnt == n_compact).all().item()) and
                      np.array_equal(xyz_host, np.transpose(first["xyzw"][0][..., :3], (1, 0, 2))[m]))
    # one synchronous call (returns the count): launch + 8-byte read-back + stream synchronisation
    t0 = time.perf_counter()
    for _ in range(20):
        rec.pointcloud_compact_device(d_xyzw.data_ptr(), d_msk.data_ptr(), d_xyz.data_ptr(), npx,
                                      capi.SLC_ORDER_REFERENCE, stream.cuda_stream)
    compact_sync_ms = (time.perf_counter() - t0) * 1e3 / 20
    del d_xyzw, d_msk, d_xyz

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
            "binary_cloud": {"frames_per_s": 1e3 / compact_ms, "us_per_frame": 1e3 * compact_ms, "points": n_compact,
                             "bytes": 12 * n_compact,
                             "gb_per_s": (17 * npx + 12 * n_compact) / (compact_ms * 1e-3) / 1e9,
                             "roofline_frac": (17 * npx + 12 * n_compact) / (compact_ms * 1e-3) / 1e9 / peak,
                             "launches_per_call": compact_launches, "maps_per_launch": B,
                             "row_major_us_per_frame": 1e3 * compact_rm_ms,
                             "row_major_roofline_frac": (17 * npx + 12 * n_compact) / (compact_rm_ms * 1e-3) / 1e9 / peak,
                             "row_major_checked": rm_ok,
                             "synchronous_call_us": 1e3 * compact_sync_ms,
                             "api": "slc_compact_points_device: float3 of the valid pixels in Result()'s order (u outer, v inner), "
                                    "one chained-scan launch for 16 maps, counts stay on the device; synchronous_call_us = one "
                                    "slc_pointcloud_compact_device call (launch + count read-back)",
                             "checked": compact_ok},
            "points_per_s": value * npts, "text_gb_per_s": value * nbytes / 1e9,
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": 8 * npx, "d2h_bytes_per_step": nbytes,
                    "api": "capi.Reconstructor.pointcloud_text_into -> slc_pointcloud_text_host, pinned host buffers"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_kind": f"of {peak_kind}",
                         "kernel": "pc_text_kernel (one chained-scan launch per frame)",
                         "algorithmic_bytes_per_step": alg,
                         "note": "phase A (x, y, z and their exact digits, f64) is FP64 / conversion latency bound, phase B (characters) integer-issue / shared-store bound; neither is HBM bound (DRAM traffic 18.5 MB read per frame, the text stays in L2)"},
            "cpu_baseline": {"value": 1.0 / cpu_s, "unit": "frames/s", "cores": 1, "kind": "port",
                             "sample": "one frame, oracle Result() restatement (snprintf %g), 1 thread, no file I/O"},
            "checked_against_oracle": bool(text == wtext and wn == npts),
        }
        if emit:
            print(json.dumps(line), flush=True)
    rec.close()
    return line if rank == 0 else None


def run_ingest(args, emit=True):
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
    line = None
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
            if emit:
                print(json.dumps(line), flush=True)
        rec.close()
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return line if rank == 0 else None


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


def time_e2e(call, steps, dev):
    """`steps` blocking host calls between barriers; returns the slowest rank's seconds."""
    import torch
    call()
    call()
    D.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        call()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    D.barrier()
    return D.max_over_ranks(dt, dev)


def e2e_entry(value, h2d, d2h, E, steps, seconds, api, world, ceiling):
    """One end-to-end entry; `roofline` puts the bytes it moved against the probed host-link ceiling."""
    out = {"value": value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "frame_sets_per_step_per_gpu": E, "steps": steps, "ms_per_step": 1e3 * seconds / steps, "api": api}
    up = world * h2d * steps / seconds / 1e9
    down = world * d2h * steps / seconds / 1e9
    out["link_gbs"] = {"h2d": up, "d2h": down, "total": up + down}
    if ceiling:
        # the path is bound by whichever direction is closer to its own ceiling, or by the two together
        fr = {k: v for k, v in (("h2d", up / ceiling["h2d_gbs"] if ceiling.get("h2d_gbs") else None),
                                ("d2h", down / ceiling["d2h_gbs"] if ceiling.get("d2h_gbs") else None),
                                ("bidir", (up + down) / ceiling["bidir_gbs"] if ceiling.get("bidir_gbs") else None))
              if v is not None}
        bound = max(fr, key=fr.get)
        out["roofline"] = {"bound": f"host link ({bound})", "link_peak_gbs": ceiling.get(f"{bound}_gbs"),
                           "achieved_gbs": {"h2d": up, "d2h": down, "bidir": up + down}[bound], "frac": fr[bound],
                           "fractions": fr, "source": "profiles/hostlink_ceiling.json (profiles/hostlink_probe.py)"}
    return out


def run_sequence(args, cfg):
    """--config config4: BASELINE configs[3] as written -- ONE dynamic sequence of 4096 frame sets at
    1920x1200, strong-scaled: rank r owns the contiguous shard slc_shard_range(4096, r, N).  Timed: the
    whole sequence, device-resident (launches over a resident ring of distinct frame sets: 4096 x 89.9 MB
    does not fit one GPU) and streamed from / to pinned host rings through the public host call."""
    import torch
    from structured_light_calculation_b200 import capi

    rank, local_rank, world = D.env_rank_world()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        D.init_process_group("nccl")
    total = args.sequence
    lo, hi = capi.shard_range(total, rank, world)
    mine = hi - lo
    cal, scene, stacks = build_inputs(cfg, args.pool)
    rec = capi.Reconstructor(cfg, device=local_rank, max_batch=args.e2e_chunk, num_slots=args.e2e_slots)
    rec.set_calibration(cal)
    ring = min(args.batch, max(mine, 1))
    d_in = torch.empty((ring, cfg.planes, cfg.height, cfg.width), dtype=torch.uint8, device=dev)
    d_xyzw = torch.empty((ring, cfg.height, cfg.width, 4), dtype=torch.float32, device=dev)
    d_mask = torch.empty((ring, cfg.height, cfg.width), dtype=torch.uint8, device=dev)
    pool_dev = [torch.from_numpy(st).to(dev) for st in stacks]
    for i in range(ring):
        d_in[i].copy_(pool_dev[(lo + i) % len(pool_dev)])
    del pool_dev
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)

    def sequence():
        for done in range(0, mine, ring):
            rec.reconstruct_device(d_in.data_ptr(), min(ring, mine - done), d_xyzw.data_ptr(), d_mask.data_ptr(), None,
                                   stream.cuda_stream)

    for _ in range(args.warmup):
        sequence()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = rec.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    D.barrier()
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(args.steps):
        sequence()
    e1.record(stream)
    torch.cuda.synchronize()
    D.barrier()
    ms = D.max_over_ranks(e0.elapsed_time(e1), dev) / args.steps          # one whole sequence, slowest rank
    launches = int(D.sum_over_ranks(rec.launch_count() - l0, dev))
    clocks = sampler.stop() if rank == 0 else None

    # streamed: the shard through slc_reconstruct_host, E frame sets per call from a pinned host ring
    E = args.e2e_stacks
    h_in = capi.PinnedArray((E, cfg.planes, cfg.height, cfg.width), np.uint8)
    h_xyzw = capi.PinnedArray((E, cfg.height, cfg.width, 4), np.float32)
    h_mask = capi.PinnedArray((E, cfg.height, cfg.width), np.uint8)
    for i in range(E):
        h_in.array[i] = stacks[i % len(stacks)]

    def streamed():
        for done in range(0, mine, E):
            rec.reconstruct_into(h_in, min(E, mine - done), h_xyzw, h_mask)

    rec.reconstruct_into(h_in, min(E, max(mine, 1)), h_xyzw, h_mask)
    D.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    streamed()
    torch.cuda.synchronize()
    s_stream = D.max_over_ranks(time.perf_counter() - t0, dev)
    D.barrier()

    # the same sequence through ONE process: rank 0 drives every GPU with slc_pool (frame sets handed out on
    # demand, PE per call from a pinned host ring), full maps and the depth-only result; the other ranks wait
    streamed_pool = None
    if world > 1:
        D.barrier()
        if rank == 0:
            pool = capi.Pool(cfg, list(range(world)), max_batch=args.e2e_chunk, num_slots=args.e2e_slots)
            pool.set_calibration(cal)
            PE = E * world
            p_in = capi.PinnedArray((PE, cfg.planes, cfg.height, cfg.width), np.uint8)
            for i in range(PE):
                p_in.array[i] = stacks[i % len(stacks)]
            streamed_pool = {"frame_sets_per_call": PE}
            for fmt, name in ((capi.SLC_RESULT_XYZW, "full_maps"), (capi.SLC_RESULT_DEPTH, "depth_only")):
                pb, pres = capi.alloc_result(cfg, PE, fmt, pinned=True)
                pool.reconstruct_into_ex(p_in, PE, pres)
                t0 = time.perf_counter()
                done = 0
                while done < total:
                    n = min(PE, total - done)
                    pool.reconstruct_into_ex(p_in, n, pres)
                    done += n
                secs = time.perf_counter() - t0
                streamed_pool[name] = {"seconds": secs, "value": total / secs, "last_call_shares": pool.last_shares()}
                del pb
            pool.close()
            del p_in
        D.barrier()

    if rank == 0:
        from oracle import sl_oracle as O
        ocfg = O.make_config(cfg.width, cfg.height, cfg.projector_width, cfg.gray_digits, cfg.phase_steps,
                             cfg.fov_min, cfg.fov_max, cfg.modulation_min, host_threads())
        want = O.reconstruct(ocfg, O.make_calib(cal.cam, cal.pro, cal.R, cal.T), stacks[lo % len(stacks)])
        tol = 1e-5 * (cfg.fov_max - cfg.fov_min)
        checked = bool(np.array_equal(d_mask[0].cpu().numpy(), want["mask"]) and
                       np.abs(d_xyzw[0, :, :, 2].cpu().numpy() - want["z"]).max() <= tol and
                       np.array_equal(h_mask.array[0], O.reconstruct(ocfg, O.make_calib(cal.cam, cal.pro, cal.R, cal.T),
                                                                     stacks[0])["mask"]))
        peak, peak_kind = hbm_peak()
        alg = cfg.algorithmic_bytes_per_pixel * cfg.pixels * total
        bound_ms = alg / world / (peak * 1e9) * 1e3
        per_gpu = -(-total // world)
        ceiling = link_ceiling(world)
        h2d, d2h = per_gpu * cfg.stack_bytes, per_gpu * cfg.pixels * 17
        line = {
            "metric": METRIC, "value": total / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"one dynamic sequence of {total} frame sets: " + workload_name(cfg),
                       "frame_sets_total": total, "frame_sets_per_gpu": per_gpu,
                       "l2_policy": f"device-resident ring of {ring} distinct-address frame sets per GPU "
                                    f"({ring * cfg.stack_bytes / 1e9:.1f} GB read + {ring * cfg.pixels * 17 / 1e9:.1f} GB "
                                    f"written per launch, far beyond L2); {total} x {cfg.algorithmic_bytes_per_pixel * cfg.pixels / 1e6:.1f} MB "
                                    f"does not fit one GPU",
                       "parallelism": f"contiguous shards (slc_shard_range) over {world} GPU(s), no collective"},
            "mpix_per_s": total / (ms * 1e-3) * cfg.pixels / 1e6,
            "sequence": {"frame_sets": total, "device_resident_ms": ms, "hbm_bound_ms": bound_ms,
                         "frac_of_bound": bound_ms / ms, "streamed_s": s_stream, "streamed_value": total / s_stream,
                         "streamed_pool": streamed_pool},
            "e2e": e2e_entry(total / s_stream, h2d, d2h, per_gpu, 1, s_stream,
                             f"capi.Reconstructor.reconstruct_into -> slc_reconstruct_host, the rank's whole shard, pinned host "
                             f"rings of {E} frame sets, {args.e2e_slots} stream slots x {args.e2e_chunk} frame sets", world, ceiling),
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": alg / world / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": bound_ms / ms, "traffic": None, "peak_kind": f"of {peak_kind}",
                         "kernel": "slc::reconstruct_vec_kernel", "algorithmic_bytes_per_sequence": alg},
            "cpu_baseline": None, "clocks": clocks, "checked_against_oracle": checked,
        }
        print(json.dumps(line), flush=True)
    rec.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return 0


def next_rows(args):
    """The SURVEY 8(f) rows in the default line: each path's own bench at a bounded size (timed regions
    well under 2 s), reduced to {value, unit, roofline_frac, e2e, checked}."""
    import copy
    out = {}
    small = copy.copy(args)
    small.steps, small.warmup = 3, 3
    small.batch = 64                    # dynamic: one sequence of 100 frames per step
    small.pc_frames, small.ingest_reps, small.dyna_e2e_frames = 4, 8, 12
    for name, fn in (("dynamic", run_dynamic), ("pointcloud", run_pointcloud), ("ingest", run_ingest)):
        try:
            t0 = time.perf_counter()
            ln = fn(small, emit=False)
            out[name] = {"metric": ln["metric"], "value": ln["value"], "unit": ln["unit"],
                         "roofline_frac": ln["roofline"]["frac"], "kernel": ln["roofline"]["kernel"],
                         "e2e_value": ln["e2e"]["value"], "checked": ln["checked_against_oracle"],
                         "workload": ln["config"]["workload"], "wall_s": time.perf_counter() - t0}
            if name == "pointcloud":
                out[name]["binary_cloud"] = ln.get("binary_cloud")
            if name == "dynamic":
                out[name]["e2e_depth_value"] = ln["e2e_compact"]["depth"]["value"]
        except Exception as e:      # a next row must never take the headline line down
            out[name] = {"error": repr(e)}
    return out


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
    ap.add_argument("--e2e-chunk", type=int, default=1, help="frame sets per upload/launch/download chunk "
                    "(1: the shortest pipeline ramp; measured 982 / 939 / 900 frame sets/s with 1 / 2 / 4 on one box)")
    ap.add_argument("--e2e-slots", type=int, default=4)
    ap.add_argument("--e2e-mem", default="pinned", choices=["pinned", "wc", "huge"],
                    help="host memory of the end-to-end input ring: cudaHostAlloc, write-combined, 2 MB pages")
    ap.add_argument("--pxt", type=int, default=0, help="tuning: pixels per thread (4/8/16)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-numa-bind", action="store_true")
    ap.add_argument("--no-next-rows", action="store_true")
    ap.add_argument("--sustained-seconds", type=float, default=2.2)
    ap.add_argument("--sequence", type=int, default=4096, help="--config config4: frame sets of the sequence")
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
        run_dynamic(args)
        return 0
    if args.path == "pointcloud":
        run_pointcloud(args)
        return 0
    if args.path == "ingest":
        run_ingest(args)
        return 0
    if args.path == "app":
        return run_app(args)
    if args.config == "config4":
        return run_sequence(args, cfg)

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

    F = args.batch
    cal, scene, stacks = build_inputs(cfg, args.pool)
    rec = capi.Reconstructor(cfg, device=local_rank, max_batch=args.e2e_chunk, num_slots=args.e2e_slots)
    rec.set_calibration(cal)
    if args.pxt:
        rec.set_pixels_per_thread(args.pxt)
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
    clocks = sampler.stop() if rank == 0 else None

    # ---- the same kernel for >= 2 s back to back: does the burst figure hold at sustained clocks? ----
    sustained = None
    if args.sustained_seconds > 0:
        n_sus = max(args.steps, int(args.sustained_seconds / (statistics.fmean(per_launch_ms) * 1e-3)) + 1)
        sampler2 = ClockSampler(local_rank)
        if rank == 0:
            sampler2.start()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        D.barrier()
        torch.cuda.synchronize()
        s0.record(stream)
        for _ in range(n_sus):
            step()
        s1.record(stream)
        torch.cuda.synchronize()
        D.barrier()
        sus_ms = D.max_over_ranks(s0.elapsed_time(s1), dev)
        clocks2 = sampler2.stop() if rank == 0 else None
        peak_s, _ = hbm_peak()
        sus_gbs = cfg.algorithmic_bytes_per_pixel * cfg.pixels * F * n_sus / (sus_ms * 1e-3) / 1e9
        sustained = {"seconds": sus_ms * 1e-3, "launches_per_gpu": n_sus, "value": world * F * n_sus / (sus_ms * 1e-3),
                     "unit": UNIT, "achieved_gbs_per_gpu": sus_gbs, "frac": sus_gbs / peak_s, "clocks": clocks2}

    # ---- the DEPTH layout of the same kernel (z + bit mask: 4.125 instead of 17 B/px written) ----
    d_depth = d_xyzw.view(-1)[: F * cfg.pixels].view(F, cfg.height, cfg.width)
    d_bits = d_mask.view(-1)[: (F * capi.bits_bytes(cfg.pixels) + 3) // 4 * 4]
    res_depth = capi.make_result(capi.SLC_RESULT_DEPTH, depth=d_depth.data_ptr(), mask_bits=d_bits.data_ptr())
    for _ in range(3):
        rec.reconstruct_device_ex(d_in.data_ptr(), F, res_depth, stream.cuda_stream)
    q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    q0.record(stream)
    for _ in range(args.steps):
        rec.reconstruct_device_ex(d_in.data_ptr(), F, res_depth, stream.cuda_stream)
    q1.record(stream)
    torch.cuda.synchronize()
    depth_ms = D.max_over_ranks(q0.elapsed_time(q1), dev) / args.steps
    depth_bytes = (cfg.planes + 4.125) * cfg.pixels * F
    depth_z = d_depth[0].cpu().numpy()
    depth_bits = d_bits[: capi.bits_bytes(cfg.pixels)].cpu().numpy()
    step()                                # the spot check below reads the xyzw + mask layout again
    torch.cuda.synchronize()

    # ---- end to end through the public host call, pinned host buffers ----
    E = args.e2e_stacks
    memflag = {"pinned": 0, "wc": capi.SLC_HOST_WRITE_COMBINED, "huge": capi.SLC_HOST_HUGE_PAGES}[args.e2e_mem]
    h_in = capi.PinnedArray((E, cfg.planes, cfg.height, cfg.width), np.uint8, memflag)
    outflag = capi.SLC_HOST_HUGE_PAGES if args.e2e_mem == "huge" else 0
    h_xyzw = capi.PinnedArray((E, cfg.height, cfg.width, 4), np.float32, outflag)
    h_mask = capi.PinnedArray((E, cfg.height, cfg.width), np.uint8, outflag)
    for i in range(E):
        h_in.array[i] = stacks[i % len(stacks)]
    e2e_steps = args.steps
    ceiling = link_ceiling(world)
    slots_note = f"pinned host buffers ({args.e2e_mem}), {args.e2e_slots} stream slots x {args.e2e_chunk} frame sets"
    e2e_s = time_e2e(lambda: rec.reconstruct_into(h_in, E, h_xyzw, h_mask), e2e_steps, dev)
    e2e = e2e_entry(world * E * e2e_steps / e2e_s, E * cfg.stack_bytes, E * cfg.pixels * 17, E, e2e_steps, e2e_s,
                    "capi.Reconstructor.reconstruct_into -> slc_reconstruct_host, " + slots_note, world, ceiling)

    # ---- the same with the reduced result formats (fewer bytes back over the link) ----
    # (every rank runs the same calls, so a failure here is the same on every rank and the barriers stay matched;
    #  it is recorded, never allowed to take the headline down)
    e2e_compact = {}
    bufs_d = bufs_p = None
    try:
        bufs_d, res_d = capi.alloc_result(cfg, E, capi.SLC_RESULT_DEPTH, pinned=True)
        s_d = time_e2e(lambda: rec.reconstruct_into_ex(h_in, E, res_d), e2e_steps, dev)
        e2e_compact["depth"] = e2e_entry(world * E * e2e_steps / s_d, E * cfg.stack_bytes,
                                         E * (cfg.pixels * 4 + capi.bits_bytes(cfg.pixels)), E, e2e_steps, s_d,
                                         "slc_reconstruct_host_ex SLC_RESULT_DEPTH (z + bit mask), " + slots_note, world, ceiling)
    except capi.SlcError as e:
        e2e_compact["depth"] = {"error": str(e)}
        bufs_d = None
    try:
        bufs_p, res_p = capi.alloc_result(cfg, E, capi.SLC_RESULT_POINTS, capi.SLC_ORDER_REFERENCE, pinned=True)
        s_p = time_e2e(lambda: rec.reconstruct_into_ex(h_in, E, res_p), e2e_steps, dev)
        n_pts = int(bufs_p["n_points"].array.sum())
        e2e_compact["points"] = e2e_entry(world * E * e2e_steps / s_p, E * cfg.stack_bytes,
                                          12 * n_pts + E * capi.bits_bytes(cfg.pixels) + 8 * E, E, e2e_steps, s_p,
                                          "slc_reconstruct_host_ex SLC_RESULT_POINTS (float3 of the valid pixels in Result()'s "
                                          "order + bit mask), " + slots_note, world, ceiling)
        e2e_compact["points"]["valid_fraction"] = n_pts / (E * cfg.pixels)
    except capi.SlcError as e:
        e2e_compact["points"] = {"error": str(e)}
        bufs_p = None

    # ---- one process, a feeder thread per GPU (slc_pool): rank 0 drives every GPU, the other ranks wait.
    #      Frame sets are handed out on demand, so GPUs behind a faster host link take more of them. ----
    e2e_pool = None
    if world > 1:
        D.barrier()
        if rank == 0:
            pool = capi.Pool(cfg, list(range(world)), max_batch=args.e2e_chunk, num_slots=args.e2e_slots)
            pool.set_calibration(cal)
            PE = E * world
            p_in = capi.PinnedArray((PE, cfg.planes, cfg.height, cfg.width), np.uint8, memflag)
            for i in range(PE):
                p_in.array[i] = stacks[i % len(stacks)]

            def pool_run(fmt, d2h_per_set, what):
                pb, pres = capi.alloc_result(cfg, PE, fmt, pinned=True)
                for _ in range(2):
                    pool.reconstruct_into_ex(p_in, PE, pres)
                t0 = time.perf_counter()
                for _ in range(e2e_steps):
                    pool.reconstruct_into_ex(p_in, PE, pres)
                ps = time.perf_counter() - t0
                ent = e2e_entry(PE * e2e_steps / ps, E * cfg.stack_bytes, E * d2h_per_set, E, e2e_steps, ps,
                                f"capi.Pool.reconstruct_into_ex -> slc_pool_reconstruct_host ({what}): ONE process, {world} feeder "
                                f"threads, frame sets handed out on demand (the other ranks idle at a barrier), " + slots_note,
                                world, ceiling)
                ent["frame_sets_taken_per_gpu_last_call"] = pool.last_shares()
                return pb, ent

            pb, e2e_pool = pool_run(capi.SLC_RESULT_XYZW, cfg.pixels * 17, "xyzw + mask")
            e2e_pool["equals_per_rank_result"] = bool(np.array_equal(pb["mask"].array[E], h_mask.array[0]) and
                                                      np.array_equal(pb["xyzw"].array[PE - E], h_xyzw.array[0]))
            del pb
            pbd, ent = pool_run(capi.SLC_RESULT_DEPTH, cfg.pixels * 4 + capi.bits_bytes(cfg.pixels), "depth + bit mask")
            ent["equals_per_rank_result"] = bool(bufs_d is not None and
                                                 np.array_equal(pbd["depth"].array[PE - E], bufs_d["depth"].array[0]))
            e2e_pool["depth"] = ent
            pool.close()
            del p_in, pbd
        D.barrier()

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
        # the reduced formats are selections of that output, bit for bit
        sel = np.transpose(h_xyzw.array[0, ..., :3], (1, 0, 2))[h_mask.array[0].T.astype(bool)]
        formats_ok = bool(bufs_d is not None and bufs_p is not None and np.array_equal(depth_z, z_dev) and
                          np.array_equal(np.unpackbits(depth_bits, bitorder="little")[: cfg.pixels], m_dev.reshape(-1)) and
                          np.array_equal(bufs_d["depth"].array[0], h_xyzw.array[0, :, :, 2]) and
                          int(bufs_p["n_points"].array[0]) == len(sel) and
                          np.array_equal(bufs_p["points"].array[0, : len(sel)], sel))
        e2e_compact["checked_bit_equal_to_full_map"] = formats_ok
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_baseline(cfg, cal, stacks[0])
    del d_in, d_xyzw, d_mask, d_depth, d_bits
    torch.cuda.empty_cache()
    rows = None
    if world == 1 and not args.no_next_rows and args.config == "config2":
        rec.close()
        rows = next_rows(args)

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
            "config": bench_config(cfg, args, world),
            "mpix_per_s": value * cfg.pixels / 1e6,
            "e2e": e2e,
            "e2e_compact": e2e_compact,
            "e2e_pool": e2e_pool,
            "gpu_launches": launches_all,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": (traffic_ps * F if traffic_ps else None),
                         "traffic_source": "ncu launch list of this command at 256 frame sets per launch "
                                           "(profiles/traffic.json), scaled to the batch",
                         "peak_kind": f"of {peak_kind} copy bandwidth (a read+write copy; a write-only stream measures "
                                      f"higher on this part, so frac can exceed 1)",
                         "kernel": "slc::reconstruct_vec_kernel",
                         "algorithmic_bytes_per_launch": alg_bytes, "mean_launch_ms": mean_launch_ms,
                         "min_launch_ms": min(per_launch_ms), "max_launch_ms": max(per_launch_ms)},
            "sustained": sustained,
            "depth_layout": {"value": world * F / (depth_ms * 1e-3), "unit": UNIT, "ms_per_launch": depth_ms,
                             "algorithmic_bytes_per_launch": depth_bytes,
                             "achieved_gbs": depth_bytes / (depth_ms * 1e-3) / 1e9,
                             "frac": depth_bytes / (depth_ms * 1e-3) / 1e9 / peak,
                             "api": "slc_reconstruct_device_ex SLC_RESULT_DEPTH: the fused kernel writes z + one bit per pixel"},
            "next_rows": rows,
            "cpu_baseline": cpu,
            "clocks": clocks,
            "kernel": {"variant": info.kernel_variant, "regs": info.kernel_regs, "block": info.kernel_block,
                       "smem": info.kernel_smem},
            "checked_against_oracle": checked,
            "numa": numa,
        }
        print(json.dumps(line), flush=True)
    if rows is None:
        rec.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
