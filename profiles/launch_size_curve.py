"""Roofline fraction of the fused first-frame kernel against the size of a launch, BASELINE configs[1].

    python profiles/launch_size_curve.py [--out gpurun_out/r02_launch_size_curve.txt]

For n = 1, 2, 4, 8, 16, 64, 256 frame sets per launch:
  * back to back: 256/n launches in a row over 256 DISTINCT resident frame sets (13 GB in, 10 GB out), so
    no input is served from L2 and the writes a launch leaves dirty in L2 are paid for by the next one
    (only the last launch of the row keeps that advantage: 1/(256/n) of the total) -- the sustained
    throughput of a stream of n-set calls;
  * isolated: one launch between CUDA events after a 512 MB write has flushed L2, median of 15 -- the
    latency of a single call (its tail of dirty lines still sits in L2 when the event fires, so this
    figure flatters small launches; it is reported as a latency, not as a roofline fraction).
"""
import argparse
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from structured_light_calculation_b200 import capi, synth  # noqa: E402
from structured_light_calculation_b200.calibration import load_calibration  # noqa: E402
from structured_light_calculation_b200.configs import CONFIGS  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r02_launch_size_curve.txt"))
    ap.add_argument("--total", type=int, default=256)
    ap.add_argument("--format", default="xyzw", choices=["xyzw", "depth"])
    ap.add_argument("--pxt", type=int, default=0)
    args = ap.parse_args()
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    cfg = CONFIGS["config2"]
    base = load_calibration(os.path.join(ROOT, "tests", "golden", "Result.yml"))
    cal = synth.synthetic_calibration(cfg, base)
    scene = synth.make_scene(cfg, cal)
    dev = torch.device("cuda", 0)
    pool = [torch.from_numpy(synth.render_stack(cfg, scene, noise_sigma=1.0, seed=1234 + i)).to(dev) for i in range(4)]
    T = args.total
    d_in = torch.empty((T, cfg.planes, cfg.height, cfg.width), dtype=torch.uint8, device=dev)
    for i in range(T):
        d_in[i].copy_(pool[i % 4])
    d_xyzw = torch.empty((T, cfg.height, cfg.width, 4), dtype=torch.float32, device=dev)
    d_mask = torch.empty((T, cfg.height, cfg.width), dtype=torch.uint8, device=dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    rec = capi.Reconstructor(cfg, device=0, max_batch=1, num_slots=1)
    rec.set_calibration(cal)
    if args.pxt:
        rec.set_pixels_per_thread(args.pxt)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    bpp = cfg.algorithmic_bytes_per_pixel if args.format == "xyzw" else cfg.planes + 4.125
    npx, bb = cfg.pixels, capi.bits_bytes(cfg.pixels)

    def launch(first, n):
        if args.format == "xyzw":
            rec.reconstruct_device(d_in[first].data_ptr(), n, d_xyzw[first].data_ptr(), d_mask[first].data_ptr(), None,
                                   stream.cuda_stream)
        else:
            res = capi.make_result(capi.SLC_RESULT_DEPTH, depth=d_xyzw.data_ptr() + 4 * npx * first,
                                   mask_bits=d_mask.data_ptr() + bb * first)
            rec.reconstruct_device_ex(d_in[first].data_ptr(), n, res, stream.cuda_stream)

    lines = [f"# {cfg.width}x{cfg.height} G{cfg.gray_digits} N{cfg.phase_steps}, {bpp} B/px algorithmic ({args.format} layout), "
             f"HBM copy peak {peak:.0f} GB/s, kernel variant {rec.info().kernel_variant}",
             f"{'sets/launch':>11s} {'launches':>8s} {'us/launch b2b':>14s} {'sets/s b2b':>11s} {'GB/s b2b':>9s} {'frac b2b':>8s} "
             f"{'us isolated':>11s} {'frac isolated':>13s}"]
    for n in (1, 2, 4, 8, 16, 64, 256):
        k = T // n
        for _ in range(2):
            for j in range(k):
                launch(j * n, n)
        torch.cuda.synchronize()
        reps = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for j in range(k):
                launch(j * n, n)
            e1.record(stream)
            torch.cuda.synchronize()
            reps.append(e0.elapsed_time(e1) * 1e3 / k)
        b2b = statistics.median(reps)
        iso = []
        for r in range(15):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            launch(((r * n) % (T - n + 1)), n)
            e1.record(stream)
            torch.cuda.synchronize()
            iso.append(e0.elapsed_time(e1) * 1e3)
        iso_us = statistics.median(iso)
        gbs = bpp * npx * n / (b2b * 1e-6) / 1e9
        lines.append(f"{n:11d} {k:8d} {b2b:14.2f} {n / (b2b * 1e-6):11.0f} {gbs:9.0f} {gbs / peak:8.3f} "
                     f"{iso_us:11.2f} {bpp * npx * n / (iso_us * 1e-6) / 1e9 / peak:13.3f}")
    rec.close()
    text = "\n".join(lines)
    print(text)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        f.write(text + "\n")


if __name__ == "__main__":
    main()
