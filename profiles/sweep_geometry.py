"""Device-resident throughput of the fused first-frame kernel over (W, H, G, N): frame sets/s and the
fraction of the measured HBM copy peak (algorithmic bytes 2G+N+17 per pixel).  Run on a B200:
    python profiles/sweep_geometry.py"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from structured_light_calculation_b200 import capi, synth  # noqa: E402
from structured_light_calculation_b200.calibration import load_calibration  # noqa: E402
from structured_light_calculation_b200.configs import StackConfig  # noqa: E402

peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
base = load_calibration(os.path.join(ROOT, "tests", "golden", "Result.yml"))
dev = torch.device("cuda", 0)
cases = [(1280, 1024, 1280, 6, 4), (1280, 1024, 1280, 7, 4), (1280, 1024, 1280, 8, 4), (1280, 1024, 2560, 9, 4),
         (1920, 1200, 2560, 6, 4), (1920, 1200, 2560, 7, 4), (1920, 1200, 2560, 8, 4), (1920, 1200, 2560, 9, 4),
         (1280, 1040, 1280, 7, 4), (1296, 1024, 1280, 7, 4),
         (1280, 1024, 1280, 8, 3), (1280, 1024, 1280, 7, 6), (1280, 1024, 1280, 8, 5),     # <G, 0> instances
         (1280, 1024, 4096, 11, 4), (1280, 1024, 1280, 4, 4), (1280, 1024, 4096, 12, 3),
         (1280, 1024, 1280, 7, 3), (1920, 1200, 2560, 9, 3),                              # <G, 3> instances
         (1920, 1200, 2560, 9, 5), (1920, 1200, 2560, 9, 6), (1920, 1200, 2560, 9, 8), (1920, 1200, 2560, 9, 12)]
if os.environ.get("SWEEP_QUICK"):          # one small batch per Gray depth, for an ncu capture
    cases = cases[:4]
for W, H, PW, G, N in cases:
    cfg = StackConfig(W, H, PW, G, N)
    cal = synth.synthetic_calibration(cfg, base)
    rec = capi.Reconstructor(cfg, device=0, max_batch=1, num_slots=1)
    rec.set_calibration(cal)
    if os.environ.get("SWEEP_PXT"):                 # pixels per thread of the vector kernel: 4 / 8 / 16
        rec.set_pixels_per_thread(int(os.environ["SWEEP_PXT"]))
    F = 16 if os.environ.get("SWEEP_QUICK") else max(8, int(6e9 // (cfg.planes * cfg.pixels)))
    scene = synth.make_scene(cfg, cal)                   # rendered stack (random bytes would make most pixels invalid)
    stack = torch.from_numpy(synth.render_stack(cfg, scene, noise_sigma=1.0, seed=5)).to(dev)
    d_in = stack.unsqueeze(0).expand(F, -1, -1, -1).contiguous()
    d_xyzw = torch.empty((F, H, W, 4), dtype=torch.float32, device=dev)
    d_mask = torch.empty((F, H, W), dtype=torch.uint8, device=dev)
    rec.time_device(d_in.data_ptr(), F, d_xyzw.data_ptr(), d_mask.data_ptr(), 1 if os.environ.get("SWEEP_QUICK") else 3)
    # the fastest of three short bursts: a sweep of 16 cases runs into the power cap, and which case meets the
    # dip differs from run to run (round 1: 1280x1040 at 0.854; the first round-2 run: the 1920x1200 rows)
    runs = [rec.time_device(d_in.data_ptr(), F, d_xyzw.data_ptr(), d_mask.data_ptr(), 1 if os.environ.get("SWEEP_QUICK") else 6)
            for _ in range(1 if os.environ.get("SWEEP_QUICK") else 3)]
    ms = min(runs)
    gbs = (cfg.planes + 17) * cfg.pixels * F / (ms * 1e-3) / 1e9
    print(f"{W}x{H} G={G} N={N} P={cfg.planes}: {F / (ms * 1e-3):9.0f} frame sets/s  {gbs:7.0f} GB/s  {gbs / peak:.3f} of peak  "
          f"(batch {F}, best of {len(runs)} bursts {ms:.3f} ms, slowest {max(runs):.3f} ms, variant {rec.info().kernel_variant})")
    time.sleep(0.5)
    rec.close()
    del d_in, d_xyzw, d_mask
    torch.cuda.empty_cache()
