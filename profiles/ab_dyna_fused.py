#!/usr/bin/env python
"""A/B of dyna_fused_kernel launch shapes in ONE process (one set of inputs, one GPU):
for every variant name, set the variant's environment switch, check 12 frames against the CPU oracle
(mask and f32(U) bit-exact, z and deltaZ in tolerance) and time whole sequences with CUDA events.  Per-kernel times come from
the ncu launch list of the same command (profiles/r01_dyna_ab_launches.csv).

The variants lived in a temporary patch of launch_dyna_fused() that read SLC_DYNA_FUSED
(o = the shipped kernel, a-e / p2-p6 = the shapes and prefetch schemes DESIGN.md 9.1 lists); none
was faster, so the patch was dropped and the shipped library ignores the variable: run against it,
every name times the shipped kernel.  The same holds for the later patches behind "pipe:C:P:K"
(SLC_DYNA_PIPE: frame-chunked strip / fused pipeline on two streams) and "lean:3" / "lean:4"
(SLC_DYNA_LEAN: row constants recomputed per frame, 3 or 4 blocks per SM).  Kept as the harness for the
next attempt.

    python profiles/ab_dyna_fused.py [comma-separated variants, default "a,b,c,d,e"] [frames, default 100]
"""
import os
import sys
import json

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from structured_light_calculation_b200 import capi, synth  # noqa: E402
from structured_light_calculation_b200.calibration import load_calibration  # noqa: E402
from structured_light_calculation_b200.configs import CONFIGS  # noqa: E402
from oracle import sl_oracle as O  # noqa: E402   (checker only)


def main():
    variants = (sys.argv[1] if len(sys.argv) > 1 else "a,b,c,d,e").split(",")
    F = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    window, S, steps = 21, 4, 5
    dev = torch.device("cuda", 0)
    cfg = CONFIGS["reference_default"]
    base = load_calibration(os.path.join(ROOT, "tests", "golden", "Result.yml"))
    cal = synth.synthetic_calibration(cfg, base)
    scene = synth.make_scene(cfg, cal)
    stack = synth.render_stack(cfg, scene, noise_sigma=1.0, seed=77)
    pool = synth.render_dyna_frames(cfg, cal, 8, stripe_period=20.0, z_step=0.3, noise_sigma=1.5)
    order = [k if k < 8 else 14 - k for k in (f % 14 for f in range(F))]
    frames = np.stack([pool[k] for k in order])
    rec = capi.Reconstructor(cfg, device=0, max_batch=1, num_slots=1)
    rec.set_calibration(cal)
    u0 = rec.reconstruct(stack, parity=True)["proj_u"][0]
    ocfg = O.make_config(cfg.width, cfg.height, cfg.projector_width, cfg.gray_digits, cfg.phase_steps)
    ocal = O.make_calib(cal.cam, cal.pro, cal.R, cal.T)
    first = O.reconstruct(ocfg, ocal, stack)
    NCHK = 12
    want = O.dyna_sequence(ocfg, ocal, first["proj_u"], first["z"], frames[:NCHK + 1], window)
    tol = 1e-5 * (cfg.fov_max - cfg.fov_min)

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

    for v in variants:
        # "name" -> SLC_DYNA_FUSED=name (the dropped kernel-shape patch); "pipe:C:P:K" -> SLC_DYNA_PIPE="C,P,K"
        # (chunks of C frames, strip stream priority P, strip blocks per SM capped at K; the pipeline experiment)
        os.environ.pop("SLC_DYNA_PIPE", None)
        os.environ.pop("SLC_DYNA_LEAN", None)
        if v.startswith("lean:"):          # row constants recomputed per frame; 3 or 4 blocks per SM
            os.environ["SLC_DYNA_LEAN"] = v[5:]
        elif v.startswith("pipe:"):
            os.environ["SLC_DYNA_PIPE"] = v[5:].replace(":", ",")
        else:
            os.environ["SLC_DYNA_FUSED"] = v
        d_xyzw.zero_(); d_mask.zero_(); d_dz.zero_()
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        xyzw = d_xyzw[0, :NCHK].cpu().numpy(); mask = d_mask[0, :NCHK].cpu().numpy(); dz = d_dz[0, :NCHK].cpu().numpy()
        ok = all(np.array_equal(mask[f], want[f]["mask"]) and
                 np.array_equal(xyzw[f, ..., 3], want[f]["proj_u"].astype(np.float32)) and
                 np.abs(xyzw[f, ..., 2] - want[f]["z"]).max() <= tol for f in range(NCHK))
        zprev = first["z"]
        dz_err = 0.0
        for f in range(NCHK):
            dz_err = max(dz_err, float(np.abs(dz[f] - (want[f]["z"] - zprev)).max()))
            zprev = want[f]["z"]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(steps):
            step()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / (steps * S)
        print(json.dumps({"variant": v, "frames": F, "us_per_sequence": 1e3 * ms,
                          "frames_per_s": (F - 1) / (ms * 1e-3), "parity_ok": bool(ok), "dz_max_err": dz_err}), flush=True)
    rec.close()


if __name__ == "__main__":
    main()
