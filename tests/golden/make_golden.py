#!/usr/bin/env python
"""Generate the committed golden fixtures that pin the CPU oracle.

Run in the build container (needs cv2 4.13 and, for the Gray-code table check,
/root/reference):  python tests/golden/make_golden.py

The reference ships no tests or golden vectors, so the pins are made here (the
reference's own sources are compiled in place against a minimal OpenCV stand-in,
oracle/Makefile -> oracle/_ref, for item 7; cv2 4.13 pins the OpenCV primitives):
  1. fast_atan2.npz    -- the container's *scalar* cv2.fastAtan2 on all 511^2
                          (s, c) pairs reachable from 4-step u8 input: sha256 of
                          the full f32 table + an explicit 8k-sample subset;
  2. gray_code_g6.npz  -- the gray2bin table as CDecodeGray::Decode builds it from
                          the reference's own Patterns/vGrayCode.txt;
  3. pipeline_*.npz    -- a whole small stack pushed through an INDEPENDENT
                          numpy + cv2 restatement of the reference loops
                          (cv2.subtract for the saturating difference,
                          cv2.fastAtan2 for the arctangent, cv2.gemm for P),
                          with every intermediate plane stored;
  4. dyna_g6.npz       -- the dynamic-frame path (StripRegression, FillOtherDeltaProU) restated
                          with numpy cumsum/argmin and cv2.blur;
  5. kat.npz           -- the hand-derived known-answer tables of SURVEY.md 8(c);
  7. reference_run_g6n4.npz -- inputs and outputs of the REFERENCE ITSELF: oracle/_ref/dynaframe_ref
                          (the reference's own path sources compiled in place, oracle/Makefile) run on a
                          small synthetic stack + dynamic sequence laid out as the reference reads it;
                          every plane it computes and the text clouds it writes.
  6. bmp_cases.npz     -- BMP files (8-bit gray / colour palette, 24- and 32-bit, bottom-up and top-down,
                          padded rows) with what cv2.imread(path, IMREAD_GRAYSCALE) -- the call of
                          CSensorV.cpp:111-114 -- returns for each.
Nothing here imports the oracle: the fixtures are independent of it.
"""
import hashlib
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from structured_light_calculation_b200 import synth  # noqa: E402
from structured_light_calculation_b200.calibration import load_calibration  # noqa: E402
from structured_light_calculation_b200.configs import StackConfig  # noqa: E402

REF = "/root/reference/DynaFrame/DynaFrame"


def golden_fast_atan2():
    v = (np.arange(-255, 256) / 2).astype(np.float32)
    s, c = np.meshgrid(v, v, indexing="ij")
    table = np.array([cv2.fastAtan2(float(a), float(b)) for a, b in zip(s.ravel(), c.ravel())], dtype=np.float32)
    digest = hashlib.sha256(table.tobytes()).hexdigest()
    rng = np.random.Generator(np.random.PCG64(7))
    idx = np.sort(rng.choice(table.size, 8192, replace=False))
    # a few non-half-integer inputs too (N-step sums are arbitrary floats)
    ys = rng.uniform(-400, 400, 4096).astype(np.float32)
    xs = rng.uniform(-400, 400, 4096).astype(np.float32)
    ys[:8] = [0, 0, 1, -1, 0, 3.5, -3.5, 1e-20]
    xs[:8] = [0, 1, 0, 0, -1, 3.5, 3.5, 1e-20]
    free = np.array([cv2.fastAtan2(float(a), float(b)) for a, b in zip(ys, xs)], dtype=np.float32)
    np.savez_compressed(os.path.join(HERE, "fast_atan2.npz"), sha256=np.array(digest), idx=idx,
                        sample=table[idx], free_y=ys, free_x=xs, free_out=free)
    print("fast_atan2: sha256", digest, "max", table.max())


def golden_gray_code():
    path = os.path.join(REF, "Patterns", "vGrayCode.txt")
    rows = np.loadtxt(path, dtype=np.int64)
    lut = np.zeros(64, dtype=np.int16)
    for b, g in rows:   # CDecodeGray.cpp:120-125
        lut[g] = b
    np.savez_compressed(os.path.join(HERE, "gray_code_g6.npz"), rows=rows.astype(np.int16), lut=lut)
    print("gray code rows", rows.shape)


def numpy_cv2_pipeline(cfg: StackConfig, cal, planes: np.ndarray) -> dict:
    """The reference loops restated with numpy + cv2 primitives (N = 4 only)."""
    assert cfg.phase_steps == 4
    H, W, G = cfg.height, cfg.width, cfg.gray_digits
    gp = cfg.projector_width // (1 << G)
    T = cfg.projector_width // (1 << (G - 1))
    # a3/a4 CDecodeGray.cpp:155-202
    code = np.zeros((H, W), dtype=np.int64)
    for b in range(G):
        diff = cv2.subtract(planes[2 * b], planes[2 * b + 1])   # saturating u8
        binp = np.where(diff > 0, 255, 0).astype(np.uint8)
        code += (binp == 255).astype(np.int64) << b
    n = 1 << G
    lut = np.zeros(n, dtype=np.int16)
    for b in range(n):
        lut[b ^ (b >> 1)] = b
    kbin = lut[code]
    gray_val = kbin.astype(np.float64) * float(gp)
    # a6 CDecodePhase.cpp:54-77
    ph = planes[2 * G:].astype(np.float32)
    sinv = (ph[0] - ph[2]) / np.float32(2)
    cosv = (ph[1] - ph[3]) / np.float32(2)
    x = np.array([cv2.fastAtan2(float(a), float(b)) for a, b in zip(sinv.ravel(), cosv.ravel())],
                 dtype=np.float32).reshape(H, W)
    q = x / np.float32(360)                                    # f32 / int -> f32
    pix = (q.astype(np.float64) * float(T)).astype(np.float32)  # * (double)T, narrowed
    pix = (pix.astype(np.float64) + 0.5).astype(np.float32)    # += 0.5 (double literal)
    over = pix > np.float32(T)
    pix = np.where(over, (pix - np.float32(T)).astype(np.float32), pix)
    phase_pix = pix.astype(np.float64)
    # a7 CCalculation.cpp:562-589
    phase = phase_pix.copy()
    even = ((gray_val / gp).astype(np.int64) % 2) == 0
    corr = np.zeros((H, W), dtype=np.int8)
    m = even & (phase_pix > T * 0.75)
    phase[m] = phase_pix[m] - T
    corr[m] = -1
    m = ~even & (phase_pix < T * 0.25)
    phase[m] = phase_pix[m] + T
    corr[m] = 1
    phase[~even] = phase[~even] - 0.5 * T
    U = gray_val + phase
    # a8 CCalculation.cpp:135-166
    RT = np.concatenate([cal.R, cal.T.reshape(3, 1)], axis=1)
    P = cv2.gemm(cal.pro, RT, 1.0, None, 0.0)
    fu, fv, cu, cv_ = cal.cam[0, 0], cal.cam[1, 1], cal.cam[0, 2], cal.cam[1, 2]
    A = fu * fv * P[0, 3]
    B = fu * fv * P[2, 3]
    u = np.arange(W, dtype=np.float64)[None, :]
    v = np.arange(H, dtype=np.float64)[:, None]
    cC = (u - cu) * fv * P[0, 0] + (v - cv_) * fu * P[0, 1] + fu * fv * P[0, 2]
    cD = (u - cu) * fv * P[2, 0] + (v - cv_) * fu * P[2, 1] + fu * fv * P[2, 2]
    # a9 CCalculation.cpp:672-708
    with np.errstate(divide="ignore", invalid="ignore"):
        z = -(A - B * U) / (cC - cD * U)
    has = U != 0
    bad = (z < cfg.fov_min) | (z > cfg.fov_max)
    mask = (has & ~bad).astype(np.uint8)
    z = np.where(mask.astype(bool), z, 0.0)
    # a10 CCalculation.cpp:756-771
    xx = z * (u - cu) / fu
    yy = z * (v - cv_) / fv
    return dict(kbin=kbin.astype(np.int16), gray_val=gray_val, phase_pix=phase_pix, corr=corr, proj_u=U,
                x=xx, y=yy, z=z, mask=mask, P=P, A=np.float64(A), B=np.float64(B), cC=cC, cD=cD)


def golden_pipeline():
    base = load_calibration(os.path.join(HERE, "Result.yml"))
    cases = {
        # the reference's own digit counts at a reduced camera size
        "pipeline_g6n4": (StackConfig(96, 64, 1280, 6, 4), 1.0, 11),
        # config-2 style (8 bits + complementary LSB), ragged width (not a multiple of 16)
        "pipeline_g9n4": (StackConfig(88, 40, 2560, 9, 4), 2.0, 12),
    }
    for name, (cfg, sigma, seed) in cases.items():
        cal = synth.synthetic_calibration(cfg, base)
        scene = synth.make_scene(cfg, cal)
        planes = synth.render_stack(cfg, scene, noise_sigma=sigma, seed=seed)
        out = numpy_cv2_pipeline(cfg, cal, planes)
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            cfg=np.array([cfg.width, cfg.height, cfg.projector_width, cfg.gray_digits, cfg.phase_steps]),
            cam=cal.cam, pro=cal.pro, R=cal.R, T=cal.T, planes=planes, true_U=scene.U, true_z=scene.z, **out)
        print(name, "valid", out["mask"].mean(), "corr", np.unique(out["corr"], return_counts=True))


def numpy_cv2_dyna(frames: np.ndarray, U0: np.ndarray, window: int = 21) -> dict:
    """StripRegression + FillOtherDeltaProU + U accumulation with numpy cumsum / cv2.blur
    (CCalculation.cpp:789-892, 595-663), independent of the oracle."""
    F, H, W = frames.shape
    half = window // 2
    strips = np.zeros((F, H, W, 2), np.float32)
    for f in range(F):
        cs = np.cumsum(np.vstack([np.zeros((1, W), np.int64), frames[f].astype(np.int64)]), axis=0)
        S = np.zeros((H, W), np.float32)
        for h in range(half, H - half):
            S[h, half:W - half] = (cs[h + half + 1] - cs[h - half])[half:W - half]
        for h in range(half, H - half):
            for w in range(half, W - half):
                win = S[h, w - half:w + half]
                M, m = win.max(), win.min()
                strips[f, h, w, 1] = 0 if S[h, w] == M else np.argmax(win) - half
                strips[f, h, w, 0] = 0 if S[h, w] == m else np.argmin(win) - half
    dps, Us = [], []
    U = U0.copy()
    for f in range(1, F):
        dB = strips[f - 1, ..., 0] - strips[f, ..., 0]
        dW = strips[f - 1, ..., 1] - strips[f, ..., 1]
        t = np.where(np.abs(dB) < np.abs(dW), dB, dW).astype(np.float32)
        dP = cv2.blur(t, (3, 3))
        U = U + dP.astype(np.float64)
        dps.append(dP)
        Us.append(U.copy())
    return dict(strips=strips, delta_p=np.stack(dps), proj_u=np.stack(Us))


def golden_dyna():
    base = load_calibration(os.path.join(HERE, "Result.yml"))
    cfg = StackConfig(96, 72, 1280, 6, 4)
    cal = synth.synthetic_calibration(cfg, base)
    scene = synth.make_scene(cfg, cal)
    planes = synth.render_stack(cfg, scene, noise_sigma=1.0, seed=21)
    first = numpy_cv2_pipeline(cfg, cal, planes)
    frames = synth.render_dyna_frames(cfg, cal, 4, stripe_period=14.0, z_step=0.4, noise_sigma=1.5)
    out = numpy_cv2_dyna(frames, first["proj_u"])
    np.savez_compressed(os.path.join(HERE, "dyna_g6.npz"),
                        cfg=np.array([cfg.width, cfg.height, cfg.projector_width, cfg.gray_digits, cfg.phase_steps]),
                        cam=cal.cam, pro=cal.pro, R=cal.R, T=cal.T, frames=frames, U0=first["proj_u"], z0=first["z"],
                        **out)
    print("dyna: nonzero deltaP", [(d != 0).mean() for d in out["delta_p"]])


def make_bmp(pixels: np.ndarray, bpp: int, palette: np.ndarray | None = None, top_down: bool = False,
             clr_used: int = 0) -> bytes:
    """A BITMAPINFOHEADER BMP.  pixels: [H][W] indices (8 bpp) or [H][W][3|4] BGR(A)."""
    import struct
    H, W = pixels.shape[:2]
    bytes_pp = bpp // 8
    stride = (W * bytes_pp + 3) & ~3
    rows = pixels if top_down else pixels[::-1]
    body = bytearray()
    for r in rows:
        raw = np.ascontiguousarray(r, dtype=np.uint8).tobytes()
        body += raw + b"\0" * (stride - len(raw))
    pal = b""
    if bpp == 8:
        n = clr_used if clr_used else 256
        pal = np.concatenate([palette[:n].astype(np.uint8), np.zeros((n, 1), np.uint8)], axis=1).tobytes()   # B G R 0
    off = 14 + 40 + len(pal)
    head = struct.pack("<2sIHHI", b"BM", off + len(body), 0, 0, off)
    info = struct.pack("<IiiHHIIiiII", 40, W, -H if top_down else H, 1, bpp, 0, len(body), 2835, 2835, clr_used, 0)
    return head + info + pal + bytes(body)


def golden_bmp():
    import tempfile
    rng = np.random.Generator(np.random.PCG64(99))
    ident = np.repeat(np.arange(256, dtype=np.uint8)[:, None], 3, axis=1)
    colour = rng.integers(0, 256, (256, 3)).astype(np.uint8)
    cases = {
        "gray8_37x11": make_bmp(rng.integers(0, 256, (11, 37)), 8, ident),
        "gray8_64x16_topdown": make_bmp(rng.integers(0, 256, (16, 64)), 8, ident, top_down=True),
        "pal8_colour_50x9": make_bmp(rng.integers(0, 256, (9, 50)), 8, colour),
        "pal8_used100_21x7": make_bmp(rng.integers(0, 100, (7, 21)), 8, colour, clr_used=100),
        "bgr24_33x10": make_bmp(rng.integers(0, 256, (10, 33, 3)), 24),
        "bgr24_40x6_topdown": make_bmp(rng.integers(0, 256, (6, 40, 3)), 24, top_down=True),
        "bgra32_19x5": make_bmp(rng.integers(0, 256, (5, 19, 4)), 32),
        "bgr24_extremes_16x4": make_bmp(np.array([[[255, 255, 255], [0, 0, 0], [255, 0, 0], [0, 255, 0], [0, 0, 255],
                                                   [1, 1, 1], [254, 255, 253], [128, 127, 129]] * 2] * 4), 24),
    }
    out = {}
    with tempfile.TemporaryDirectory() as d:
        for name, data in cases.items():
            path = os.path.join(d, name + ".bmp")
            with open(path, "wb") as f:
                f.write(data)
            img = cv2.imread(path, cv2.IMREAD_GRAYSCALE)      # CSensorV.cpp:111-114
            assert img is not None and img.dtype == np.uint8, name
            out[name + "__file"] = np.frombuffer(data, np.uint8)
            out[name + "__gray"] = img
    np.savez_compressed(os.path.join(HERE, "bmp_cases.npz"), **out)
    print("bmp cases", list(cases))


def golden_reference_run():
    """Run the reference's own compiled sources and keep what they produce."""
    from oracle import ref_runner as R          # builds / runs the reference binary only (not the oracle port)
    if not R.build():
        print("reference binary unavailable; keeping the committed reference_run fixture")
        return
    base = load_calibration(os.path.join(HERE, "Result.yml"))
    cfg = StackConfig(72, 56, 1280, 6, 4)       # the reference's own digit counts and projector width
    cal = synth.synthetic_calibration(cfg, base)
    scene = synth.make_scene(cfg, cal)
    planes = synth.render_stack(cfg, scene, noise_sigma=1.0, seed=17)
    frames = synth.render_dyna_frames(cfg, cal, 4, stripe_period=14.0, z_step=0.4, noise_sigma=1.5)
    ws = R.Workspace(cfg, cal, planes, frames)
    full = ws.run_full()
    first = ws.run_first()
    ws.close()
    out = {"planes": planes, "dyna_frames": frames, "cam": cal.cam, "pro": cal.pro, "R": cal.R, "T": cal.T,
           "gray": first["gray"], "phase": first["phase"], "A": first["A"], "B": first["B"], "P": first["P"],
           "cC": first["cC"], "cD": first["cD"]}
    for f, fr in enumerate(full["frames"]):
        for k, v in fr.items():
            out[f"f{f}_{k}"] = v
        out[f"cloud{f}"] = np.frombuffer(full["clouds"][f], np.uint8)
    assert np.array_equal(first["projU"], full["frames"][0]["projU"]) and np.array_equal(first["z"], full["frames"][0]["z"])
    np.savez_compressed(os.path.join(HERE, "reference_run_g6n4.npz"), **out)
    print("reference run:", {k: (v.shape if hasattr(v, "shape") else v) for k, v in list(out.items())[:6]}, "...")


def golden_kat():
    # SURVEY.md 8(c) KAT-E (G=6, PW=1280, T=40, gp=20) and KAT-T (Result.yml)
    kat_e = np.array([
        # U_true, I0, I1, I2, I3, grayCode, kbin, grayVal, fastAtan(deg), pix, corr, U_decoded
        [347.30, 16, 66, 238, 188, 25, 17, 340, 241.20962524414062, 27.301069259643555, 0, 347.30106925964355],
        [320.20, 121, 254, 133, 0, 24, 16, 320, 357.295654296875, 0.19951629638671875, 0, 320.1995162963867],
        [339.90, 139, 1, 115, 253, 24, 16, 320, 174.5604705810547, 19.895606994628906, 0, 339.8956069946289],
        [359.95, 116, 254, 138, 0, 25, 17, 340, 355.0505065917969, 39.95005798339844, 0, 359.95005798339844],
        [0.60, 129, 254, 125, 0, 0, 0, 0, 0.9020314812660217, 0.6002257466316223, 0, 0.6002257466316223],
        [1279.40, 105, 252, 149, 2, 32, 63, 1260, 350.01837158203125, 39.39093017578125, 0, 1279.3909301757812],
    ], dtype=np.float64)
    kat_t = np.array([
        # u, v, U, cC, cD, z, x, y, inFOV
        [320, 256, 640.0, 1.2926350522e9, 1.4560922605e6, 26.684693078303, 0.010991564619, 0.010972373744, 1],
        [0, 0, 300.25, 5.2914256140e8, 1.4741196490e6, 88.385787100032, -23.263796885137, -18.571274707752, 1],
        [639, 511, 900.5, 2.0537325088e9, 1.4380963737e6, 14.685488718517, 3.865329912375, 3.085657255057, 1],
        [500, 50, 777.125, 1.7061394634e9, 1.4024924342e6, 16.916177918698, 2.515398233678, -2.858791264529, 1],
        [100, 400, 512.0, 7.8232606973e8, 1.5082204172e6, 877.825796462349, 0, 0, 0],
    ], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "kat.npz"), kat_e=kat_e, kat_t=kat_t,
                        A=np.float64(-5901248111.511674), B=np.float64(5820127.3056066735))


if __name__ == "__main__":
    golden_reference_run()
    golden_bmp()
    golden_fast_atan2()
    if os.path.isdir(REF):
        golden_gray_code()
    golden_pipeline()
    golden_dyna()
    golden_kat()
