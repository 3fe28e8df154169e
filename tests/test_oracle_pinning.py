"""CPU tests: pin the oracle against the committed golden fixtures
(tests/golden/make_golden.py: container cv2 4.13 scalar primitives + an
independent numpy/cv2 restatement of the reference loops) and the SURVEY 8(c)
known-answer tables."""
import hashlib
import os

import numpy as np
import pytest

from conftest import GOLDEN, bits_equal, make_case, oracle_run


def test_fast_atan2_matches_cv2_on_every_reachable_4step_input(oracle):
    g = np.load(os.path.join(GOLDEN, "fast_atan2.npz"))
    v = (np.arange(-255, 256) / 2).astype(np.float32)
    s, c = np.meshgrid(v, v, indexing="ij")
    table = oracle.fast_atan2(s.ravel(), c.ravel())
    assert hashlib.sha256(table.tobytes()).hexdigest() == str(g["sha256"])
    assert bits_equal(table[g["idx"]], g["sample"])
    assert bits_equal(oracle.fast_atan2(g["free_y"], g["free_x"]), g["free_out"])
    assert oracle.fast_atan2(np.float32(0), np.float32(0))[()] == 0.0
    # accuracy of the polynomial itself vs true atan2: 1.67e-4 rad (SURVEY hard part 1)
    true = np.degrees(np.arctan2(s.astype(np.float64), c.astype(np.float64))) % 360
    d = np.abs(table.reshape(s.shape) - true)
    d = np.minimum(d, 360 - d)
    assert 1.6e-4 < np.radians(d.max()) < 1.7e-4


def test_atan_coefficients_bit_patterns(oracle):
    # the f32 constants hard-coded in csrc/slc_device.cuh
    want = [0x4265226F, 0xC19556EE, 0x410E9FBF, 0xC0228AD9]
    got = [np.float32(np.float32(c) * np.float32(180 / np.pi)).view(np.uint32)
           for c in (0.9997878412794807, -0.3258083974640975, 0.1555786518463281, -0.04432655554792128)]
    assert [int(x) for x in got] == want
    for bits, txt in zip(want, ("0x1.ca44dep+5", "-0x1.2aaddcp+4", "0x1.1d3f7ep+3", "-0x1.4515b2p+1")):
        assert np.float32(float.fromhex(txt)).view(np.uint32) == bits


def test_gray_lut_matches_reference_code_file(oracle):
    g = np.load(os.path.join(GOLDEN, "gray_code_g6.npz"))
    assert bits_equal(oracle.default_gray_lut(6), g["lut"])
    rows = g["rows"].astype(np.int64)
    assert np.array_equal(rows[:, 1], rows[:, 0] ^ (rows[:, 0] >> 1))
    ref = "/root/reference/DynaFrame/DynaFrame/Patterns/vGrayCode.txt"
    if os.path.exists(ref):   # only in the build container
        assert np.array_equal(np.loadtxt(ref, dtype=np.int64), rows)
    for G in (1, 7, 9, 10, 12):
        lut = oracle.default_gray_lut(G)
        b = np.arange(1 << G)
        assert np.array_equal(lut[b ^ (b >> 1)], b)


@pytest.mark.parametrize("name", ["pipeline_g6n4", "pipeline_g9n4"])
def test_pipeline_golden_bit_exact(oracle, name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    W, H, PW, G, N = [int(v) for v in g["cfg"]]
    cfg = oracle.make_config(W, H, PW, G, N)
    cal = oracle.make_calib(g["cam"], g["pro"], g["R"], g["T"])
    r = oracle.reconstruct(cfg, cal, g["planes"])
    for key in ("kbin", "gray_val", "phase_pix", "corr", "proj_u", "x", "y", "z", "mask"):
        assert bits_equal(r[key], g[key]), key
    A, B, cC, cD, P = oracle.calibration(cfg, cal)
    assert A == g["A"] and B == g["B"]
    assert bits_equal(P, g["P"]) and bits_equal(cC, g["cC"]) and bits_equal(cD, g["cD"])
    # and the decode lands on the rendered ground truth
    m = r["mask"].astype(bool) & (np.abs(r["proj_u"] - g["true_U"]) < 1)
    assert m.mean() > 0.85
    assert np.abs((r["z"] - g["true_z"])[m]).max() < 0.5


def test_kat_e_single_pixels(oracle):
    kat = np.load(os.path.join(GOLDEN, "kat.npz"))["kat_e"]
    cfg = oracle.make_config(16, 1, 1280, 6, 4)
    cal = oracle.make_calib(np.eye(3), np.eye(3), np.eye(3), np.array([1.0, 0, 1.0]))
    for row in kat:
        U_true, I0, I1, I2, I3, gcode, kbin, gval, deg, pix, corr, U = row
        planes = np.zeros((16, 1, 16), np.uint8)
        for b in range(6):
            bit = (int(gcode) >> b) & 1
            planes[2 * b] = 200 if bit else 20
            planes[2 * b + 1] = 20 if bit else 200
        for k, val in enumerate((I0, I1, I2, I3)):
            planes[12 + k] = int(val)
        r = oracle.reconstruct(cfg, cal, planes)
        assert r["kbin"][0, 0] == kbin and r["gray_val"][0, 0] == gval
        s, c = (I0 - I2) / 2, (I1 - I3) / 2
        assert float(oracle.fast_atan2(np.float32(s), np.float32(c)).reshape(-1)[0]) == deg
        assert r["phase_pix"][0, 0] == pix
        assert r["corr"][0, 0] == corr
        assert r["proj_u"][0, 0] == U
        assert abs(U - U_true) < 0.02


def test_kat_t_triangulation(oracle, base_calibration):
    k = np.load(os.path.join(GOLDEN, "kat.npz"))
    cfg = oracle.make_config(640, 512, 1280, 6, 4)
    c = base_calibration
    A, B, cC, cD, _ = oracle.calibration(cfg, oracle.make_calib(c.cam, c.pro, c.R, c.T))
    assert A == float(k["A"]) and B == float(k["B"])
    for u, v, U, eC, eD, ez, ex, ey, infov in k["kat_t"]:
        u, v = int(u), int(v)
        assert abs(cC[v, u] / eC - 1) < 1e-9 and abs(cD[v, u] / eD - 1) < 1e-9
        z = -(A - B * U) / (cC[v, u] - cD[v, u] * U)
        assert abs(z - ez) < 1e-9
        assert (10 <= z <= 100) == bool(infov)
        if infov:
            assert abs(z * (u - c.cam[0, 2]) / c.cam[0, 0] - ex) < 1e-9
            assert abs(z * (v - c.cam[1, 2]) / c.cam[1, 1] - ey) < 1e-9


def test_wrap_correction_branches(oracle):
    """Every branch of CCalculation.cpp:570-584, including U == 0 -> invalid."""
    G, PW = 3, 64   # gp = 8, T = 16
    cfg = oracle.make_config(16, 1, PW, G, 4)
    cal = oracle.make_calib(np.diag([100.0, 100.0, 1.0]), np.diag([100.0, 100.0, 1.0]), np.eye(3),
                            np.array([-5.0, 0.0, 0.0]))
    planes = np.zeros((2 * G + 4, 1, 16), np.uint8)
    planes[1::2][:G] = 100          # all inverse images brighter -> code 0 -> kbin 0 (even)
    # phase exactly T: s = -0, c = +: x = 360*(T-0.5)/T = 348.75 deg -> pix == T -> U == 0
    # choose (s, c) by search over u8 values for pix > 0.75 T (corr -1) in column 0
    planes[2 * G + 0, 0, 0], planes[2 * G + 2, 0, 0] = 0, 200     # s = -100
    planes[2 * G + 1, 0, 0], planes[2 * G + 3, 0, 0] = 100, 100   # c = 0 -> 270 deg -> pix = 12.5 > 12
    r = oracle.reconstruct(cfg, cal, planes)
    assert r["kbin"][0, 0] == 0 and r["corr"][0, 0] == -1
    assert r["proj_u"][0, 0] == 12.5 - 16
    # odd kbin with small phase -> +T - T/2
    planes[0], planes[1] = 100, 0   # bit 0 set -> code 1 -> kbin 1
    planes[2 * G + 0, 0, 0], planes[2 * G + 2, 0, 0] = 110, 100   # s = +5
    planes[2 * G + 1, 0, 0], planes[2 * G + 3, 0, 0] = 200, 0     # c = +100 -> ~2.86 deg
    r = oracle.reconstruct(cfg, cal, planes)
    assert r["kbin"][0, 0] == 1 and r["corr"][0, 0] == 1
    assert abs(r["proj_u"][0, 0] - (8 + r["phase_pix"][0, 0] + 16 - 8)) < 1e-12


def test_threads_do_not_change_results(oracle, base_calibration):
    from structured_light_calculation_b200.configs import CONFIGS
    cfg = CONFIGS["config3"].with_(width=160, height=96)
    cal, _, planes = make_case(cfg, base_calibration, noise=2.0, seed=3)
    a = oracle_run(oracle, cfg, cal, planes, threads=1)
    b = oracle_run(oracle, cfg, cal, planes, threads=4)
    for k in a:
        assert bits_equal(a[k], b[k]), k


def test_ext_nstep_phase_agrees_with_true_phase(oracle, base_calibration):
    """[EXT] N = 8 / 12 / 5: decoded offsets track the rendered truth (noise-free)."""
    from structured_light_calculation_b200.configs import StackConfig
    for N in (8, 12, 5, 3):
        cfg = StackConfig(128, 64, 2048, 8, N)
        cal, scene, planes = make_case(cfg, base_calibration, noise=0.0)
        r = oracle_run(oracle, cfg, cal, planes)
        m = r["mask"].astype(bool) & scene.lit & (scene.albedo > 0.2)
        err = np.abs(r["proj_u"] - scene.U)[m]
        assert np.percentile(err, 99) < 0.08, (N, np.percentile(err, 99))


def test_modulation_mask(oracle, base_calibration):
    from structured_light_calculation_b200.configs import CONFIGS
    cfg = CONFIGS["config3"].with_(width=160, height=128)
    cal, scene, planes = make_case(cfg, base_calibration, noise=1.0)
    r = oracle_run(oracle, cfg, cal, planes)
    dark = scene.albedo < 0.05
    assert dark.sum() > 100
    assert r["mod_ok"][dark].mean() < 0.05          # low-albedo patch rejected
    assert r["mod_ok"][scene.lit & (scene.albedo > 0.3)].mean() > 0.99
    assert not (r["mask"].astype(bool) & ~r["mod_ok"].astype(bool)).any()
    off = oracle_run(oracle, cfg.with_(modulation_min=0.0), cal, planes)
    assert off["mod_ok"].all()


def test_invalid_configurations_rejected(oracle):
    cal = oracle.make_calib(np.eye(3), np.eye(3), np.eye(3), np.zeros(3))
    planes = np.zeros((1, 1, 1), np.uint8)
    for bad in (dict(gray_digits=0), dict(gray_digits=17), dict(phase_steps=2), dict(projector_width=32, gray_digits=6)):
        kw = dict(width=16, height=1, projector_width=1280, gray_digits=6, phase_steps=4)
        kw.update(bad)
        cfg = oracle.make_config(**kw)
        with pytest.raises(ValueError):
            oracle.reconstruct(cfg, cal, np.zeros((2 * max(kw["gray_digits"], 0) + kw["phase_steps"], 1, 16), np.uint8))
    del planes


def test_div360_fma_sequence_equals_ieee_division(oracle):
    """The CUDA kernel divides the angle by 360 with q0 = x*y, r = fma(-q0, 360, x),
    q = fma(r, y, q0), y = RN(1/360) (csrc/slc_device.cuh div360_rn).  Exhaustive proof that
    this equals the reference's IEEE `x / 360` for every f32 angle in [1e-30, 360.5]."""
    bad, n = oracle.check_div360(1e-30, 360.5)
    assert n > 900_000_000 and bad == 0
    assert np.float32(1.0) / np.float32(360.0) == np.float32(float.fromhex("0x1.6c16c2p-9"))


def test_dynamic_frame_golden_bit_exact(oracle):
    """StripRegression / FillOtherDeltaProU / U accumulation against the numpy + cv2.blur fixture."""
    g = np.load(os.path.join(GOLDEN, "dyna_g6.npz"))
    W, H, PW, G, N = [int(v) for v in g["cfg"]]
    cfg = oracle.make_config(W, H, PW, G, N)
    cal = oracle.make_calib(g["cam"], g["pro"], g["R"], g["T"])
    frames = g["frames"]
    B0, W0 = oracle.strip_regression(cfg, frames[0])
    assert bits_equal(B0, g["strips"][0, ..., 0]) and bits_equal(W0, g["strips"][0, ..., 1])
    seq = oracle.dyna_sequence(cfg, cal, g["U0"], g["z0"], frames)
    assert len(seq) == frames.shape[0] - 1
    for f, r in enumerate(seq):
        assert bits_equal(r["strip_b"], g["strips"][f + 1, ..., 0]) and bits_equal(r["strip_w"], g["strips"][f + 1, ..., 1])
        assert bits_equal(r["delta_p"], g["delta_p"][f])
        assert bits_equal(r["proj_u"], g["proj_u"][f])
    # U accumulates exactly: every addend is a multiple of 2^-27 far inside a double
    total = g["U0"] + g["delta_p"].astype(np.float64).sum(axis=0)
    assert bits_equal(seq[-1]["proj_u"], total)


def test_result_text_restatement_is_printf(oracle):
    """CCalculation::Result (CCalculation.cpp:323-357): the oracle's loop against an independent
    Python rendering (u outer / v inner, z filter, '%g' == ostream << double)."""
    rng = np.random.default_rng(5)
    W, H = 37, 23
    x = rng.uniform(-30, 30, (H, W))
    y = rng.uniform(-30, 30, (H, W))
    z = rng.uniform(0, 120, (H, W))
    z[rng.random((H, W)) < 0.2] = 0.0
    x[3, 4] = 1.25e-7
    y[3, 4] = -123456.5
    z[3, 4] = 10.0
    cfg = oracle.make_config(W, H, 1280, 6, 4)
    text, n = oracle.result_text(cfg, x, y, z)
    lines = []
    for u in range(W):
        for v in range(H):
            if z[v, u] < 10 or z[v, u] > 100:
                continue
            lines.append("%g %g %g\n" % (x[v, u], y[v, u], z[v, u]))
    assert text == "".join(lines).encode() and n == len(lines)
    crlf3, _ = oracle.result_text(cfg, x, y, z, oracle.TEXT_CRLF | oracle.TEXT_EXP3)
    assert crlf3 == "".join(lines).replace("e-07", "e-007").replace("\n", "\r\n").encode()
