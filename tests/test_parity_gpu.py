"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against
the CPU oracle on identical seeded inputs.

Bars (BASELINE.json north_star):
  * Gray period index, wrap correction, phase offset, ProjectorU (f64) and the
    validity mask: BIT-EXACT;
  * unwrapped phase 2*pi*U/T taken from the f32 `w` output: <= 1e-4 rad;
  * XYZ: <= 1e-5 of the depth range (range 10..100 => 9e-4 units).  The residual
    is FP32-vs-f64: the kernel solves z in f32 (f64 only inside guard bands at the
    FOV limits, which is what keeps the mask bit-exact) and stores f32.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, bits_equal, make_case, oracle_run

pytestmark = pytest.mark.gpu

XYZ_REL_TOL = 1e-5      # of the depth range
PHASE_TOL_RAD = 1e-4


def _reconstructor(cfg, cal, **kw):
    from structured_light_calculation_b200 import capi
    rec = capi.Reconstructor(cfg, device=0, **kw)
    rec.set_calibration(cal)
    return rec


# named test geometries whose f32 `w` cannot meet 1e-4 rad (half an ulp of U below PW, times 2 pi / T):
# PW 4096 with T = 4 (1.9e-4 rad) and PW 65536 with T = 2 (6.1e-3 rad).  Their exact U is the proj_u plane.
W_PHASE_OUT_OF_RANGE = {"generic_g11n4", "g16"}


def w_phase_bound(cfg):
    """Worst-case error of f32(U) as an unwrapped phase, for U inside the projector raster."""
    below = np.nextafter(np.float32(cfg.projector_width), np.float32(0))
    return float(np.spacing(below)) / 2 * 2 * np.pi / cfg.phase_period


def check_parity(got, want, cfg, stack=0, w_out_of_range_ok=False):
    assert bits_equal(got["kbin"][stack], want["kbin"]), "period index"
    assert bits_equal(got["corr"][stack], want["corr"]), "wrap correction"
    assert bits_equal(got["phase_pix"][stack].astype(np.float64), want["phase_pix"]), "phase offset"
    assert bits_equal(got["proj_u"][stack], want["proj_u"]), "ProjectorU"
    assert bits_equal(got["mask"][stack], want["mask"]), "validity mask"
    T = cfg.phase_period
    # w is ProjectorU rounded ONCE to f32 (asserted bit for bit), so as a phase its error is at most
    # half an f32 ulp of U times 2 pi / T.  Inside the projector raster (U < PW) that is <= 1e-4 rad exactly
    # when w_phase_bound(cfg) <= 1e-4 -- every BASELINE geometry (PW <= 4096 with T >= 8: 1.2e-4 px x 2 pi / 8
    # = 9.6e-5 rad); a pixel decoded one wrap beyond the last column (U >= PW, where nothing was projected)
    # sits in the next binade.  include/slcalc_b200.h states this range; outside it the exact U is the f64
    # proj_u plane, and the named geometries that are outside are listed in W_PHASE_OUT_OF_RANGE.
    assert bits_equal(got["xyzw"][stack, :, :, 3], want["proj_u"].astype(np.float32)), "w != f32(U)"
    w = got["xyzw"][stack, :, :, 3].astype(np.float64)
    raster = want["proj_u"] < cfg.projector_width
    phase_err = (np.abs(w - want["proj_u"])[raster].max() if raster.any() else 0.0) * 2 * np.pi / T
    bound = w_phase_bound(cfg)
    assert phase_err <= bound * (1 + 1e-12), f"w is not one rounding of U: {phase_err} rad > {bound}"
    if bound <= PHASE_TOL_RAD:
        assert phase_err <= PHASE_TOL_RAD, f"unwrapped phase error {phase_err} rad"
    else:
        assert w_out_of_range_ok or cfg.name in W_PHASE_OUT_OF_RANGE, \
            f"{cfg.name or cfg}: f32 w cannot hold 1e-4 rad here (bound {bound:.2e}) and the geometry is not listed"
    tol = XYZ_REL_TOL * (cfg.fov_max - cfg.fov_min)
    errs = {}
    for ch, key in enumerate("xyz"):
        a = got["xyzw"][stack, :, :, ch].astype(np.float64)
        b = want[key]
        assert np.array_equal(np.isnan(a), np.isnan(b))
        err = np.nanmax(np.abs(a - b)) if a.size else 0.0
        errs[key] = err
        assert err <= tol, f"{key}: {err} > {tol}"
    invalid = want["mask"] == 0
    assert not got["xyzw"][stack, :, :, :3][invalid].any(), "invalid pixels must be (0,0,0)"
    return errs, phase_err


def check_fast_modes(rec, planes, got_parity):
    """The kernel instances without parity planes (MODE 0 / 1, the ones bench.py times)
    must produce exactly what the parity-checked instance (MODE 2) produced."""
    fast = rec.reconstruct(planes, parity=False)
    assert bits_equal(fast["xyzw"], got_parity["xyzw"]), "MODE 0/1 xyzw differs from MODE 2"
    assert bits_equal(fast["mask"], got_parity["mask"]), "MODE 0/1 mask differs from MODE 2"
    check_result_formats(rec, planes, got_parity)


def check_result_formats(rec, planes, full):
    """SLC_RESULT_DEPTH / SLC_RESULT_POINTS are selections of the checked xyzw + mask output, bit for bit."""
    from structured_light_calculation_b200 import capi
    cfg = rec.cfg
    n, npx = full["mask"].shape[0], cfg.pixels
    mask = full["mask"].reshape(n, npx)
    d = rec.reconstruct_ex(planes, capi.SLC_RESULT_DEPTH)
    assert bits_equal(d["depth"], np.ascontiguousarray(full["xyzw"][..., 2])), "DEPTH z differs from the xyzw map"
    assert np.array_equal(capi.unpack_mask_bits(d["mask_bits"], n, npx), mask), "DEPTH bit mask differs from the byte mask"
    if cfg.width % 8 != 0:
        return
    for order in (capi.SLC_ORDER_ROW_MAJOR, capi.SLC_ORDER_REFERENCE):
        p = rec.reconstruct_ex(planes, capi.SLC_RESULT_POINTS, order)
        assert np.array_equal(capi.unpack_mask_bits(p["mask_bits"], n, npx), mask), "POINTS bit mask"
        for i in range(n):
            xyz, m = full["xyzw"][i, ..., :3], full["mask"][i].astype(bool)
            want = xyz[m] if order == capi.SLC_ORDER_ROW_MAJOR else np.transpose(xyz, (1, 0, 2))[m.T]
            assert int(p["n_points"][i]) == want.shape[0], f"POINTS count, order {order}"
            assert bits_equal(np.ascontiguousarray(p["points"][i, : want.shape[0]]), np.ascontiguousarray(want)), \
                f"POINTS list differs from mask-selecting the map, order {order}"


CASES = [
    # (name, G, N, PW, W, H, noise, modulation)
    ("ref_default", 6, 4, 1280, 256, 96, 1.0, 0.0),
    ("config1", 7, 4, 1280, 320, 128, 1.0, 0.0),
    ("config2", 9, 4, 2560, 384, 160, 2.0, 0.0),
    ("config3", 8, 8, 2048, 272, 128, 1.0, 8.0),
    ("config5", 10, 12, 4096, 512, 96, 1.0, 0.0),
    ("generic_g5n6", 5, 6, 640, 128, 64, 1.0, 4.0),
    ("generic_g11n4", 11, 4, 4096, 128, 64, 0.0, 0.0),
    ("odd_n5", 8, 5, 2048, 128, 64, 1.0, 0.0),
    ("odd_n3", 7, 3, 1280, 128, 64, 1.0, 2.0),
    ("g8n4", 8, 4, 2048, 256, 64, 1.0, 0.0),          # 4-step instances beside the BASELINE ones
    ("g10n4", 10, 4, 4096, 256, 64, 1.0, 3.0),
    ("g5n4", 5, 4, 640, 128, 64, 1.0, 0.0),
    ("g9n8", 9, 8, 2560, 192, 64, 1.0, 0.0),          # fixed Gray depth, generic even / odd step count
    ("g6n7", 6, 7, 1280, 192, 64, 1.0, 2.0),
    ("g1", 1, 4, 64, 64, 32, 1.0, 0.0),
    ("g16", 16, 4, 65536, 64, 32, 0.0, 0.0),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("pxt", [16, 8, 4])
def test_fused_kernel_matches_oracle(built_library, oracle, base_calibration, case, pxt):
    from structured_light_calculation_b200 import capi
    from structured_light_calculation_b200.configs import StackConfig
    name, G, N, PW, W, H, noise, mod = case
    cfg = StackConfig(W, H, PW, G, N, modulation_min=mod, name=name)
    cal, scene, planes = make_case(cfg, base_calibration, noise=noise, seed=G * 100 + N)
    rec = _reconstructor(cfg, cal)
    rec.set_pixels_per_thread(pxt)          # per context: nothing process-wide
    got = rec.reconstruct(planes, parity=True)
    assert rec.launch_count() == 1
    assert rec.info().kernel_variant in (0, 1)
    check_fast_modes(rec, planes, got)
    rec.close()
    want = oracle_run(oracle, cfg, cal, planes)
    check_parity(got, want, cfg)
    del capi


def test_random_geometries_match_oracle(built_library, oracle, base_calibration):
    """A seeded sweep over geometries nobody picked by hand: Gray depth 1..12, 3..16 phase steps, widths
    that are and are not multiples of 4 / 8 / 16 (vector and scalar kernels), projector widths with and
    without a remainder in PW / 2^G, noise and modulation thresholds.  Every output is held to the same bar
    as the named cases, and the kernel instances without parity planes must reproduce the checked one."""
    from structured_light_calculation_b200.configs import StackConfig
    rng = np.random.default_rng(20261018)
    done = in_range = 0
    for trial in range(40):
        G = int(rng.integers(1, 13))
        N = int(rng.integers(3, 17))
        gp = int(rng.integers(1, 12))
        PW = gp * (1 << G) + int(rng.integers(0, 2)) * int(rng.integers(0, 1 << G))   # sometimes not a multiple of 2^G
        if PW > 8192:
            continue
        W = int(rng.choice([40, 52, 64, 72, 97, 128, 136, 200]))
        H = int(rng.integers(9, 49))
        mod = float(rng.choice([0.0, 0.0, 3.0, 10.0]))
        noise = float(rng.choice([0.0, 1.0, 2.5]))
        cfg = StackConfig(W, H, PW, G, N, modulation_min=mod, name=f"rand{trial}")
        cal, scene, planes = make_case(cfg, base_calibration, noise=noise, seed=1000 + trial)
        rec = _reconstructor(cfg, cal)
        got = rec.reconstruct(planes, parity=True)
        check_fast_modes(rec, planes, got)
        rec.close()
        want = oracle_run(oracle, cfg, cal, planes)
        try:
            # a random geometry may lie outside the range in which an f32 w holds 1e-4 rad (w_phase_bound);
            # check_parity then holds w to its exact one-rounding bound instead
            check_parity(got, want, cfg, w_out_of_range_ok=True)
        except AssertionError as e:
            raise AssertionError(f"trial {trial}: W={W} H={H} PW={PW} G={G} N={N} mod={mod} noise={noise}: {e}") from e
        done += 1
        in_range += w_phase_bound(cfg) <= PHASE_TOL_RAD
    assert done >= 25 and 0 < in_range < done      # the sweep saw both sides of the range


@pytest.mark.parametrize("flags_name", ["z_fp64", "scalar"])
def test_kernel_modes_match_oracle(built_library, oracle, base_calibration, flags_name):
    from structured_light_calculation_b200 import capi
    from structured_light_calculation_b200.configs import CONFIGS
    flags = capi.SLC_FLAG_Z_FP64 if flags_name == "z_fp64" else capi.SLC_FLAG_SCALAR_KERNEL
    cfg = CONFIGS["config2"].with_(width=320, height=200)
    cal, scene, planes = make_case(cfg, base_calibration, noise=1.0, seed=21)
    rec = _reconstructor(cfg, cal, flags=flags)
    got = rec.reconstruct(planes, parity=True)
    if flags_name == "scalar":
        assert rec.info().kernel_variant == 2
    rec.close()
    want = oracle_run(oracle, cfg, cal, planes)
    errs, _ = check_parity(got, want, cfg)
    if flags_name == "z_fp64":
        # z is then the f32 rounding of the reference's own f64 value
        z = got["xyzw"][0, :, :, 2]
        assert np.array_equal(z, want["z"].astype(np.float32))


@pytest.mark.parametrize("W,H", [(100, 37), (17, 5), (24, 9), (1, 1), (1000, 3), (36, 7)])
def test_ragged_sizes(built_library, oracle, base_calibration, W, H):
    """Widths that are not a multiple of 16 / 8 / 4 fall back to narrower vectors or the scalar kernel."""
    from structured_light_calculation_b200.configs import StackConfig
    cfg = StackConfig(W, H, 1280, 6, 4)
    cal, scene, planes = make_case(cfg, base_calibration, noise=1.0, seed=W)
    rec = _reconstructor(cfg, cal)
    got = rec.reconstruct(planes, parity=True)
    rec.close()
    check_parity(got, oracle_run(oracle, cfg, cal, planes), cfg)


@pytest.mark.parametrize("name", ["pipeline_g6n4", "pipeline_g9n4"])
def test_golden_fixture_through_cuda(built_library, name):
    """CUDA path against the committed numpy+cv2 fixtures directly (no oracle in between)."""
    from structured_light_calculation_b200.calibration import Calibration
    from structured_light_calculation_b200.configs import StackConfig
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    W, H, PW, G, N = [int(v) for v in g["cfg"]]
    cfg = StackConfig(W, H, PW, G, N)
    rec = _reconstructor(cfg, Calibration(g["cam"], g["pro"], g["R"], g["T"]))
    got = rec.reconstruct(g["planes"], parity=True)
    rec.close()
    want = {k: g[k] for k in ("kbin", "corr", "phase_pix", "proj_u", "mask", "x", "y", "z")}
    check_parity(got, want, cfg)


def test_extreme_inputs(built_library, oracle, base_calibration):
    """All-black, all-white, saturated and random-noise stacks (no structure at all)."""
    from structured_light_calculation_b200.configs import CONFIGS
    cfg = CONFIGS["config1"].with_(width=128, height=48)
    cal, _, _ = make_case(cfg, base_calibration)
    rng = np.random.Generator(np.random.PCG64(99))
    stacks = [
        np.zeros((cfg.planes, cfg.height, cfg.width), np.uint8),
        np.full((cfg.planes, cfg.height, cfg.width), 255, np.uint8),
        rng.integers(0, 256, (cfg.planes, cfg.height, cfg.width), dtype=np.uint8),
        rng.integers(0, 2, (cfg.planes, cfg.height, cfg.width), dtype=np.uint8) * 255,
        rng.integers(126, 130, (cfg.planes, cfg.height, cfg.width), dtype=np.uint8),
    ]
    rec = _reconstructor(cfg, cal, max_batch=len(stacks))
    got = rec.reconstruct(np.stack(stacks), parity=True)
    rec.close()
    for i, st in enumerate(stacks):
        check_parity(got, oracle_run(oracle, cfg, cal, st), cfg, stack=i)


def test_all_4step_phase_inputs_bit_exact(built_library, oracle):
    """Every (I0-I2, I1-I3) pair a 4-step u8 stack can produce, in both kbin parities:
    the arctan polynomial, the offset arithmetic and both wrap branches, exhaustively."""
    from structured_light_calculation_b200.calibration import Calibration
    from structured_light_calculation_b200.configs import StackConfig
    W, H = 512, 511
    cfg = StackConfig(W, H, 1280, 6, 4)
    cal = Calibration(np.array([[2400.0, 0, 255.5], [0, 2400.0, 255.5], [0, 0, 1]]),
                      np.array([[2000.0, 0, 640], [0, 2000.0, 400], [0, 0, 1]]), np.eye(3),
                      np.array([-8.0, 0.0, 0.5]))
    planes = np.zeros((cfg.planes, H, W), np.uint8)
    d = np.arange(-255, 256)
    ds, dc = np.meshgrid(d, d, indexing="ij")          # 511 x 511 differences
    i0 = np.where(ds >= 0, ds, 0); i2 = np.where(ds >= 0, 0, -ds)
    i1 = np.where(dc >= 0, dc, 0); i3 = np.where(dc >= 0, 0, -dc)
    for k, img in enumerate((i0, i1, i2, i3)):
        planes[12 + k, :, :511] = img.astype(np.uint8)
    # kbin parity alternates by row; higher bits vary by column block
    rows = np.arange(H)[:, None]
    cols = np.arange(W)[None, :]
    code = ((rows & 1) | ((cols >> 4) << 1)) & 63
    for b in range(6):
        bit = ((code >> b) & 1).astype(bool)
        planes[2 * b] = np.where(bit, 180, 40)
        planes[2 * b + 1] = np.where(bit, 40, 180)
    rec = _reconstructor(cfg, cal)
    got = rec.reconstruct(planes, parity=True)
    check_fast_modes(rec, planes, got)
    rec.close()
    want = oracle_run(oracle, cfg, cal, planes)
    check_parity(got, want, cfg)
    assert set(np.unique(want["corr"])) == {-1, 0, 1}


def test_device_arctan_and_offset_bit_exact(built_library, oracle):
    """cvFastArctan + the offset arithmetic on the device (safe-range divisions, no FCHK
    slow path) against the oracle's IEEE path: every 4-step pair, every N-step-like sum of
    u8 differences we can draw, plus adversarial magnitudes."""
    from structured_light_calculation_b200 import capi
    from structured_light_calculation_b200.configs import StackConfig
    rng = np.random.Generator(np.random.PCG64(2024))
    v = (np.arange(-255, 256) / 2).astype(np.float32)
    s4, c4 = [a.ravel() for a in np.meshgrid(v, v, indexing="ij")]
    # N-step sums: sum_k d_k * f32(cos(2 pi k / N)) with integer d_k in [-255, 255]
    parts_s, parts_c = [s4], [c4]
    for N in (6, 8, 12, 16, 5, 7):
        n_terms = N // 2 if N % 2 == 0 else N
        k = np.arange(n_terms)
        ck = np.cos(2 * np.pi * k / N).astype(np.float32)
        sk = np.sin(2 * np.pi * k / N).astype(np.float32)
        d = rng.integers(-255, 256, (400_000, n_terms)).astype(np.float32)
        ss = np.zeros(d.shape[0], np.float32)
        cc = np.zeros(d.shape[0], np.float32)
        for j in range(n_terms):   # f32 accumulation (rounding differs from fmaf; any f32 value is fair input)
            ss = (ss + d[:, j] * ck[j]).astype(np.float32)
            cc = (cc + d[:, j] * sk[j]).astype(np.float32)
        parts_s.append(ss)
        parts_c.append(cc)
    # adversarial: tiny / huge ratios, equal magnitudes, zeros, signed zeros
    extra = np.array([0.0, -0.0, 1e-7, -1e-7, 1e-3, 0.5, 1.0, 127.5, 255.0, 16320.0, -16320.0, 3e-5, 65000.0],
                     np.float32)
    es, ec = [a.ravel() for a in np.meshgrid(extra, extra, indexing="ij")]
    parts_s.append(es)
    parts_c.append(ec)
    log_s = (rng.choice([-1, 1], 500_000) * np.exp(rng.uniform(np.log(1e-6), np.log(7e4), 500_000))).astype(np.float32)
    log_c = (rng.choice([-1, 1], 500_000) * np.exp(rng.uniform(np.log(1e-6), np.log(7e4), 500_000))).astype(np.float32)
    parts_s.append(log_s)
    parts_c.append(log_c)
    s = np.concatenate(parts_s)
    c = np.concatenate(parts_c)
    for T in (40, 10, 8, 16, 2):
        cfg = StackConfig(64, 16, T, 1, 4)      # G = 1: phase period == projector width == T
        rec = capi.Reconstructor(cfg, device=0)
        deg, pix = rec.eval_phase(s, c)
        rec.close()
        want_deg = oracle.fast_atan2(s, c)
        assert bits_equal(deg, want_deg), f"arctan differs in {(deg.view(np.uint32) != want_deg.view(np.uint32)).sum()} of {s.size}"
        assert bits_equal(pix, oracle.phase_pix(want_deg, T)), "offset arithmetic differs"


def test_fov_boundary_mask_is_bit_exact(built_library, oracle, base_calibration):
    """A plane swept through z = fov_min and z = fov_max: pixels arbitrarily close to the
    limits must take the same side as the reference's f64 comparison (SURVEY hard part 3)."""
    from structured_light_calculation_b200.configs import StackConfig
    cfg = StackConfig(640, 512, 1280, 6, 4)
    cal = base_calibration
    from structured_light_calculation_b200 import capi
    rec = capi.Reconstructor(cfg, device=0)
    rec.set_calibration(cal)
    ocfg = oracle.make_config(cfg.width, cfg.height, cfg.projector_width, cfg.gray_digits, cfg.phase_steps)
    A, B, cC, cD, _ = oracle.calibration(ocfg, oracle.make_calib(cal.cam, cal.pro, cal.R, cal.T))
    rng = np.random.Generator(np.random.PCG64(5))
    n_flips_seen = 0
    for limit in (cfg.fov_min, cfg.fov_max):
        # U that puts z exactly on the limit, then jitter by a few f32 ulps of U
        U0 = (A + limit * cC) / (B + limit * cD)
        U = U0 + rng.integers(-3, 4, U0.shape) * np.spacing(U0.astype(np.float32)).astype(np.float64)
        xyzw, mask = rec.triangulate(U)
        z = -(A - B * U) / (cC - cD * U)
        want = ((U != 0) & ~((z < cfg.fov_min) | (z > cfg.fov_max))).astype(np.uint8)
        assert bits_equal(mask, want)
        n_flips_seen += int((want == 0).sum() > 1000 and (want == 1).sum() > 1000)
    rec.close()
    assert n_flips_seen == 2   # the sweep really straddled both limits


def test_fused_kernel_fov_guard_band(built_library, oracle, base_calibration):
    """Same question for the fused kernel's f32 solve + f64 guard band: render scenes whose
    surface sits within ~1e-6 of the FOV limits so thousands of pixels are borderline."""
    from structured_light_calculation_b200 import synth
    from structured_light_calculation_b200.configs import StackConfig
    cfg = StackConfig(640, 256, 1280, 6, 4)
    cal = synth.synthetic_calibration(cfg, base_calibration)
    scene = synth.make_scene(cfg, cal)
    ocfg = oracle.make_config(cfg.width, cfg.height, cfg.projector_width, cfg.gray_digits, cfg.phase_steps)
    ocal = oracle.make_calib(cal.cam, cal.pro, cal.R, cal.T)
    planes = synth.render_stack(cfg, scene, noise_sigma=0.5, seed=3)
    base = oracle.reconstruct(ocfg, ocal, planes)
    zs = base["z"][base["mask"] == 1]
    rec = _reconstructor(cfg, cal)
    for lim_lo, lim_hi in ((float(np.median(zs)), 100.0), (10.0, float(np.median(zs))),
                           (float(np.percentile(zs, 30)), float(np.percentile(zs, 70)))):
        c2 = cfg.with_(fov_min=lim_lo, fov_max=lim_hi)
        rec2 = _reconstructor(c2, cal)
        got = rec2.reconstruct(planes, parity=True)
        rec2.close()
        want = oracle_run(oracle, c2, cal, planes)
        assert 0.05 < want["mask"].mean() < 0.95
        check_parity(got, want, c2)
    rec.close()


def test_custom_gray_lut(built_library, oracle, base_calibration):
    """A non-reflected code table (CDecodeGray.cpp:120-125 loads whatever the file holds)."""
    from structured_light_calculation_b200.configs import CONFIGS
    cfg = CONFIGS["config1"].with_(width=160, height=64)
    cal, _, planes = make_case(cfg, base_calibration, seed=8)
    rng = np.random.Generator(np.random.PCG64(1))
    lut = rng.permutation(1 << cfg.gray_digits).astype(np.int16)
    rec = _reconstructor(cfg, cal)
    rec.set_gray_lut(lut)
    got = rec.reconstruct(planes, parity=True)
    check_parity(got, oracle_run(oracle, cfg, cal, planes, lut=lut), cfg)
    # handing back the standard table returns to the arithmetic decode
    rec.set_gray_lut(oracle.default_gray_lut(cfg.gray_digits))
    got = rec.reconstruct(planes, parity=True)
    check_parity(got, oracle_run(oracle, cfg, cal, planes), cfg)
    rec.close()


def test_decoder_objects(built_library, oracle, base_calibration):
    """CDecodeGray / CDecodePhase stand-alone entry points (f64 planes like GetResult())."""
    from structured_light_calculation_b200.configs import CONFIGS
    for cfg in (CONFIGS["reference_default"].with_(width=200, height=50),
                CONFIGS["config3"].with_(width=96, height=40)):
        cal, _, planes = make_case(cfg, base_calibration, seed=4)
        want = oracle_run(oracle, cfg, cal, planes)
        rec = _reconstructor(cfg, cal)
        val, kbin = rec.decode_gray(planes[: 2 * cfg.gray_digits])
        pix, mod = rec.decode_phase(planes[2 * cfg.gray_digits:])
        xyzw, mask = rec.triangulate(want["proj_u"])
        rec.close()
        assert bits_equal(val, want["gray_val"]) and bits_equal(kbin, want["kbin"])
        assert bits_equal(pix, want["phase_pix"]) and bits_equal(mod, want["mod_ok"])
        # triangulate() knows nothing about modulation: compare where modulation passed
        ok = want["mod_ok"].astype(bool)
        assert np.array_equal(mask[ok], want["mask"][ok])
        assert np.array_equal(xyzw[..., 2][ok], want["z"].astype(np.float32)[ok])


def test_batched_and_pipelined_host_paths(built_library, oracle, base_calibration):
    from structured_light_calculation_b200 import capi
    from structured_light_calculation_b200.configs import CONFIGS
    cfg = CONFIGS["config2"].with_(width=192, height=80)
    cal, scene, _ = make_case(cfg, base_calibration)
    from structured_light_calculation_b200 import synth
    n = 7
    stacks = np.stack([synth.render_stack(cfg, scene, noise_sigma=1.5, seed=100 + i) for i in range(n)])
    wants = [oracle_run(oracle, cfg, cal, stacks[i]) for i in range(n)]
    # (a) one launch for the whole batch
    rec = _reconstructor(cfg, cal, max_batch=n, num_slots=1)
    got = rec.reconstruct(stacks, parity=True)
    assert rec.launch_count() == 1
    rec.close()
    for i in range(n):
        check_parity(got, wants[i], cfg, stack=i)
    # (b) chunked over 3 stream slots, max_batch 2 -> 4 launches, same answers
    rec = _reconstructor(cfg, cal, max_batch=2, num_slots=3)
    got2 = rec.reconstruct(stacks, parity=True)
    assert rec.launch_count() == 4
    for k in got:
        assert bits_equal(got[k], got2[k]), k
    # (c) explicit double buffering with pinned memory
    pin_in = capi.PinnedArray(stacks.shape, np.uint8)
    pin_in.array[...] = stacks
    pin_xyzw = capi.PinnedArray((n, cfg.height, cfg.width, 4), np.float32)
    pin_mask = capi.PinnedArray((n, cfg.height, cfg.width), np.uint8)
    sb = cfg.stack_bytes
    for i in range(n):
        slot = i % 3
        if i >= 3:
            rec.wait(slot)
        rec.submit(slot, pin_in.ptr + i * sb, 1, pin_xyzw.ptr + i * cfg.pixels * 16, pin_mask.ptr + i * cfg.pixels)
    with pytest.raises(capi.SlcError):      # slot still in flight
        rec.submit((n - 1) % 3, pin_in.ptr, 1, pin_xyzw.ptr, pin_mask.ptr)
    for s in range(3):
        rec.wait(s)
    assert bits_equal(pin_xyzw.array, got["xyzw"]) and bits_equal(pin_mask.array, got["mask"])
    rec.close()


def test_error_paths(built_library, base_calibration):
    from structured_light_calculation_b200 import capi
    from structured_light_calculation_b200.configs import CONFIGS, StackConfig
    cfg = CONFIGS["config1"].with_(width=64, height=32)
    rec = capi.Reconstructor(cfg, device=0)
    with pytest.raises(capi.SlcError) as e:     # Calculate* before Init (CCalculation.cpp:176-181)
        rec.reconstruct(np.zeros((cfg.planes, cfg.height, cfg.width), np.uint8))
    assert e.value.status == capi.SLC_ERR_NOT_INITIALISED
    with pytest.raises(capi.SlcError):
        rec.set_gray_lut(np.zeros(5, np.int16))
    rec.close()
    for bad in (StackConfig(64, 32, 1280, 0, 4), StackConfig(64, 32, 1280, 17, 4), StackConfig(64, 32, 1280, 6, 2),
                StackConfig(64, 32, 32, 6, 4), StackConfig(0, 32, 1280, 6, 4), StackConfig(64, 32, 1280, 6, 66)):
        with pytest.raises(capi.SlcError) as e:
            capi.Reconstructor(bad, device=0)
        assert e.value.status == capi.SLC_ERR_INVALID_ARG
    with pytest.raises(capi.SlcError):
        capi.Reconstructor(cfg, device=1000)


def test_empty_batch_is_a_no_op(built_library, base_calibration):
    """Zero frame sets: every reconstruct entry point returns SLC_OK without touching a buffer (NULL
    pointers allowed) or launching anything; a dynamic sequence needs its first frame (n_frames >= 1)."""
    from structured_light_calculation_b200 import capi
    from structured_light_calculation_b200.configs import CONFIGS
    cfg = CONFIGS["config1"].with_(width=64, height=32)
    cal, _, _ = make_case(cfg, base_calibration)
    rec = _reconstructor(cfg, cal)
    before = rec.launch_count()
    out = rec.reconstruct(np.zeros((0, cfg.planes, cfg.height, cfg.width), np.uint8), parity=True)
    assert out["xyzw"].shape == (0, cfg.height, cfg.width, 4) and out["mask"].shape == (0, cfg.height, cfg.width)
    rec._check(rec.lib.slc_reconstruct_host(rec.h, None, 0, None, None, None))
    rec._check(rec.lib.slc_reconstruct_device(rec.h, None, 0, None, None, None, None))
    assert rec.launch_count() == before
    with pytest.raises(capi.SlcError):
        rec._check(rec.lib.slc_reconstruct_host(rec.h, None, 1, None, None, None))       # NULL buffers, one frame set
    with pytest.raises(capi.SlcError):
        rec._check(rec.lib.slc_reconstruct_device(rec.h, None, -1, None, None, None, None))
    with pytest.raises(capi.SlcError):
        rec._check(rec.lib.slc_dyna_track_device(rec.h, None, 0, 21, None, None, None, None, None, None))
    rec.close()


@pytest.mark.parametrize("name", ["config1", "config2", "config3", "config5"])
def test_full_size_properties(built_library, oracle, base_calibration, name):
    """BASELINE sizes: full parity against the (multi-threaded) oracle on one stack, plus
    size-independent properties on a batch: batch == single, determinism, ground truth."""
    from structured_light_calculation_b200 import synth
    from structured_light_calculation_b200.configs import CONFIGS
    cfg = CONFIGS[name]
    cal, scene, planes = make_case(cfg, base_calibration, noise=1.0, seed=77)
    rec = _reconstructor(cfg, cal, max_batch=3)
    got = rec.reconstruct(planes, parity=True)
    want = oracle_run(oracle, cfg, cal, planes, threads=min(8, oracle.max_threads()))
    errs, perr = check_parity(got, want, cfg)
    print(f"{name}: max |dx|,|dy|,|dz| = {errs}, phase err {perr:.3e} rad")
    # decode recovers the rendered ground truth
    m = want["mask"].astype(bool) & scene.lit & (scene.albedo > 0.2) & (scene.z < cfg.fov_max)
    assert np.percentile(np.abs(got["xyzw"][0, :, :, 2] - scene.z)[m], 99) < 0.25
    # batch of 3 (two copies + a different stack) == singles; rerun is deterministic
    p2 = synth.render_stack(cfg, scene, noise_sigma=2.0, seed=78)
    batch = rec.reconstruct(np.stack([planes, p2, planes]))
    assert bits_equal(batch["xyzw"][0], got["xyzw"][0]) and bits_equal(batch["xyzw"][2], got["xyzw"][0])
    assert bits_equal(batch["mask"][0], got["mask"][0])
    again = rec.reconstruct(np.stack([planes, p2, planes]))
    assert bits_equal(again["xyzw"], batch["xyzw"]) and bits_equal(again["mask"], batch["mask"])
    rec.close()


@pytest.mark.parametrize("W,H,window,n_frames", [(320, 128, 21, 6), (200, 75, 21, 4), (96, 64, 9, 5), (64, 48, 33, 3),
                                                 (512, 160, 21, 4),    # interior tiles of the window-21 fast path
                                                 (204, 70, 21, 3),     # W % 8 != 0: generic 3x3-sum kernel
                                                 (202, 66, 21, 3)])    # W % 4 != 0: generic strip kernel
def test_dynamic_frames_match_oracle(built_library, oracle, base_calibration, W, H, window, n_frames):
    """CalculateOther (StripRegression + FillOtherDeltaProU + FillCoordinate + deltaZ): strips,
    blurred deltaP and the accumulated ProjectorU bit-exact, mask bit-exact, XYZ / deltaZ in tolerance."""
    from structured_light_calculation_b200 import synth
    from structured_light_calculation_b200.configs import StackConfig
    cfg = StackConfig(W, H, 1280, 6, 4)
    cal, scene, planes = make_case(cfg, base_calibration, noise=1.0, seed=9)
    first = oracle_run(oracle, cfg, cal, planes)
    frames = synth.render_dyna_frames(cfg, cal, n_frames, stripe_period=14.0, z_step=0.4, noise_sigma=1.5)
    if window == 33:   # exercise ties: large flat regions; and the largest deltas the window allows
        frames[:, : H // 2, :] = 100
        ramp = (np.arange(W) * 7 % 256).astype(np.uint8)
        frames[1, H // 2:, :] = ramp                      # extremum at one end of the window ...
        frames[2, H // 2:, :] = ramp[::-1]                # ... then at the other: |delta| up to 31
    if W in (512, 204):  # ties inside the fast path: flat, saturated and two-level regions
        frames[:, : H // 3, : W // 2] = 255
        frames[1:, H // 3: H // 2, W // 4:] = 0
        frames[:, H // 2: H // 2 + 30, ::7] = 17
        frames[:, H // 2: H // 2 + 30, 1::7] = 17
    ocfg = oracle.make_config(W, H, cfg.projector_width, cfg.gray_digits, cfg.phase_steps)
    ocal = oracle.make_calib(cal.cam, cal.pro, cal.R, cal.T)
    want = oracle.dyna_sequence(ocfg, ocal, first["proj_u"], first["z"], frames, window)
    rec = _reconstructor(cfg, cal)
    got = rec.dyna_track(frames, first["proj_u"], window=window, parity=True)
    plain = rec.dyna_track(frames, first["proj_u"], window=window, parity=False)
    assert rec.launch_count() == (4 if W % 8 == 0 else 6)      # strips + fused track, or strips + 3x3 sums + track
    rec.close()
    assert bits_equal(plain["xyzw"], got["xyzw"]) and bits_equal(plain["mask"], got["mask"])
    B0, W0 = oracle.strip_regression(ocfg, frames[0], window)
    assert bits_equal(got["strips"][0, ..., 0].astype(np.float32), B0)
    assert bits_equal(got["strips"][0, ..., 1].astype(np.float32), W0)
    tol = XYZ_REL_TOL * (cfg.fov_max - cfg.fov_min)
    moved = 0
    for f, w in enumerate(want):
        assert bits_equal(got["strips"][f + 1, ..., 0].astype(np.float32), w["strip_b"]), f"stripB frame {f + 1}"
        assert bits_equal(got["strips"][f + 1, ..., 1].astype(np.float32), w["strip_w"]), f"stripW frame {f + 1}"
        assert bits_equal(got["delta_p"][f], w["delta_p"]), f"deltaP frame {f + 1}"
        assert bits_equal(got["proj_u"][f], w["proj_u"]), f"ProjectorU frame {f + 1}"
        assert bits_equal(got["mask"][f], w["mask"]), f"mask frame {f + 1}"
        assert bits_equal(got["xyzw"][f, ..., 3], w["proj_u"].astype(np.float32))
        for ch, key in enumerate("xyz"):
            assert np.abs(got["xyzw"][f, ..., ch] - w[key]).max() <= tol, (f, key)
        assert np.abs(got["delta_z"][f] - w["delta_z"]).max() <= 2 * tol, f"deltaZ frame {f + 1}"
        moved += int((w["delta_p"] != 0).sum())
    if window == 21:
        assert moved > 0     # the tracker saw motion


def test_dynamic_frames_random_sweep(built_library, oracle, base_calibration):
    """Seeded sweep of the dynamic path: widths that take the fused kernel (W % 8 == 0) and the generic
    kernels, every odd window 3..33, heights around the tile sizes, and image content made of rendered
    stripes, pure noise, constant and two-level regions (ties in the window minimum / maximum)."""
    from structured_light_calculation_b200 import synth
    from structured_light_calculation_b200.configs import StackConfig
    rng = np.random.default_rng(977)
    for trial in range(14):
        W = int(rng.choice([64, 72, 96, 128, 136, 150, 168, 256, 260, 384]))
        H = int(rng.choice([40, 47, 64, 65, 72, 96, 130]))
        window = int(rng.choice(np.arange(3, 35, 2)))
        if W <= window or H <= window:
            continue
        n_frames = int(rng.integers(2, 7))
        cfg = StackConfig(W, H, 1280, 6, 4)
        cal, scene, planes = make_case(cfg, base_calibration, noise=1.0, seed=300 + trial)
        first = oracle_run(oracle, cfg, cal, planes)
        frames = synth.render_dyna_frames(cfg, cal, n_frames, stripe_period=float(rng.choice([9.0, 14.0, 23.0])),
                                          z_step=0.4, noise_sigma=1.5)
        kind = trial % 4
        if kind == 1:      # pure noise: extrema anywhere in the window, large deltas
            frames = rng.integers(0, 256, frames.shape, dtype=np.uint8)
        elif kind == 2:    # coarse levels: many equal column sums
            frames = (frames // 64 * 64).astype(np.uint8)
        elif kind == 3:    # flat and saturated blocks next to stripes
            frames[:, : H // 2, : W // 2] = 255
            frames[1:, H // 2:, W // 3: 2 * W // 3] = 0
        ocfg = oracle.make_config(W, H, cfg.projector_width, cfg.gray_digits, cfg.phase_steps)
        ocal = oracle.make_calib(cal.cam, cal.pro, cal.R, cal.T)
        want = oracle.dyna_sequence(ocfg, ocal, first["proj_u"], first["z"], frames, window)
        rec = _reconstructor(cfg, cal)
        got = rec.dyna_track(frames, first["proj_u"], window=window, parity=True)
        plain = rec.dyna_track(frames, first["proj_u"], window=window, parity=False)
        rec.close()
        tag = f"trial {trial}: W={W} H={H} window={window} frames={n_frames} kind={kind}"
        assert bits_equal(plain["xyzw"], got["xyzw"]) and bits_equal(plain["mask"], got["mask"]), tag
        tol = XYZ_REL_TOL * (cfg.fov_max - cfg.fov_min)
        for f, w in enumerate(want):
            assert bits_equal(got["strips"][f + 1, ..., 0].astype(np.float32), w["strip_b"]), f"{tag}: stripB frame {f + 1}"
            assert bits_equal(got["strips"][f + 1, ..., 1].astype(np.float32), w["strip_w"]), f"{tag}: stripW frame {f + 1}"
            assert bits_equal(got["delta_p"][f], w["delta_p"]), f"{tag}: deltaP frame {f + 1}"
            assert bits_equal(got["proj_u"][f], w["proj_u"]), f"{tag}: ProjectorU frame {f + 1}"
            assert bits_equal(got["mask"][f], w["mask"]), f"{tag}: mask frame {f + 1}"
            for ch, key in enumerate("xyz"):
                assert np.abs(got["xyzw"][f, ..., ch] - w[key]).max() <= tol, (tag, f, key)
            assert np.abs(got["delta_z"][f] - w["delta_z"]).max() <= 2 * tol, f"{tag}: deltaZ frame {f + 1}"


def test_dynamic_frames_long_sequence(built_library, oracle, base_calibration):
    """40 frames through one launch of the frame-walking kernel: the accumulated ProjectorU and the
    masks stay bit-exact to the last frame (the plane oscillates, so U returns and leaves again)."""
    from structured_light_calculation_b200 import synth
    from structured_light_calculation_b200.configs import StackConfig
    W, H, n_frames = 136, 72, 40
    cfg = StackConfig(W, H, 1280, 6, 4)
    cal, scene, planes = make_case(cfg, base_calibration, noise=1.0, seed=9)
    first = oracle_run(oracle, cfg, cal, planes)
    pool = synth.render_dyna_frames(cfg, cal, 6, stripe_period=14.0, z_step=0.4, noise_sigma=1.5)
    frames = np.stack([pool[k if k < 6 else 10 - k] for k in (f % 10 for f in range(n_frames))])
    ocfg = oracle.make_config(W, H, cfg.projector_width, cfg.gray_digits, cfg.phase_steps)
    ocal = oracle.make_calib(cal.cam, cal.pro, cal.R, cal.T)
    want = oracle.dyna_sequence(ocfg, ocal, first["proj_u"], first["z"], frames, 21)
    rec = _reconstructor(cfg, cal)
    got = rec.dyna_track(frames, first["proj_u"], window=21, parity=True)
    rec.close()
    tol = XYZ_REL_TOL * (cfg.fov_max - cfg.fov_min)
    for f, w in enumerate(want):
        assert bits_equal(got["proj_u"][f], w["proj_u"]), f"ProjectorU frame {f + 1}"
        assert bits_equal(got["delta_p"][f], w["delta_p"]), f"deltaP frame {f + 1}"
        assert bits_equal(got["mask"][f], w["mask"]), f"mask frame {f + 1}"
        assert np.abs(got["xyzw"][f, ..., 2] - w["z"]).max() <= tol
        assert np.abs(got["delta_z"][f] - w["delta_z"]).max() <= 2 * tol
    assert any((w["delta_p"] != 0).any() for w in want)


def test_smoke_entry_point(built_library):
    import __graft_entry__
    __graft_entry__.smoke()


def test_device_entry_points_stay_inside_their_buffers(built_library, oracle, base_calibration, tmp_path):
    """compute-sanitizer is not available on the GPU pool, so every device entry point that writes a
    caller-owned buffer is run with canary bytes after (and before) the region it may touch."""
    import torch
    from structured_light_calculation_b200 import capi, synth
    from structured_light_calculation_b200.capi import SlcDynaParity
    from structured_light_calculation_b200.configs import StackConfig
    W, H, n_frames = 200, 75, 5
    cfg = StackConfig(W, H, 1280, 7, 4)
    cal, scene, planes = make_case(cfg, base_calibration, noise=1.0, seed=3)
    frames = synth.render_dyna_frames(cfg, cal, n_frames, stripe_period=14.0, z_step=0.4, noise_sigma=1.5)
    npx, PAD, CAN = W * H, 4096, 0xA5
    dev = torch.device("cuda", 0)

    def guarded(nbytes):
        t = torch.full((nbytes + 2 * PAD,), CAN, dtype=torch.uint8, device=dev)
        return t, t.data_ptr() + PAD

    def intact(t, nbytes):
        return bool((t[:PAD] == CAN).all()) and bool((t[PAD + nbytes:] == CAN).all())

    rec = _reconstructor(cfg, cal)
    d_stack = torch.from_numpy(planes).to(dev)
    bufs = {"xyzw": 16 * npx, "mask": npx, "kbin": 2 * npx, "corr": npx, "pix": 4 * npx, "u": 8 * npx}
    g = {k: guarded(n) for k, n in bufs.items()}
    par = capi.SlcParityPlanes(g["kbin"][1], g["corr"][1], g["pix"][1], g["u"][1])
    rec.reconstruct_device(d_stack.data_ptr(), 1, g["xyzw"][1], g["mask"][1], par)
    rec.synchronize()
    assert all(intact(g[k][0], n) for k, n in bufs.items())
    d_u0 = g["u"][0][PAD:PAD + 8 * npx].clone()

    d_frames = torch.from_numpy(frames).to(dev)
    no = n_frames - 1
    dbufs = {"xyzw": 16 * npx * no, "mask": npx * no, "dz": 4 * npx * no, "strips": 2 * npx * n_frames,
             "dp": 4 * npx * no, "pu": 8 * npx * no}
    dg = {k: guarded(n) for k, n in dbufs.items()}
    dpar = SlcDynaParity(dg["strips"][1], dg["dp"][1], dg["pu"][1])
    rec._check(rec.lib.slc_dyna_track_device(rec.h, d_frames.data_ptr(), n_frames, 21, d_u0.data_ptr(), dg["xyzw"][1],
                                             dg["mask"][1], dg["dz"][1], dpar, None))
    rec.synchronize()
    assert all(intact(dg[k][0], n) for k, n in dbufs.items())

    # text cloud: exactly `bytes` bytes written, nothing after them
    cap = 43 * npx + 16
    tt, tptr = guarded(cap)
    nb, npts = rec.pointcloud_text_device(g["u"][1], tptr, cap)
    assert npts > 0 and intact(tt, cap) and bool((tt[PAD + nb:PAD + cap] == CAN).all())
    ct, cptr = guarded(12 * npx)
    n = rec.pointcloud_compact_device(g["xyzw"][1], g["mask"][1], cptr, npx, capi.SLC_ORDER_REFERENCE)
    assert n == npts and intact(ct, 12 * npx) and bool((ct[PAD + 12 * n:PAD + 12 * npx] == CAN).all())

    # BMP unpack: odd width, padded rows
    img = np.random.default_rng(1).integers(0, 256, (37, 53)).astype(np.uint8)
    for bpp in (8, 24):
        data = synth.encode_bmp(img, bpp)
        info = capi.bmp_parse(data)
        raw = torch.frombuffer(bytearray(data[info.pixel_offset:]), dtype=torch.uint8).to(dev)
        pt, pptr = guarded(img.size)
        rec._check(rec.lib.slc_bmp_unpack_device(rec.h, raw.data_ptr(), info, pptr, None))
        rec.synchronize()
        assert intact(pt, img.size)
        assert np.array_equal(pt[PAD:PAD + img.size].cpu().numpy().reshape(img.shape), img)
    rec.close()
