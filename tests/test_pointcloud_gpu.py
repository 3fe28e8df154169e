"""Point-cloud output: CCalculation::Result (CCalculation.cpp:323-357) on the
device -- number formatting against printf, the whole text against the oracle's restatement of
the reference loop byte for byte, and the binary compaction against numpy."""
from fractions import Fraction

import numpy as np
import pytest

from conftest import make_case, oracle_run

pytestmark = pytest.mark.gpu


def _rec(cfg, cal, **kw):
    from structured_light_calculation_b200 import capi
    rec = capi.Reconstructor(cfg, device=0, max_batch=1, num_slots=1, **kw)
    rec.set_calibration(cal)
    return rec


def _exact_ties():
    """doubles whose 7th significant digit is an exact 5: |x| * 10^k = D + 0.5 exactly"""
    out = []
    rng = np.random.default_rng(3)
    for k in range(-8, 9):
        for _ in range(40):
            D = int(rng.integers(100000, 1000000))
            fr = Fraction(2 * D + 1, 2) / Fraction(10) ** k
            x = float(fr)
            if Fraction(x) == fr:
                out.append(x)
    # k > 0 needs 5^k | (2D+1)
    for k in range(1, 8):
        step = 5 ** k
        for m in range(1, 400, 2):
            n = m * step
            if n % 2 == 1 and 200001 <= n <= 1999999:
                fr = Fraction(n, 2) / Fraction(10) ** k
                x = float(fr)
                if Fraction(x) == fr:
                    out.append(x)
    return np.array(out)


def test_format_g6_matches_printf(built_library, oracle, base_calibration):
    from structured_light_calculation_b200 import capi
    from structured_light_calculation_b200.configs import StackConfig
    cfg = StackConfig(64, 32, 1280, 6, 4)
    cal, _, _ = make_case(cfg, base_calibration)
    rec = _rec(cfg, cal)
    rng = np.random.default_rng(11)
    ties = _exact_ties()
    assert ties.size > 300
    vals = np.concatenate([
        np.array([0.0, -0.0, 1.0, -1.0, 0.1, 0.5, 1.5, 2.5, 100.0, 1e5, 99999.95, 999999.5, 999999.4999999999, 1e6,
                  123456.5, 1234565.0, 1e-4, 9.9999995e-5, 9.999995e-5, 0.00099999999999999, 0.001, 1e-5, 1e22,
                  1.5e-300, 5e-324, 2.2250738585072014e-308, 1.7976931348623157e308, 26.684693078303,
                  -23.263796885137, 0.010991564619, 88.385787100032, float("inf"), -float("inf")]),
        ties, -ties,
        rng.uniform(-120, 120, 200000),                                   # coordinates
        rng.uniform(-1, 1, 50000) * 10.0 ** rng.integers(-12, 12, 50000),
        rng.uniform(1, 10, 50000) * 10.0 ** rng.integers(-300, 300, 50000),
        np.round(rng.uniform(-100, 100, 50000), 4),                       # short decimals
        np.nextafter(10.0 ** np.arange(-20, 21), 0), 10.0 ** np.arange(-20, 21), np.nextafter(10.0 ** np.arange(-20, 21), np.inf),
    ])
    for flags in (0, capi.SLC_TEXT_EXP3):
        got = rec.format_g6(vals, flags)
        bad = []
        for v, g in zip(vals, got):
            want = oracle.format_g6(v, flags)
            exact_domain = (v == 0) or (1e-17 <= abs(v) < 1e28) or not np.isfinite(v)
            if g != want and exact_domain:
                bad.append((v, g, want))
            if flags == 0 and np.isfinite(v):
                assert want == (b"%g" % v)                                # the oracle is printf
        assert not bad, bad[:10]
        # outside the exact domain (not a coordinate) only the last digit may differ
        for v, g in zip(vals, got):
            if np.isfinite(v) and v != 0 and not (1e-17 <= abs(v) < 1e28):
                assert abs(float(g) - v) <= 1.01e-5 * abs(v), (v, g)
    rec.close()


@pytest.mark.parametrize("W,H,flags", [(192, 128, 0), (200, 75, 3), (64, 40, 1)])
def test_pointcloud_text_matches_oracle(built_library, oracle, base_calibration, W, H, flags):
    from structured_light_calculation_b200.configs import StackConfig
    cfg = StackConfig(W, H, 1280, 7, 4)
    cal, scene, planes = make_case(cfg, base_calibration, noise=1.0, seed=21)
    want = oracle_run(oracle, cfg, cal, planes)
    ocfg = oracle.make_config(W, H, cfg.projector_width, cfg.gray_digits, cfg.phase_steps)
    text, npts = oracle.result_text(ocfg, want["x"], want["y"], want["z"], flags)
    assert npts == int(want["mask"].sum()) and npts > 0.3 * W * H
    rec = _rec(cfg, cal)
    got = rec.reconstruct(planes, parity=True)
    gtext, gpts = rec.pointcloud_text(got["proj_u"][0], flags)
    assert gpts == npts
    assert gtext == text
    # the f64 plane from the oracle gives the same bytes
    assert rec.pointcloud_text(want["proj_u"], flags)[0] == text
    # empty cloud, and a buffer that is too small
    assert rec.pointcloud_text(np.zeros((H, W)), flags) == (b"", 0)
    from structured_light_calculation_b200.capi import SlcError
    with pytest.raises(SlcError) as ei:
        rec.pointcloud_text(got["proj_u"][0], flags, capacity=len(text) - 1)
    assert str(len(text)) in str(ei.value)
    assert rec.pointcloud_text(got["proj_u"][0], flags, capacity=len(text))[0] == text
    rec.close()


@pytest.mark.parametrize("W,H", [(192, 128), (200, 75)])
def test_pointcloud_compact_matches_numpy(built_library, oracle, base_calibration, W, H):
    from structured_light_calculation_b200 import capi
    from structured_light_calculation_b200.configs import StackConfig
    cfg = StackConfig(W, H, 1280, 7, 4)
    cal, scene, planes = make_case(cfg, base_calibration, noise=1.0, seed=22)
    rec = _rec(cfg, cal)
    got = rec.reconstruct(planes)
    xyzw, mask = got["xyzw"][0], got["mask"][0]
    rm = rec.pointcloud_compact(xyzw, mask, capi.SLC_ORDER_ROW_MAJOR)
    assert np.array_equal(rm, xyzw[mask != 0][:, :3])
    ref = rec.pointcloud_compact(xyzw, mask, capi.SLC_ORDER_REFERENCE)
    uu, vv = np.nonzero(mask.T)
    assert np.array_equal(ref, xyzw[vv, uu, :3])
    none = rec.pointcloud_compact(xyzw, np.zeros_like(mask))
    assert none.shape == (0, 3)
    every = rec.pointcloud_compact(xyzw, np.ones_like(mask), capi.SLC_ORDER_REFERENCE)
    assert np.array_equal(every, xyzw.transpose(1, 0, 2).reshape(-1, 4)[:, :3])
    rec.close()


def test_pointcloud_and_ingest_error_paths(built_library, base_calibration):
    """Bad arguments come back as status codes with a message, never as a crash."""
    import torch
    from structured_light_calculation_b200 import capi
    from structured_light_calculation_b200.capi import SlcError
    from structured_light_calculation_b200.configs import StackConfig
    cfg = StackConfig(64, 32, 1280, 6, 4)
    cal, _, planes = make_case(cfg, base_calibration)
    dev = torch.device("cuda", 0)
    rec = capi.Reconstructor(cfg, device=0, max_batch=1, num_slots=1)
    d_u = torch.zeros((32, 64), dtype=torch.float64, device=dev)
    d_text = torch.empty((43 * 2048 + 32,), dtype=torch.uint8, device=dev)
    with pytest.raises(SlcError) as ei:                       # calibration not set (CCalculation.cpp:176-181)
        rec.pointcloud_text_device(d_u.data_ptr(), d_text.data_ptr(), d_text.numel())
    assert ei.value.status == capi.SLC_ERR_NOT_INITIALISED
    rec.set_calibration(cal)
    with pytest.raises(SlcError) as ei:                       # output must be 16-byte aligned
        rec.pointcloud_text_device(d_u.data_ptr(), d_text.data_ptr() + 4, 1024)
    assert "aligned" in ei.value.message
    got = rec.reconstruct(planes)
    d_xyzw = torch.from_numpy(got["xyzw"][0]).to(dev)
    d_mask = torch.from_numpy(got["mask"][0]).to(dev)
    d_xyz = torch.empty((2048, 3), dtype=torch.float32, device=dev)
    with pytest.raises(SlcError) as ei:                       # unknown order
        rec.pointcloud_compact_device(d_xyzw.data_ptr(), d_mask.data_ptr(), d_xyz.data_ptr(), 2048, order=7)
    assert "order" in ei.value.message
    with pytest.raises(SlcError) as ei:                       # too small: the needed size is reported
        rec.pointcloud_compact_device(d_xyzw.data_ptr(), d_mask.data_ptr(), d_xyz.data_ptr(), 10)
    assert "needs" in ei.value.message
    with pytest.raises(SlcError):                             # NULL map
        rec.pointcloud_compact_device(0, d_mask.data_ptr(), d_xyz.data_ptr(), 2048)
    with pytest.raises(SlcError) as ei:                       # a path that does not exist names itself
        rec.load_bmp_planes(["/nonexistent/vGrayCam0.bmp"], d_mask.data_ptr())
    assert "vGrayCam0.bmp" in ei.value.message
    rec.load_bmp_planes([], d_mask.data_ptr())                # nothing to do
    # the context is still usable after every failure
    assert rec.pointcloud_compact_device(d_xyzw.data_ptr(), d_mask.data_ptr(), d_xyz.data_ptr(), 2048) == int(got["mask"].sum())
    rec.close()


def test_modulation_rejected_pixels_have_no_projector_column(built_library, oracle, base_calibration):
    """[EXT] modulation mask on: a rejected pixel carries U = 0 (the reference's own "no value" sentinel,
    CCalculation.cpp:678) in w and in the proj_u plane, so everything downstream that only tests U == 0 --
    Result()'s text cloud, FillCoordinate(i) on the plane -- agrees with the validity mask."""
    from structured_light_calculation_b200.configs import StackConfig
    cfg = StackConfig(272, 128, 2048, 8, 8, modulation_min=8.0)
    cal, scene, planes = make_case(cfg, base_calibration, noise=1.0, seed=23)
    want = oracle_run(oracle, cfg, cal, planes)
    rejected = want["mod_ok"] == 0
    assert rejected.any() and not rejected.all()
    assert not want["proj_u"][rejected].any()
    rec = _rec(cfg, cal)
    got = rec.reconstruct(planes, parity=True)
    assert np.array_equal(got["proj_u"][0], want["proj_u"]) and np.array_equal(got["mask"][0], want["mask"])
    assert not got["proj_u"][0][rejected].any() and not got["xyzw"][0][rejected].any()
    fast = rec.reconstruct(planes)                               # MODE 1 (no parity planes): same w
    assert np.array_equal(fast["xyzw"], got["xyzw"])
    text, npts = rec.pointcloud_text(got["proj_u"][0])
    assert npts == int(got["mask"][0].sum()) == text.count(b"\n")
    xyzw, mask = rec.triangulate(got["proj_u"][0])
    assert np.array_equal(mask, got["mask"][0])
    rec.close()
