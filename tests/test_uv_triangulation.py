"""[EXT] SURVEY 8(f) rank 4: both projector coordinates.  The reference only has the m_vertical
switch of its Gray decoder (CDecodeGray.cpp:182-185) and never uses row 1 of P, so there is no
reference output to match: the definition (least squares over the two per-coordinate equations,
oracle/sl_oracle.c) is checked against ground truth on CPU, and the CUDA path against the oracle
bit for bit."""
import numpy as np
import pytest

from conftest import bits_equal, oracle_run


def _setup(base_calibration, W=192, H=128, noise=1.0):
    from structured_light_calculation_b200 import synth
    from structured_light_calculation_b200.configs import StackConfig
    cfg_v = StackConfig(W, H, 1280, 6, 4)            # columns: GRAY_V_NUMDIGIT = 6 over PROJECTOR_RESLINE = 1280
    cfg_h = StackConfig(W, H, 800, 5, 4)             # rows:    GRAY_H_NUMDIGIT = 5 over PROJECTOR_RESROW  = 800
    cal = synth.synthetic_calibration(cfg_v, base_calibration)
    scene = synth.make_scene(cfg_v, cal)
    stack_v = synth.render_stack(cfg_v, scene, noise_sigma=noise, seed=81)
    stack_h = synth.render_stack(cfg_h, scene, noise_sigma=noise, seed=82, horizontal=True)
    return cfg_v, cfg_h, cal, scene, stack_v, stack_h


def test_uv_definition_recovers_ground_truth(oracle, base_calibration):
    """Exact U and V of the scene -> the scene's z, to rounding; decoded U, V -> z to quantisation."""
    cfg_v, cfg_h, cal, scene, stack_v, stack_h = _setup(base_calibration, noise=0.0)
    ocfg = oracle.make_config(cfg_v.width, cfg_v.height, cfg_v.projector_width, cfg_v.gray_digits, cfg_v.phase_steps)
    ocal = oracle.make_calib(cal.cam, cal.pro, cal.R, cal.T)
    exact = oracle.triangulate_uv(ocfg, ocal, scene.U, scene.V)
    inside = (scene.z >= cfg_v.fov_min) & (scene.z <= cfg_v.fov_max)
    assert np.array_equal(exact["mask"].astype(bool), inside)
    assert np.abs(exact["z"] - scene.z)[inside].max() < 1e-9
    assert np.abs(exact["x"] - scene.xyz[..., 0])[inside].max() < 1e-9
    # decoded coordinates (u8 quantisation only)
    U = oracle_run(oracle, cfg_v, cal, stack_v)["proj_u"]
    V = oracle_run(oracle, cfg_h, cal, stack_h)["proj_u"]
    lit_h = scene.lit & (scene.V >= 0) & (scene.V < cfg_h.projector_width)
    good = inside & scene.lit & lit_h & (scene.albedo > 0.1)
    assert good.mean() > 0.5
    assert np.abs(V - scene.V)[good].max() < 0.2 and np.abs(U - scene.U)[good].max() < 0.2
    both = oracle.triangulate_uv(ocfg, ocal, U, V)
    single = oracle_run(oracle, cfg_v, cal, stack_v)
    err_both = np.abs(both["z"] - scene.z)[good & (both["mask"] > 0)]
    err_single = np.abs(single["z"] - scene.z)[good & (single["mask"] > 0)]
    assert err_both.max() < 0.15 and err_both.mean() <= 1.05 * err_single.mean()      # dz/dU ~ 0.12 units per projector px


@pytest.mark.gpu
def test_uv_triangulation_matches_oracle(built_library, oracle, base_calibration):
    from structured_light_calculation_b200 import capi
    cfg_v, cfg_h, cal, scene, stack_v, stack_h = _setup(base_calibration)
    rec_v = capi.Reconstructor(cfg_v, device=0, max_batch=1, num_slots=1)
    rec_v.set_calibration(cal)
    rec_h = capi.Reconstructor(cfg_h, device=0, max_batch=1, num_slots=1)
    rec_h.set_calibration(cal)
    U = rec_v.reconstruct(stack_v, parity=True)["proj_u"][0]
    V = rec_h.reconstruct(stack_h, parity=True)["proj_u"][0]          # horizontal patterns through the same decode
    assert bits_equal(U, oracle_run(oracle, cfg_v, cal, stack_v)["proj_u"])
    assert bits_equal(V, oracle_run(oracle, cfg_h, cal, stack_h)["proj_u"])
    xyzw, mask = rec_v.triangulate_uv(U, V)
    ocfg = oracle.make_config(cfg_v.width, cfg_v.height, cfg_v.projector_width, cfg_v.gray_digits, cfg_v.phase_steps)
    want = oracle.triangulate_uv(ocfg, oracle.make_calib(cal.cam, cal.pro, cal.R, cal.T), U, V)
    assert bits_equal(mask, want["mask"]) and mask.mean() > 0.4
    for ch, k in enumerate("xyz"):
        assert bits_equal(xyzw[..., ch], want[k].astype(np.float32)), k
    assert bits_equal(xyzw[..., 3], U.astype(np.float32))
    # a pixel with only one coordinate decoded is invalid
    V2 = V.copy()
    V2[10:20, 30:60] = 0.0
    _, mask2 = rec_v.triangulate_uv(U, V2)
    assert not mask2[10:20, 30:60].any()
    with pytest.raises(capi.SlcError):
        rec_h._check(rec_h.lib.slc_triangulate_uv_host(rec_h.h, None, None, None, None))
    rec_v.close()
    rec_h.close()
