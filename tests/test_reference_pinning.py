"""The oracle and the CUDA path against the REFERENCE ITSELF.

oracle/_ref/dynaframe_ref is the reference's own path sources (CDecodeGray.cpp, CDecodePhase.cpp,
CCalculation.cpp, CSensorV.cpp, GlobalFunction.cpp) compiled in place from /root/reference by
oracle/Makefile against a minimal OpenCV stand-in (oracle/ref_shim/).  tests/golden/
reference_run_g6n4.npz holds inputs and outputs of one run of it (made by make_golden.py), so the
pin exists even where the binary does not."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, bits_equal, make_case, oracle_run

XYZ_REL_TOL = 1e-5


def _fixture():
    return np.load(os.path.join(GOLDEN, "reference_run_g6n4.npz"))


def _cfg_cal(z):
    from structured_light_calculation_b200.calibration import Calibration
    from structured_light_calculation_b200.configs import StackConfig
    H, W = z["planes"].shape[1:]
    return StackConfig(W, H, 1280, 6, 4), Calibration(z["cam"], z["pro"], z["R"], z["T"])


def test_oracle_matches_committed_reference_run(oracle):
    """Every plane the reference computed -- decoders, ProjectorU, x, y, z, strips, deltaP, deltaZ --
    and the text clouds it wrote, bit for bit / byte for byte."""
    z = _fixture()
    cfg, cal = _cfg_cal(z)
    ocfg = oracle.make_config(cfg.width, cfg.height, cfg.projector_width, cfg.gray_digits, cfg.phase_steps)
    ocal = oracle.make_calib(cal.cam, cal.pro, cal.R, cal.T)
    A, B, cC, cD, P = oracle.calibration(ocfg, ocal)
    assert A == float(z["A"]) and B == float(z["B"]) and bits_equal(P, z["P"])
    assert bits_equal(cC, z["cC"]) and bits_equal(cD, z["cD"])
    want = oracle.reconstruct(ocfg, ocal, z["planes"])
    assert bits_equal(want["gray_val"], z["gray"]) and bits_equal(want["phase_pix"], z["phase"])
    assert bits_equal(want["proj_u"], z["f0_projU"])
    for k in "xyz":
        assert bits_equal(want[k], z[f"f0_{k}"]), k
    text, _ = oracle.result_text(ocfg, want["x"], want["y"], want["z"])
    assert text == z["cloud0"].tobytes()
    frames = z["dyna_frames"]
    B0, W0 = oracle.strip_regression(ocfg, frames[0], 21)
    assert bits_equal(B0, z["f0_stripB"]) and bits_equal(W0, z["f0_stripW"])
    seq = oracle.dyna_sequence(ocfg, ocal, want["proj_u"], want["z"], frames, 21)
    for f, w in enumerate(seq, start=1):
        for a, b in (("stripB", "strip_b"), ("stripW", "strip_w"), ("deltaP", "delta_p"), ("projU", "proj_u"),
                     ("x", "x"), ("y", "y"), ("z", "z"), ("deltaZ", "delta_z")):
            assert bits_equal(w[b], z[f"f{f}_{a}"]), (f, a)
        text, _ = oracle.result_text(ocfg, w["x"], w["y"], w["z"])
        assert text == z[f"cloud{f}"].tobytes(), f


@pytest.fixture(scope="module")
def ref():
    from oracle import ref_runner
    if not ref_runner.available():
        if os.path.isdir("/root/reference"):
            ref_runner.build()
        if not ref_runner.available():
            pytest.skip("oracle/_ref/dynaframe_ref not built (needs /root/reference)")
    return ref_runner


def test_reference_binary_reproduces_committed_run(ref):
    z = _fixture()
    cfg, cal = _cfg_cal(z)
    ws = ref.Workspace(cfg, cal, z["planes"], z["dyna_frames"])
    full = ws.run_full()
    ws.close()
    for f, fr in enumerate(full["frames"]):
        for k, v in fr.items():
            assert bits_equal(v, z[f"f{f}_{k}"]), (f, k)
        assert full["clouds"][f] == z[f"cloud{f}"].tobytes()


def test_reference_gray_table_is_the_shipped_file(ref):
    """The runner generates Patterns/vGrayCode.txt; for the reference's own digit count it must be the
    file the reference ships (only checkable where /root/reference exists)."""
    path = "/root/reference/DynaFrame/DynaFrame/Patterns/vGrayCode.txt"
    if not os.path.exists(path):
        pytest.skip("reference tree absent")
    shipped = np.loadtxt(path, dtype=np.int64)
    assert np.array_equal(shipped, np.array([[b, b ^ (b >> 1)] for b in range(64)]))


@pytest.mark.parametrize("W,H,PW,G", [(160, 128, 1280, 6), (200, 75, 1280, 7), (96, 80, 2560, 9), (64, 40, 2048, 8)])
def test_oracle_matches_reference_binary(ref, oracle, base_calibration, W, H, PW, G):
    """The same loops at other geometries (ref_params.cpp takes them from the environment): the
    oracle restatement against the reference's compiled sources, first frame, all f64 planes."""
    from structured_light_calculation_b200.configs import StackConfig
    cfg = StackConfig(W, H, PW, G, 4)
    cal, scene, planes = make_case(cfg, base_calibration, noise=1.5, seed=G)
    ws = ref.Workspace(cfg, cal, planes)
    got = ws.run_first()
    ws.close()
    want = oracle_run(oracle, cfg, cal, planes)
    for a, b in (("gray", "gray_val"), ("phase", "phase_pix"), ("projU", "proj_u"), ("x", "x"), ("y", "y"), ("z", "z")):
        assert bits_equal(got[a], want[b]), a
    assert want["mask"].mean() > 0.3


@pytest.mark.gpu
@pytest.mark.parametrize("W,H,PW,G", [(192, 128, 1280, 6), (200, 75, 2560, 9)])
def test_cuda_first_frame_matches_reference_binary(ref, built_library, base_calibration, W, H, PW, G):
    """The fused kernel through the C ABI against the reference's own compiled loops."""
    from structured_light_calculation_b200 import capi
    from structured_light_calculation_b200.configs import StackConfig
    cfg = StackConfig(W, H, PW, G, 4)
    cal, scene, planes = make_case(cfg, base_calibration, noise=1.0, seed=50 + G)
    ws = ref.Workspace(cfg, cal, planes)
    want = ws.run_first()
    ws.close()
    rec = capi.Reconstructor(cfg, device=0, max_batch=1, num_slots=1)
    rec.set_calibration(cal)
    got = rec.reconstruct(planes, parity=True)
    rec.close()
    gp = PW // (1 << G)
    assert bits_equal(got["kbin"][0].astype(np.float64) * gp, want["gray"])          # period index
    assert bits_equal(got["phase_pix"][0].astype(np.float64), want["phase"])
    assert bits_equal(got["proj_u"][0], want["projU"])
    ref_mask = ((want["projU"] != 0) & (want["z"] >= cfg.fov_min) & (want["z"] <= cfg.fov_max)).astype(np.uint8)
    assert bits_equal(got["mask"][0], ref_mask)
    tol = XYZ_REL_TOL * (cfg.fov_max - cfg.fov_min)
    for ch, k in enumerate("xyz"):
        assert np.abs(got["xyzw"][0, ..., ch] - want[k]).max() <= tol, k


@pytest.mark.gpu
def test_cuda_dynamic_frames_and_clouds_match_reference_binary(ref, built_library, base_calibration):
    """CalculateFirst + CalculateOther of the reference (its own text clouds included) against
    slc_reconstruct_host + slc_dyna_track_host + slc_pointcloud_text_host."""
    from structured_light_calculation_b200 import capi, synth
    from structured_light_calculation_b200.configs import StackConfig
    cfg = StackConfig(160, 96, 1280, 6, 4)
    cal, scene, planes = make_case(cfg, base_calibration, noise=1.0, seed=61)
    frames = synth.render_dyna_frames(cfg, cal, 5, stripe_period=14.0, z_step=0.4, noise_sigma=1.5)
    ws = ref.Workspace(cfg, cal, planes, frames)
    want = ws.run_full()
    ws.close()
    rec = capi.Reconstructor(cfg, device=0, max_batch=1, num_slots=1)
    rec.set_calibration(cal)
    first = rec.reconstruct(planes, parity=True)
    assert bits_equal(first["proj_u"][0], want["frames"][0]["projU"])
    assert rec.pointcloud_text(first["proj_u"][0])[0] == want["clouds"][0]
    dyn = rec.dyna_track(frames, first["proj_u"][0], window=21, parity=True)
    tol = XYZ_REL_TOL * (cfg.fov_max - cfg.fov_min)
    for f in range(1, 5):
        fr = want["frames"][f]
        assert bits_equal(dyn["strips"][f, ..., 0].astype(np.float32), fr["stripB"])
        assert bits_equal(dyn["strips"][f, ..., 1].astype(np.float32), fr["stripW"])
        assert bits_equal(dyn["delta_p"][f - 1], fr["deltaP"])
        assert bits_equal(dyn["proj_u"][f - 1], fr["projU"])
        assert np.abs(dyn["xyzw"][f - 1, ..., 2] - fr["z"]).max() <= tol
        assert np.abs(dyn["delta_z"][f - 1] - fr["deltaZ"]).max() <= 2 * tol
        assert rec.pointcloud_text(dyn["proj_u"][f - 1])[0] == want["clouds"][f]
    rec.close()


@pytest.mark.gpu
def test_whole_program_writes_the_reference_programs_bytes(ref, built_library, base_calibration, tmp_path):
    """examples/dynaframe_main.cpp (the reference main() on the library, file-backed sensor) and the
    reference's own program on the same files: every text cloud byte for byte."""
    import subprocess
    from structured_light_calculation_b200 import synth
    from structured_light_calculation_b200.configs import StackConfig
    from conftest import ROOT
    exe = os.path.join(ROOT, "structured_light_calculation_b200", "bin", "dynaframe_main")
    if not os.path.exists(exe):
        import __graft_entry__
        __graft_entry__.build()
    cfg = StackConfig(200, 96, 1280, 6, 4)
    cal, scene, planes = make_case(cfg, base_calibration, noise=1.0, seed=71)
    n = 5
    frames = synth.render_dyna_frames(cfg, cal, n, stripe_period=14.0, z_step=0.4, noise_sigma=1.5)
    ws = ref.Workspace(cfg, cal, planes, frames, root=str(tmp_path))
    want = ws.run_app()
    out = tmp_path / "ours"
    out.mkdir()
    res = subprocess.run([exe, ws.data, str(cfg.width), str(cfg.height), str(cfg.projector_width), str(cfg.gray_digits),
                          str(cfg.phase_steps), str(n), str(out)], cwd=ws.cwd, stdout=subprocess.PIPE,
                         stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 0, res.stdout
    for f in range(n):
        name = "iFrame.txt" if f == 0 else f"cFrame{f}.txt"
        assert (out / name).read_bytes() == want["clouds"][f], name
    assert len(want["clouds"][0]) > 1000


@pytest.mark.gpu
def test_baseline_config0_full_size_against_reference_binary(ref, built_library, base_calibration):
    """BASELINE configs[0] as stated: one 1280x1024 stack, 7-bit Gray + 4-step, decoded and triangulated
    by the reference's own CPU path -- here its compiled sources -- against the fused kernel."""
    from structured_light_calculation_b200 import capi
    from structured_light_calculation_b200.configs import CONFIGS
    cfg = CONFIGS["config1"]
    cal, scene, planes = make_case(cfg, base_calibration, noise=1.0, seed=7)
    ws = ref.Workspace(cfg, cal, planes)
    want = ws.run_first()
    ws.close()
    rec = capi.Reconstructor(cfg, device=0, max_batch=1, num_slots=1)
    rec.set_calibration(cal)
    got = rec.reconstruct(planes, parity=True)
    rec.close()
    assert bits_equal(got["kbin"][0].astype(np.float64) * cfg.gray_period, want["gray"])
    assert bits_equal(got["phase_pix"][0].astype(np.float64), want["phase"])
    assert bits_equal(got["proj_u"][0], want["projU"])
    ref_mask = ((want["projU"] != 0) & (want["z"] >= cfg.fov_min) & (want["z"] <= cfg.fov_max)).astype(np.uint8)
    assert bits_equal(got["mask"][0], ref_mask) and ref_mask.mean() > 0.8
    tol = XYZ_REL_TOL * (cfg.fov_max - cfg.fov_min)
    errs = [float(np.abs(got["xyzw"][0, ..., ch] - want[k]).max()) for ch, k in enumerate("xyz")]
    print("configs[0] max |dx|, |dy|, |dz| vs the reference binary:", errs)
    assert max(errs) <= tol
