"""The reference's class API (CDecodeGray / CDecodePhase / CCalculation) re-hosted in C++
on the C ABI: compile a reference-style program against include/dynaframe_b200.hpp and run it."""
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, bits_equal, make_case, oracle_run

SRC = os.path.join(ROOT, "tests", "cpp", "host_api_test.cpp")


def _compile(tmp_path, built_library):
    from structured_light_calculation_b200 import capi
    exe = str(tmp_path / "host_api_test")
    libdir = os.path.dirname(capi.LIB_PATH)
    cmd = ["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe,
           "-L", libdir, "-lslcalc_b200", f"-Wl,-rpath,{libdir}"]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 0, res.stdout
    return exe


def test_cpp_program_compiles_and_links(tmp_path, built_library):
    exe = _compile(tmp_path, built_library)
    assert os.path.exists(exe)
    res = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 2      # usage


@pytest.mark.gpu
@pytest.mark.parametrize("G,N,PW,W,H", [(6, 4, 1280, 160, 64), (9, 4, 2560, 96, 40)])
def test_cpp_host_api_matches_oracle(tmp_path, built_library, oracle, base_calibration, G, N, PW, W, H):
    from structured_light_calculation_b200.configs import StackConfig
    exe = _compile(tmp_path, built_library)
    cfg = StackConfig(W, H, PW, G, N)
    cal, scene, planes = make_case(cfg, base_calibration, noise=1.0, seed=31)
    d = tmp_path / "data"
    d.mkdir()
    planes.tofile(d / "planes.u8")
    # the calibration in the exact format CCalculation::Init reads (OpenCV YAML, dt: d)
    def mat(name, m):
        m = np.asarray(m, dtype=np.float64)
        rows, cols = (m.shape if m.ndim == 2 else (m.size, 1))
        vals = ", ".join(f"{v:.17e}" for v in m.reshape(-1))
        return f"{name}: !!opencv-matrix\n   rows: {rows}\n   cols: {cols}\n   dt: d\n   data: [ {vals} ]\n"
    (d / "parameters.yml").write_text("%YAML:1.0\n" + mat("CamMat", cal.cam) + mat("ProMat", cal.pro) +
                                      mat("R", cal.R) + mat("T", cal.T))
    # the Gray code table in the format of Patterns/vGrayCode.txt: "bin gray" rows
    with open(d / "vGrayCode.txt", "w") as f:
        for b in range(1 << G):
            f.write(f"{b} {b ^ (b >> 1)}\n")
    from structured_light_calculation_b200 import synth
    frames = synth.render_dyna_frames(cfg, cal, 4, stripe_period=14.0, z_step=0.4, noise_sigma=1.5)
    frames.tofile(d / "dyna.u8")
    synth.write_reference_layout(str(d / "group"), cfg, planes, dyna_frames=frames, bpp=8 if G == 6 else 24)
    res = subprocess.run([exe, str(d), str(W), str(H), str(PW), str(G), str(N)], stdout=subprocess.PIPE,
                         stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 0, res.stdout
    assert "host_api_test ok" in res.stdout and "file-backed sensor ok" in res.stdout and "pool ok" in res.stdout
    want = oracle_run(oracle, cfg, cal, planes)
    gray = np.fromfile(d / "gray.f64", np.float64).reshape(H, W)
    phase = np.fromfile(d / "phase.f64", np.float64).reshape(H, W)
    xyzw = np.fromfile(d / "xyzw.f32", np.float32).reshape(H, W, 4)
    mask = np.fromfile(d / "mask.u8", np.uint8).reshape(H, W)
    projU = np.fromfile(d / "projU.f64", np.float64).reshape(H, W)
    assert bits_equal(gray, want["gray_val"]) and bits_equal(phase, want["phase_pix"])
    assert bits_equal(projU, want["proj_u"]) and bits_equal(mask, want["mask"])
    tol = 1e-5 * (cfg.fov_max - cfg.fov_min)
    for ch, key in enumerate("xyz"):
        assert np.abs(xyzw[..., ch] - want[key]).max() <= tol
    # Result(): "x y z" per in-FOV pixel, u outer / v inner (CCalculation.cpp:336-352)
    cloud = np.loadtxt(d / "cloud.txt").reshape(-1, 3)
    vv, uu = np.nonzero(want["mask"].T)[::-1][0], np.nonzero(want["mask"].T)[0]
    assert cloud.shape[0] == int(want["mask"].sum())
    ref = np.stack([want["x"][vv, uu], want["y"][vv, uu], want["z"][vv, uu]], axis=1)
    assert np.allclose(cloud, ref, rtol=2e-5, atol=1e-4)
    # ... and byte for byte what the reference's `file << double` loop writes from its f64 planes
    ocfg0 = oracle.make_config(W, H, PW, G, N)
    text, npts = oracle.result_text(ocfg0, want["x"], want["y"], want["z"])
    assert (d / "cloud.txt").read_bytes() == text and npts == cloud.shape[0]
    ply = (d / "cloud.ply").read_bytes()
    head, body = ply.split(b"end_header\n", 1)
    assert f"element vertex {npts}".encode() in head
    pts = np.frombuffer(body, np.float32).reshape(-1, 3)
    assert pts.shape[0] == npts and np.array_equal(pts, xyzw[vv, uu, :3])
    # CalculateOther(): dynamic frames through the class API
    ocfg = oracle.make_config(W, H, PW, G, N)
    seq = oracle.dyna_sequence(ocfg, oracle.make_calib(cal.cam, cal.pro, cal.R, cal.T), want["proj_u"], want["z"], frames)
    dxyzw = np.fromfile(d / "dyna_xyzw.f32", np.float32).reshape(3, H, W, 4)
    dmask = np.fromfile(d / "dyna_mask.u8", np.uint8).reshape(3, H, W)
    ddz = np.fromfile(d / "dyna_dz.f32", np.float32).reshape(3, H, W)
    for f, r in enumerate(seq):
        assert bits_equal(dmask[f], r["mask"])
        assert bits_equal(dxyzw[f, ..., 3], r["proj_u"].astype(np.float32))
        assert np.abs(dxyzw[f, ..., 2] - r["z"]).max() <= tol
        assert np.abs(ddz[f] - r["delta_z"]).max() <= 2 * tol
    assert np.loadtxt(d / "cloud_dyn1.txt").reshape(-1, 3).shape[0] == int(seq[0]["mask"].sum())
    text1, _ = oracle.result_text(ocfg, seq[0]["x"], seq[0]["y"], seq[0]["z"])
    assert (d / "cloud_dyn1.txt").read_bytes() == text1
