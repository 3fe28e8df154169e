"""Input ingest: CSensor::LoadDatas' imread(path, CV_LOAD_IMAGE_GRAYSCALE) of
.bmp files (CSensorV.cpp:111-114).  The golden planes come from cv2.imread itself
(tests/golden/bmp_cases.npz, made by make_golden.py)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, bits_equal, make_case


def _cases():
    z = np.load(os.path.join(GOLDEN, "bmp_cases.npz"))
    return {k[:-6]: (z[k].tobytes(), z[k[:-6] + "__gray"]) for k in z.files if k.endswith("__file")}


def test_bmp_oracle_matches_cv2_fixture():
    from oracle.bmp_oracle import decode_bmp_gray
    cases = _cases()
    assert len(cases) >= 8
    for name, (data, want) in cases.items():
        assert np.array_equal(decode_bmp_gray(data), want), name


def test_bmp_header_parse(built_library):
    """slc_bmp_parse is metadata only (no pixel work): geometry, palette -> gray table, rejections."""
    from oracle.bmp_oracle import bgr_to_gray
    from structured_light_calculation_b200 import capi, synth
    for name, (data, want) in _cases().items():
        info = capi.bmp_parse(data)
        assert (info.height, info.width) == want.shape, name
        assert info.top_down == (1 if "topdown" in name else 0)
        assert info.row_stride == ((info.width * info.bits_per_pixel // 8 + 3) & ~3)
        if name.startswith("pal8_colour"):
            pal = np.frombuffer(data, np.uint8, count=1024, offset=54).reshape(256, 4)
            assert np.array_equal(np.array(info.gray[:]), bgr_to_gray(pal[:, 0], pal[:, 1], pal[:, 2]))
            assert info.palette_is_identity == 0
        if name.startswith("gray8"):
            assert info.palette_is_identity == 1
    good = synth.encode_bmp(np.zeros((4, 8), np.uint8))
    for bad in (b"", b"PNG" + good[3:], good[:60], good[:-1],                       # empty, magic, header only, truncated
                good[:30] + b"\x01\x00\x00\x00" + good[34:],                         # BI_RLE8
                good[:28] + b"\x04\x00" + good[30:]):                                # 4 bpp
        with pytest.raises(capi.SlcError):
            capi.bmp_parse(bad)


@pytest.mark.gpu
def test_bmp_decode_on_device_matches_cv2(built_library, base_calibration):
    from structured_light_calculation_b200 import capi
    from structured_light_calculation_b200.configs import StackConfig
    cfg = StackConfig(64, 32, 1280, 6, 4)
    cal, _, _ = make_case(cfg, base_calibration)
    rec = capi.Reconstructor(cfg, device=0, max_batch=1, num_slots=1)
    rec.set_calibration(cal)
    for name, (data, want) in _cases().items():
        got = rec.bmp_decode(data)
        assert np.array_equal(got, want), name
    with pytest.raises(capi.SlcError):
        rec.bmp_decode(_cases()["gray8_37x11"][0], expect_shape=(11, 38))
    rec.close()


@pytest.mark.gpu
@pytest.mark.parametrize("bpp,W,H", [(8, 192, 128), (24, 200, 75)])
def test_stack_from_reference_file_layout(tmp_path, built_library, oracle, base_calibration, bpp, W, H):
    """iFrame/vGrayCam{i}.bmp + vPhaseCam{i}.bmp -> device stack -> fused kernel == the in-memory run."""
    from structured_light_calculation_b200 import capi, synth
    from structured_light_calculation_b200.configs import StackConfig
    cfg = StackConfig(W, H, 1280, 7, 4)
    cal, scene, planes = make_case(cfg, base_calibration, noise=1.0, seed=41)
    paths = synth.write_reference_layout(str(tmp_path / "group"), cfg, planes, bpp=bpp)
    assert os.path.basename(paths[0]) == "vGrayCam0.bmp" and os.path.basename(paths[-1]) == "vPhaseCam3.bmp"
    rec = capi.Reconstructor(cfg, device=0, max_batch=1, num_slots=1)
    rec.set_calibration(cal)
    want = rec.reconstruct(planes)
    npx = W * H
    d_stack = rec.device_alloc(cfg.planes * npx)
    d_xyzw = rec.device_alloc(16 * npx)
    d_mask = rec.device_alloc(npx)
    rec.load_bmp_planes(paths, d_stack)
    back = np.empty_like(planes)
    rec.to_host(back, d_stack)
    assert np.array_equal(back, planes)
    rec.reconstruct_device(d_stack, 1, d_xyzw, d_mask)
    rec.synchronize()
    xyzw = np.empty((H, W, 4), np.float32)
    mask = np.empty((H, W), np.uint8)
    rec.to_host(xyzw, d_xyzw)
    rec.to_host(mask, d_mask)
    assert bits_equal(xyzw, want["xyzw"][0]) and bits_equal(mask, want["mask"][0])
    # a missing file names itself (CSensorV.cpp:122-129)
    with pytest.raises(capi.SlcError) as ei:
        rec.load_bmp_planes(paths[:2] + [str(tmp_path / "group" / "iFrame" / "vGrayCam99.bmp")], d_stack)
    assert "vGrayCam99.bmp" in str(ei.value)
    # wrong size
    other = tmp_path / "small.bmp"
    other.write_bytes(synth.encode_bmp(np.zeros((8, 8), np.uint8)))
    with pytest.raises(capi.SlcError):
        rec.load_bmp_planes([str(other)], d_stack)
    for p in (d_stack, d_xyzw, d_mask):
        rec.device_free(p)
    rec.close()


@pytest.mark.gpu
@pytest.mark.parametrize("W,H", [(64, 24), (200, 30)])       # wide (uint4) path / generic path
def test_bmp_batch_unpack_matches_decoder_oracle(built_library, base_calibration, W, H):
    """slc_bmp_unpack_batch_device: a set of files already in device memory -> the plane-major stack
    in one launch.  Covers every source misalignment (the pixel array of a .bmp starts 1078 bytes
    into the file), identity and colour palettes, 24-bit, top-down, pixel arrays that end exactly at
    the end of their allocation, and the bytes around the stack (canaries)."""
    import torch
    from oracle.bmp_oracle import decode_bmp_gray
    from structured_light_calculation_b200 import capi, synth
    from structured_light_calculation_b200.configs import StackConfig
    rng = np.random.default_rng(W * 1000 + H)
    cfg = StackConfig(W, H, 1280, 6, 4)
    cal, _, _ = make_case(cfg, base_calibration)
    rec = capi.Reconstructor(cfg, device=0, max_batch=1, num_slots=1)
    rec.set_calibration(cal)
    dev = torch.device("cuda", 0)
    files = []
    for k in range(9):
        img = rng.integers(0, 256, (H, W), dtype=np.uint8)
        data = bytearray(synth.encode_bmp(img, bpp=24 if k == 7 else 8, top_down=(k in (3, 8))))
        if k in (4, 5):                       # a colour palette: index -> gray through the fixed-point weights
            pal = rng.integers(0, 256, (256, 4), dtype=np.uint8)
            data[54:54 + 1024] = pal.tobytes()
        files.append(bytes(data))
    infos = [capi.bmp_parse(f) for f in files]
    want = np.stack([decode_bmp_gray(f) for f in files])
    keep, ptrs = [], []
    for k, (f, info) in enumerate(zip(files, infos)):
        raw = np.frombuffer(f, np.uint8)[info.pixel_offset:]
        mis = k % 4                           # the pixel array starts `mis` bytes into a tensor and ends with it
        t = torch.empty(mis + raw.size, dtype=torch.uint8, device=dev)
        t[mis:].copy_(torch.from_numpy(raw.copy()))
        keep.append(t)
        ptrs.append(t.data_ptr() + mis)
    PAD, CAN = 4096, 0x5A
    n, npx = len(files), W * H
    stack = torch.full((n * npx + 2 * PAD,), CAN, dtype=torch.uint8, device=dev)
    rec.bmp_unpack_batch_device(ptrs, infos, stack.data_ptr() + PAD)
    rec.synchronize()
    got = stack[PAD:PAD + n * npx].cpu().numpy().reshape(n, H, W)
    for k in range(n):
        assert np.array_equal(got[k], want[k]), f"file {k}"
    assert bool((stack[:PAD] == CAN).all()) and bool((stack[PAD + n * npx:] == CAN).all())
    # the single-plane entry point writes the same planes
    one = torch.empty((H, W), dtype=torch.uint8, device=dev)
    for k in range(n):
        rec._check(rec.lib.slc_bmp_unpack_device(rec.h, ptrs[k], infos[k], one.data_ptr(), None))
        rec.synchronize()
        assert np.array_equal(one.cpu().numpy(), want[k])
    # errors: a file of another size, a NULL pixel array
    other = capi.bmp_parse(synth.encode_bmp(np.zeros((8, 16), np.uint8)))
    with pytest.raises(capi.SlcError):
        rec.bmp_unpack_batch_device(ptrs[:1], [other], stack.data_ptr() + PAD)
    with pytest.raises(capi.SlcError):
        rec.bmp_unpack_batch_device([0], infos[:1], stack.data_ptr() + PAD)
    rec.bmp_unpack_batch_device([], [], stack.data_ptr() + PAD)      # empty set: nothing to do
    rec.close()
