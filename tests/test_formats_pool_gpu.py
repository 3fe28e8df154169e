"""GPU tests (-m gpu) of the result formats (slc_reconstruct_*_ex) and of slc_pool: every format is a
bit-for-bit selection of the parity-checked xyzw + mask output, a pool's output equals a single
context's, contexts driven from different host threads do not disturb each other, and a launch
covers any number of frame sets."""
import threading

import numpy as np
import pytest

from conftest import bits_equal, make_case

pytestmark = pytest.mark.gpu


def _cfg(name="config2", W=384, H=160):
    from structured_light_calculation_b200.configs import CONFIGS
    return CONFIGS[name].with_(width=W, height=H)


def _stacks(cfg, base_calibration, n, seed=3):
    from structured_light_calculation_b200 import synth
    cal, scene, first = make_case(cfg, base_calibration, noise=1.5, seed=seed)
    rest = [synth.render_stack(cfg, scene, noise_sigma=1.5, seed=seed + 1 + i) for i in range(n - 1)]
    return cal, np.stack([first] + rest)


def _select(full, i, order):
    from structured_light_calculation_b200 import capi
    xyz, m = full["xyzw"][i, ..., :3], full["mask"][i].astype(bool)
    return xyz[m] if order == capi.SLC_ORDER_ROW_MAJOR else np.transpose(xyz, (1, 0, 2))[m.T]


@pytest.mark.parametrize("max_batch,num_slots", [(1, 1), (2, 3), (3, 2)])
def test_host_formats_over_many_chunks(built_library, base_calibration, max_batch, num_slots):
    """7 frame sets through the pipelined host path in every format, chunked over the stream slots."""
    from structured_light_calculation_b200 import capi
    cfg = _cfg()
    cal, stacks = _stacks(cfg, base_calibration, 7)
    rec = capi.Reconstructor(cfg, device=0, max_batch=max_batch, num_slots=num_slots)
    rec.set_calibration(cal)
    full = rec.reconstruct(stacks)
    n, npx = 7, cfg.pixels
    assert full["mask"].any() and not full["mask"].all()
    d = rec.reconstruct_ex(stacks, capi.SLC_RESULT_DEPTH)
    assert bits_equal(d["depth"], np.ascontiguousarray(full["xyzw"][..., 2]))
    assert np.array_equal(capi.unpack_mask_bits(d["mask_bits"], n, npx), full["mask"].reshape(n, npx))
    for order in (capi.SLC_ORDER_ROW_MAJOR, capi.SLC_ORDER_REFERENCE):
        p = rec.reconstruct_ex(stacks, capi.SLC_RESULT_POINTS, order, full_maps=(order == capi.SLC_ORDER_REFERENCE))
        for i in range(n):
            want = _select(full, i, order)
            assert int(p["n_points"][i]) == len(want)
            assert bits_equal(np.ascontiguousarray(p["points"][i, : len(want)]), np.ascontiguousarray(want))
        assert np.array_equal(capi.unpack_mask_bits(p["mask_bits"], n, npx), full["mask"].reshape(n, npx))
        if "xyzw" in p:          # the maps can ride along
            assert bits_equal(p["xyzw"], full["xyzw"]) and bits_equal(p["mask"], full["mask"])
    rec.close()


def test_points_stride_overflow_is_reported(built_library, base_calibration):
    from structured_light_calculation_b200 import capi
    cfg = _cfg()
    cal, stacks = _stacks(cfg, base_calibration, 2)
    rec = capi.Reconstructor(cfg, device=0, max_batch=1, num_slots=2)
    rec.set_calibration(cal)
    full = rec.reconstruct(stacks)
    counts = full["mask"].reshape(2, -1).sum(axis=1)
    stride = int(counts.min()) // 2
    bufs, res = capi.alloc_result(cfg, 2, capi.SLC_RESULT_POINTS, point_stride=stride)
    bufs["points"][...] = -7.0
    with pytest.raises(capi.SlcError) as e:
        rec.reconstruct_into_ex(stacks, 2, res)
    assert e.value.status == capi.SLC_ERR_INVALID_ARG and "point_stride" in e.value.message
    assert list(bufs["n_points"]) == list(counts)                      # the full counts
    for i in range(2):                                                  # and the first `stride` points of each
        assert bits_equal(np.ascontiguousarray(bufs["points"][i]), np.ascontiguousarray(_select(full, i, 0)[:stride]))
    rec.close()


def test_device_formats(built_library, base_calibration):
    """slc_reconstruct_device_ex: DEPTH straight from the fused kernel, POINTS with one extra launch,
    on a caller stream, 5 frame sets in one call."""
    import torch
    from structured_light_calculation_b200 import capi
    cfg = _cfg("config3", 272, 128)
    cal, stacks = _stacks(cfg, base_calibration, 5)
    rec = capi.Reconstructor(cfg, device=0, max_batch=1, num_slots=1)
    rec.set_calibration(cal)
    full = rec.reconstruct(stacks)
    n, npx, H, W = 5, cfg.pixels, cfg.height, cfg.width
    dev = torch.device("cuda", 0)
    d_in = torch.from_numpy(stacks).to(dev)
    stream = torch.cuda.Stream(device=dev)
    d_depth = torch.zeros((n, H, W), dtype=torch.float32, device=dev)
    d_bits = torch.zeros((n * capi.bits_bytes(npx) + 3) // 4 * 4, dtype=torch.uint8, device=dev)
    l0 = rec.launch_count()
    rec.reconstruct_device_ex(d_in.data_ptr(), n, capi.make_result(capi.SLC_RESULT_DEPTH, depth=d_depth.data_ptr(),
                                                                   mask_bits=d_bits.data_ptr()), stream.cuda_stream)
    stream.synchronize()
    assert rec.launch_count() - l0 == 1
    assert bits_equal(d_depth.cpu().numpy(), np.ascontiguousarray(full["xyzw"][..., 2]))
    assert np.array_equal(capi.unpack_mask_bits(d_bits.cpu().numpy(), n, npx), full["mask"].reshape(n, npx))
    d_xyzw = torch.zeros((n, H, W, 4), dtype=torch.float32, device=dev)
    d_mask = torch.zeros((n, H, W), dtype=torch.uint8, device=dev)
    d_pts = torch.zeros((n, npx, 3), dtype=torch.float32, device=dev)
    d_cnt = torch.zeros(n, dtype=torch.int64, device=dev)
    for order in (capi.SLC_ORDER_ROW_MAJOR, capi.SLC_ORDER_REFERENCE, capi.SLC_ORDER_REFERENCE):   # state reuse: epochs
        l0 = rec.launch_count()
        rec.reconstruct_device_ex(d_in.data_ptr(), n, capi.make_result(
            capi.SLC_RESULT_POINTS, order, xyzw=d_xyzw.data_ptr(), mask=d_mask.data_ptr(), points=d_pts.data_ptr(),
            point_stride=npx, n_points=d_cnt.data_ptr(), mask_bits=d_bits.data_ptr()), stream.cuda_stream)
        stream.synchronize()
        assert rec.launch_count() - l0 == 2
        pts, cnt = d_pts.cpu().numpy(), d_cnt.cpu().numpy()
        for i in range(n):
            want = _select(full, i, order)
            assert int(cnt[i]) == len(want)
            assert bits_equal(np.ascontiguousarray(pts[i, : len(want)]), np.ascontiguousarray(want))
    rec.close()


def test_depth_format_on_the_any_geometry_kernel(built_library, base_calibration):
    """Widths the vector kernel cannot take (and the forced scalar kernel) write DEPTH through atomicOr bits."""
    from structured_light_calculation_b200 import capi
    for W, H, flags in ((100, 37, 0), (36, 7, 0), (384, 160, capi.SLC_FLAG_SCALAR_KERNEL)):
        cfg = _cfg("config1", W, H)
        cal, stacks = _stacks(cfg, base_calibration, 3)
        rec = capi.Reconstructor(cfg, device=0, max_batch=2, num_slots=2, flags=flags)
        rec.set_calibration(cal)
        assert rec.info().kernel_variant == 2
        full = rec.reconstruct(stacks)
        d = rec.reconstruct_ex(stacks, capi.SLC_RESULT_DEPTH)
        assert bits_equal(d["depth"], np.ascontiguousarray(full["xyzw"][..., 2]))
        assert np.array_equal(capi.unpack_mask_bits(d["mask_bits"], 3, cfg.pixels), full["mask"].reshape(3, -1))
        rec.close()


def test_full_size_formats(built_library, base_calibration):
    """configs[1] at full size: the formats against the full map, both point orders."""
    from structured_light_calculation_b200 import capi
    from structured_light_calculation_b200.configs import CONFIGS
    cfg = CONFIGS["config2"]
    cal, scene, planes = make_case(cfg, base_calibration, noise=1.0, seed=11)
    rec = capi.Reconstructor(cfg, device=0, max_batch=1, num_slots=2)
    rec.set_calibration(cal)
    full = rec.reconstruct(planes)
    d = rec.reconstruct_ex(planes, capi.SLC_RESULT_DEPTH)
    assert bits_equal(d["depth"], np.ascontiguousarray(full["xyzw"][..., 2]))
    assert np.array_equal(capi.unpack_mask_bits(d["mask_bits"], 1, cfg.pixels), full["mask"].reshape(1, -1))
    for order in (capi.SLC_ORDER_ROW_MAJOR, capi.SLC_ORDER_REFERENCE):
        p = rec.reconstruct_ex(planes, capi.SLC_RESULT_POINTS, order)
        want = _select(full, 0, order)
        assert int(p["n_points"][0]) == len(want) > 0
        assert bits_equal(np.ascontiguousarray(p["points"][0, : len(want)]), np.ascontiguousarray(want))
        # the stand-alone compactor takes the same kernel
        assert bits_equal(rec.pointcloud_compact(full["xyzw"][0], full["mask"][0], order), np.ascontiguousarray(want))
    rec.close()


def test_more_frame_sets_than_grid_y(built_library, base_calibration):
    """66 000 tiny frame sets in ONE call: the launcher loops over grid.y = 65535."""
    import torch
    from structured_light_calculation_b200 import capi
    cfg = _cfg("config1", 32, 8)
    cal, stacks = _stacks(cfg, base_calibration, 3)
    rec = capi.Reconstructor(cfg, device=0, max_batch=1, num_slots=1)
    rec.set_calibration(cal)
    full = rec.reconstruct(stacks)
    n = 66000
    dev = torch.device("cuda", 0)
    d_in = torch.from_numpy(stacks).to(dev)[torch.arange(n, device=dev) % 3].contiguous()
    d_xyzw = torch.zeros((n, cfg.height, cfg.width, 4), dtype=torch.float32, device=dev)
    d_mask = torch.zeros((n, cfg.height, cfg.width), dtype=torch.uint8, device=dev)
    l0 = rec.launch_count()
    rec.reconstruct_device(d_in.data_ptr(), n, d_xyzw.data_ptr(), d_mask.data_ptr())
    rec.synchronize()
    assert rec.launch_count() - l0 == 2
    for i in (0, 1, 65534, 65535, 65536, n - 1):
        assert bits_equal(d_xyzw[i].cpu().numpy(), full["xyzw"][i % 3]) and bits_equal(d_mask[i].cpu().numpy(), full["mask"][i % 3])
    assert rec.time_device(d_in.data_ptr(), n, d_xyzw.data_ptr(), d_mask.data_ptr(), 1) > 0
    rec.close()


@pytest.mark.parametrize("fmt_name", ["xyzw", "depth", "points"])
def test_pool_equals_single_context(built_library, base_calibration, fmt_name):
    """A pool (two or three members, here all on device 0 -- the driver's GPU tests see one GPU) must return
    exactly what one context returns: shards are contiguous, calibration is replicated."""
    from structured_light_calculation_b200 import capi
    cfg = _cfg()
    cal, stacks = _stacks(cfg, base_calibration, 7)
    rec = capi.Reconstructor(cfg, device=0, max_batch=2, num_slots=2)
    rec.set_calibration(cal)
    fmt = {"xyzw": capi.SLC_RESULT_XYZW, "depth": capi.SLC_RESULT_DEPTH, "points": capi.SLC_RESULT_POINTS}[fmt_name]
    one, res1 = capi.alloc_result(cfg, 7, fmt)
    rec.reconstruct_into_ex(stacks, 7, res1)
    rec.close()
    for members in (2, 3):
        pool = capi.Pool(cfg, [0] * members, max_batch=2, num_slots=2)
        pool.set_calibration(cal)
        for n in (7, 1, 0):                       # fewer frame sets than members, and none at all
            many, res = capi.alloc_result(cfg, 7, fmt)
            for a in many.values():
                a[...] = 0
            pool.reconstruct_into_ex(stacks, n, res)
            assert sum(pool.last_shares()) == n          # handed out on demand, every frame set exactly once
            for k in many:
                if k == "points":
                    for i in range(n):
                        c = int(one["n_points"][i])
                        assert bits_equal(np.ascontiguousarray(many[k][i, :c]), np.ascontiguousarray(one[k][i, :c]))
                elif k == "mask_bits":
                    nb = n * capi.bits_bytes(cfg.pixels)
                    assert bits_equal(many[k][:nb], one[k][:nb])
                else:
                    assert bits_equal(np.ascontiguousarray(many[k][:n]), np.ascontiguousarray(one[k][:n])), k
        pool.close()


def test_pool_device_shards(built_library, base_calibration):
    import torch
    from structured_light_calculation_b200 import capi
    cfg = _cfg()
    cal, stacks = _stacks(cfg, base_calibration, 5)
    rec = capi.Reconstructor(cfg, device=0, max_batch=1, num_slots=1)
    rec.set_calibration(cal)
    full = rec.reconstruct(stacks)
    rec.close()
    dev = torch.device("cuda", 0)
    pool = capi.Pool(cfg, [0, 0], max_batch=1, num_slots=1)
    pool.set_calibration(cal)
    d_in = torch.from_numpy(stacks).to(dev)
    d_xyzw = torch.zeros((5, cfg.height, cfg.width, 4), dtype=torch.float32, device=dev)
    d_mask = torch.zeros((5, cfg.height, cfg.width), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    shards = [capi.shard_range(5, i, 2) for i in range(2)]
    pool.reconstruct_device([d_in[lo].data_ptr() for lo, _ in shards], [hi - lo for lo, hi in shards],
                            [capi.make_result(capi.SLC_RESULT_XYZW, xyzw=d_xyzw[lo].data_ptr(), mask=d_mask[lo].data_ptr())
                             for lo, _ in shards])
    assert bits_equal(d_xyzw.cpu().numpy(), full["xyzw"]) and bits_equal(d_mask.cpu().numpy(), full["mask"])
    pool.close()


def test_two_contexts_from_two_host_threads(built_library, base_calibration):
    """The header's promise: distinct contexts may be driven from distinct host threads.  Two contexts of
    DIFFERENT geometry and per-context pixels-per-thread run 20 host calls each, concurrently; every result
    must equal the one computed alone."""
    from structured_light_calculation_b200 import capi
    cases = []
    for name, W, H, pxt in (("config2", 384, 160, 4), ("config3", 272, 128, 8)):
        cfg = _cfg(name, W, H)
        cal, stacks = _stacks(cfg, base_calibration, 3)
        rec = capi.Reconstructor(cfg, device=0, max_batch=2, num_slots=2)
        rec.set_calibration(cal)
        rec.set_pixels_per_thread(pxt)
        cases.append((rec, stacks, rec.reconstruct(stacks)))
    errors = []

    def work(rec, stacks, want):
        try:
            for it in range(20):
                got = rec.reconstruct(stacks) if it % 2 == 0 else None
                if got is not None:
                    assert bits_equal(got["xyzw"], want["xyzw"]) and bits_equal(got["mask"], want["mask"])
                else:
                    d = rec.reconstruct_ex(stacks, capi.SLC_RESULT_DEPTH)
                    assert bits_equal(d["depth"], np.ascontiguousarray(want["xyzw"][..., 2]))
        except Exception as e:      # noqa: BLE001 -- reported by the main thread
            errors.append(repr(e))

    threads = [threading.Thread(target=work, args=c) for c in cases]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for rec, _, _ in cases:
        rec.close()
    assert not errors, errors


@pytest.mark.parametrize("W,H", [(320, 128), (200, 75)])
def test_dynamic_frames_result_formats(built_library, base_calibration, W, H):
    """slc_dyna_track_host_ex: depth + bits / valid points of every dynamic frame == selections of the full maps
    (any width for DEPTH; POINTS needs a width that is a multiple of 8)."""
    from structured_light_calculation_b200 import capi, synth
    cfg = _cfg("reference_default", W, H)
    cal, scene, planes = make_case(cfg, base_calibration, noise=1.0, seed=41)
    frames = synth.render_dyna_frames(cfg, cal, 5, stripe_period=14.0, z_step=0.4, noise_sigma=1.5)
    rec = capi.Reconstructor(cfg, device=0, max_batch=1, num_slots=1)
    rec.set_calibration(cal)
    u0 = rec.reconstruct(planes, parity=True)["proj_u"][0]
    full = rec.dyna_track(frames, u0, window=21)
    n, npx = 4, cfg.pixels
    assert full["mask"].any()
    bufs, res = capi.alloc_result(cfg, n, capi.SLC_RESULT_DEPTH)
    rec.dyna_track_into_ex(frames, 5, u0, res, window=21)
    assert bits_equal(bufs["depth"], np.ascontiguousarray(full["xyzw"][..., 2]))
    assert np.array_equal(capi.unpack_mask_bits(bufs["mask_bits"], n, npx), full["mask"].reshape(n, npx))
    for order in (capi.SLC_ORDER_ROW_MAJOR, capi.SLC_ORDER_REFERENCE):
        bufs, res = capi.alloc_result(cfg, n, capi.SLC_RESULT_POINTS, order)
        if W % 8:
            with pytest.raises(capi.SlcError):
                rec.dyna_track_into_ex(frames, 5, u0, res, window=21)
            continue
        rec.dyna_track_into_ex(frames, 5, u0, res, window=21)
        for i in range(n):
            want = _select(full, i, order)
            assert int(bufs["n_points"][i]) == len(want)
            assert bits_equal(np.ascontiguousarray(bufs["points"][i, : len(want)]), np.ascontiguousarray(want))
    # a single frame has no dynamic map: nothing is written, nothing fails
    rec.dyna_track_into_ex(frames, 1, u0, res, window=21)
    rec.close()


@pytest.mark.parametrize("H", [2048, 3000])
def test_points_format_on_tall_images(built_library, base_calibration, H):
    """Result()'s order keeps 17 bytes of shared memory per image row: above ~1750 rows the kernel needs the
    opt-in for more than 48 KB (BASELINE configs[2] and [4] are 2048 and 3000 rows tall)."""
    from structured_light_calculation_b200 import capi
    cfg = _cfg("config1", 64, H)
    cal, stacks = _stacks(cfg, base_calibration, 2)
    rec = capi.Reconstructor(cfg, device=0, max_batch=2, num_slots=1)
    rec.set_calibration(cal)
    full = rec.reconstruct(stacks)
    for order in (capi.SLC_ORDER_REFERENCE, capi.SLC_ORDER_ROW_MAJOR):
        p = rec.reconstruct_ex(stacks, capi.SLC_RESULT_POINTS, order)
        for i in range(2):
            want = _select(full, i, order)
            assert int(p["n_points"][i]) == len(want) > 0
            assert bits_equal(np.ascontiguousarray(p["points"][i, : len(want)]), np.ascontiguousarray(want))
    rec.close()
