"""CPU tests of the host side: C-ABI surface, calibration reader, configs, sharding."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT


def test_library_exports_every_declared_symbol(built_library):
    from structured_light_calculation_b200 import capi
    names = capi.declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(built_library, n), f"{n} declared in include/slcalc_b200.h but not exported"
    assert built_library.slc_abi_version() == 2
    assert built_library.slc_status_string(4).decode().startswith("no CUDA device")


def test_library_is_sm100a_only(built_library):
    from structured_light_calculation_b200 import capi
    out = subprocess.run(["cuobjdump", "-lelf", capi.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    archs = {line.split(".")[-2] for line in out.splitlines() if "sm_" in line}
    assert archs == {"sm_100a"}, archs


def test_cpp_host_api_symbols_present(built_library):
    from structured_light_calculation_b200 import capi
    out = subprocess.run(["nm", "-DC", capi.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    for sym in ("dynaframe::CDecodeGray::Decode()", "dynaframe::CDecodeGray::SetNumDigit(int, bool)",
                "dynaframe::CDecodePhase::SetNumMat(int, int)", "dynaframe::CDecodePhase::Decode()",
                "dynaframe::CCalculation::Init()", "dynaframe::CCalculation::CalculateFirst()",
                "dynaframe::CCalculation::Result(", "dynaframe::ErrorHandling("):
        assert sym in out, sym


def test_no_cuda_device_fails_loudly(built_library):
    """On a box without a GPU the product path refuses to run: no CPU fallback."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from structured_light_calculation_b200 import capi
    from structured_light_calculation_b200.configs import CONFIGS
    with pytest.raises(capi.SlcError) as e:
        capi.Reconstructor(CONFIGS["config1"])
    assert e.value.status == capi.SLC_ERR_NO_DEVICE
    assert "no CPU path" in e.value.message


def test_product_never_imports_oracle():
    """Nothing under the package (Python or C/CUDA) may reference oracle/."""
    pkg = os.path.join(ROOT, "structured_light_calculation_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.sep + "build" in dirpath or os.sep + "lib" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert "sl_oracle" not in text and "oracle/" not in text.replace("the oracle", ""), (dirpath, f)


def test_calibration_yaml_reader(base_calibration):
    c = base_calibration
    assert c.cam.shape == (3, 3) and c.T.shape == (3,)
    assert c.cam[0, 0] == 1.2138714552009253e+003 and c.cam[0, 2] == 319.5 and c.cam[1, 2] == 255.5
    assert c.T[2] == 3.9430125669975382e+000
    assert abs(np.linalg.det(c.R) - 1.0) < 1e-12
    from structured_light_calculation_b200.calibration import parse_opencv_yaml
    with pytest.raises(ValueError):
        parse_opencv_yaml("A: !!opencv-matrix\n   rows: 2\n   cols: 2\n   dt: d\n   data: [ 1., 2., 3. ]\n")


def test_configs_match_baseline_table():
    from structured_light_calculation_b200.configs import CONFIGS
    table = {"config1": (18, 35, 10, 20), "config2": (22, 39, 5, 10), "config3": (24, 41, 8, 16),
             "config5": (32, 49, 4, 8)}
    for name, (planes, bpp, gp, T) in table.items():
        c = CONFIGS[name]
        assert (c.planes, c.algorithmic_bytes_per_pixel, c.gray_period, c.phase_period) == (planes, bpp, gp, T)
        assert c.width % 16 == 0
    r = CONFIGS["reference_default"]
    assert (r.width, r.height, r.projector_width, r.gray_digits, r.phase_steps) == (1280, 1024, 1280, 6, 4)
    assert (r.gray_period, r.phase_period) == (20, 40)


def test_synth_is_deterministic(base_calibration):
    from structured_light_calculation_b200 import synth
    from structured_light_calculation_b200.configs import CONFIGS
    cfg = CONFIGS["config1"].with_(width=96, height=48)
    cal = synth.synthetic_calibration(cfg, base_calibration)
    s1 = synth.make_scene(cfg, cal)
    a = synth.render_stack(cfg, s1, noise_sigma=1.0, seed=3)
    b = synth.render_stack(cfg, synth.make_scene(cfg, cal), noise_sigma=1.0, seed=3)
    assert np.array_equal(a, b) and a.shape == (cfg.planes, 48, 96) and a.dtype == np.uint8
    assert not np.array_equal(a, synth.render_stack(cfg, s1, noise_sigma=1.0, seed=4))


def test_shard_range_partitions():
    from structured_light_calculation_b200.distributed import shard_range
    for n in (0, 1, 7, 4096, 4099):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 4, 4)


def test_c_shard_range_is_the_same_partition(built_library):
    """slc_shard_range (what slc_pool and a C++ caller use) == distributed.shard_range (what bench.py's ranks use)."""
    from structured_light_calculation_b200 import capi
    from structured_light_calculation_b200.distributed import shard_range
    for n in (0, 1, 7, 4096, 4099):
        for world in (1, 2, 3, 8):
            for r in range(world):
                assert capi.shard_range(n, r, world) == shard_range(n, r, world)
    with pytest.raises(capi.SlcError):
        capi.shard_range(4, 4, 4)


def test_pool_without_a_device_fails_loudly(built_library):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from structured_light_calculation_b200 import capi
    from structured_light_calculation_b200.configs import CONFIGS
    with pytest.raises(capi.SlcError) as e:
        capi.Pool(CONFIGS["config1"], [0, 1])
    assert e.value.status == capi.SLC_ERR_NO_DEVICE and "no CPU path" in e.value.message
    assert built_library.slc_pool_size(None) == 0


def _gloo_worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from structured_light_calculation_b200 import distributed as D
    r, _, w = D.init_process_group("gloo")
    lo, hi = D.shard_range(4096, r, w)
    D.barrier()
    total = D.sum_over_ranks(hi - lo)
    slowest = D.max_over_ranks(10.0 + r)
    out.put((r, lo, hi, total, slowest))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding():
    """world_size 2 over gloo: the N>1 host logic bench.py uses (shard, barrier, max over ranks)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in res] == [(0, 2048), (2048, 4096)]
    assert all(r[3] == 4096 and r[4] == 11.0 for r in res)
