import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def base_calibration():
    from structured_light_calculation_b200.calibration import load_calibration
    return load_calibration(os.path.join(GOLDEN, "Result.yml"))


@pytest.fixture(scope="session")
def oracle():
    from oracle import sl_oracle
    sl_oracle.lib()
    return sl_oracle


@pytest.fixture(scope="session")
def built_library():
    """The CUDA library must exist: build it in-tree if this checkout has not yet."""
    from structured_light_calculation_b200 import capi
    if not os.path.exists(capi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return capi.load_library()


def make_case(cfg, base_calibration, noise=1.0, seed=1):
    """Synthetic calibration + scene + rendered stack for a StackConfig."""
    from structured_light_calculation_b200 import synth
    cal = synth.synthetic_calibration(cfg, base_calibration)
    scene = synth.make_scene(cfg, cal)
    planes = synth.render_stack(cfg, scene, noise_sigma=noise, seed=seed)
    return cal, scene, planes


def oracle_run(oracle, cfg, cal, planes, lut=None, threads=1):
    ocfg = oracle.make_config(cfg.width, cfg.height, cfg.projector_width, cfg.gray_digits, cfg.phase_steps,
                              cfg.fov_min, cfg.fov_max, cfg.modulation_min, threads)
    ocal = oracle.make_calib(cal.cam, cal.pro, cal.R, cal.T)
    return oracle.reconstruct(ocfg, ocal, planes, lut)


def bits_equal(a, b):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    return a.dtype == b.dtype and a.shape == b.shape and a.tobytes() == b.tobytes()
