"""ctypes loader for the CPU oracle (oracle/libsl_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  Nothing under
structured_light_calculation_b200/ may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsl_oracle.so")


class SloConfig(C.Structure):
    _fields_ = [
        ("width", C.c_int32),
        ("height", C.c_int32),
        ("projector_width", C.c_int32),
        ("gray_digits", C.c_int32),
        ("phase_steps", C.c_int32),
        ("fov_min", C.c_double),
        ("fov_max", C.c_double),
        ("modulation_min", C.c_float),
        ("threads", C.c_int32),
    ]


class SloCalib(C.Structure):
    _fields_ = [
        ("cam", C.c_double * 9),
        ("pro", C.c_double * 9),
        ("R", C.c_double * 9),
        ("T", C.c_double * 3),
    ]


class SloOutputs(C.Structure):
    _fields_ = [
        ("gray_val", C.c_void_p),
        ("phase_pix", C.c_void_p),
        ("proj_u", C.c_void_p),
        ("x", C.c_void_p),
        ("y", C.c_void_p),
        ("z", C.c_void_p),
        ("kbin", C.c_void_p),
        ("corr", C.c_void_p),
        ("mask", C.c_void_p),
        ("mod_ok", C.c_void_p),
    ]


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc, -ffp-contract=off)."""
    src = os.path.join(_HERE, "sl_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(
        os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "sl_oracle.h"))
    ):
        subprocess.run(["make", "-C", _HERE, "-B", "libsl_oracle.so"], check=True,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.slo_fast_atan2.restype = C.c_float
        L.slo_fast_atan2.argtypes = [C.c_float, C.c_float]
        L.slo_fast_atan2_array.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.slo_phase_pix_array.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        L.slo_check_div360.argtypes = [C.c_float, C.c_float, C.POINTER(C.c_ulonglong)]
        L.slo_check_div360.restype = C.c_ulonglong
        L.slo_default_gray_lut.argtypes = [C.c_int, C.c_void_p]
        L.slo_gray_period.argtypes = [C.POINTER(SloConfig)]
        L.slo_phase_period.argtypes = [C.POINTER(SloConfig)]
        L.slo_calibration.argtypes = [C.POINTER(SloConfig), C.POINTER(SloCalib), C.POINTER(C.c_double),
                                      C.POINTER(C.c_double), C.c_void_p, C.c_void_p, C.c_void_p]
        L.slo_reconstruct.argtypes = [C.POINTER(SloConfig), C.POINTER(SloCalib), C.c_void_p, C.c_void_p,
                                      C.POINTER(SloOutputs)]
        L.slo_decode_gray.argtypes = [C.POINTER(SloConfig), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.slo_decode_phase.argtypes = [C.POINTER(SloConfig), C.c_void_p, C.c_void_p, C.c_void_p]
        L.slo_time_reconstruct.argtypes = [C.POINTER(SloConfig), C.POINTER(SloCalib), C.c_void_p, C.c_int,
                                           C.c_void_p]
        L.slo_strip_regression.argtypes = [C.POINTER(SloConfig), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.slo_delta_p.argtypes = [C.POINTER(SloConfig)] + [C.c_void_p] * 5
        L.slo_dyna_frame.argtypes = [C.POINTER(SloConfig), C.POINTER(SloCalib)] + [C.c_void_p] * 9
        L.slo_max_threads.restype = C.c_int
        L.slo_result_text.argtypes = [C.POINTER(SloConfig), C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint, C.c_void_p,
                                      C.c_longlong, C.POINTER(C.c_longlong)]
        L.slo_result_text.restype = C.c_longlong
        L.slo_format_g6.argtypes = [C.c_double, C.c_uint, C.c_char_p]
        L.slo_format_g6.restype = C.c_int
        L.slo_triangulate_uv.argtypes = [C.POINTER(SloConfig), C.POINTER(SloCalib)] + [C.c_void_p] * 6
        _lib = L
    return _lib


def make_config(width, height, projector_width, gray_digits, phase_steps, fov_min=10.0, fov_max=100.0,
                modulation_min=0.0, threads=1) -> SloConfig:
    return SloConfig(width, height, projector_width, gray_digits, phase_steps, float(fov_min), float(fov_max),
                     float(modulation_min), threads)


def make_calib(cam, pro, R, T) -> SloCalib:
    c = SloCalib()
    c.cam[:] = [float(v) for v in np.asarray(cam, dtype=np.float64).reshape(9)]
    c.pro[:] = [float(v) for v in np.asarray(pro, dtype=np.float64).reshape(9)]
    c.R[:] = [float(v) for v in np.asarray(R, dtype=np.float64).reshape(9)]
    c.T[:] = [float(v) for v in np.asarray(T, dtype=np.float64).reshape(3)]
    return c


def fast_atan2(y, x) -> np.ndarray:
    """Vector wrapper over the scalar C restatement of cv::fastAtan2."""
    L = lib()
    y = np.asarray(y, dtype=np.float32)
    x = np.asarray(x, dtype=np.float32)
    yb, xb = np.broadcast_arrays(y, x)
    yb = np.ascontiguousarray(yb)
    xb = np.ascontiguousarray(xb)
    out = np.empty(yb.shape, dtype=np.float32)
    L.slo_fast_atan2_array(yb.ctypes.data, xb.ctypes.data, out.ctypes.data, out.size)
    return out


def phase_pix(deg, period: int) -> np.ndarray:
    deg = np.ascontiguousarray(deg, dtype=np.float32)
    out = np.empty(deg.shape, np.float32)
    lib().slo_phase_pix_array(deg.ctypes.data, out.ctypes.data, out.size, int(period))
    return out


def check_div360(lo: float, hi: float):
    n = C.c_ulonglong()
    bad = lib().slo_check_div360(lo, hi, C.byref(n))
    return int(bad), int(n.value)


def default_gray_lut(n_digits: int) -> np.ndarray:
    lut = np.zeros(1 << n_digits, dtype=np.int16)
    lib().slo_default_gray_lut(n_digits, lut.ctypes.data)
    return lut


def calibration(cfg: SloConfig, cal: SloCalib, want_luts=True):
    A = C.c_double()
    B = C.c_double()
    P = np.zeros(12, dtype=np.float64)
    cC = cD = None
    if want_luts:
        cC = np.empty((cfg.height, cfg.width), dtype=np.float64)
        cD = np.empty((cfg.height, cfg.width), dtype=np.float64)
    lib().slo_calibration(C.byref(cfg), C.byref(cal), C.byref(A), C.byref(B),
                          cC.ctypes.data if want_luts else None, cD.ctypes.data if want_luts else None,
                          P.ctypes.data)
    return A.value, B.value, cC, cD, P.reshape(3, 4)


def reconstruct(cfg: SloConfig, cal: SloCalib, planes: np.ndarray, gray_lut: np.ndarray | None = None) -> dict:
    """Run the staged CPU path on one stack; returns every plane as numpy."""
    H, W = cfg.height, cfg.width
    P = 2 * cfg.gray_digits + cfg.phase_steps
    planes = np.ascontiguousarray(planes, dtype=np.uint8)
    assert planes.shape == (P, H, W), (planes.shape, (P, H, W))
    res = {
        "gray_val": np.empty((H, W), np.float64),
        "phase_pix": np.empty((H, W), np.float64),
        "proj_u": np.empty((H, W), np.float64),
        "x": np.empty((H, W), np.float64),
        "y": np.empty((H, W), np.float64),
        "z": np.empty((H, W), np.float64),
        "kbin": np.empty((H, W), np.int16),
        "corr": np.empty((H, W), np.int8),
        "mask": np.empty((H, W), np.uint8),
        "mod_ok": np.empty((H, W), np.uint8),
    }
    out = SloOutputs(*[res[k].ctypes.data for k, _ in SloOutputs._fields_])
    lut_ptr = None
    if gray_lut is not None:
        gray_lut = np.ascontiguousarray(gray_lut, dtype=np.int16)
        assert gray_lut.size == 1 << cfg.gray_digits
        lut_ptr = gray_lut.ctypes.data
    rc = lib().slo_reconstruct(C.byref(cfg), C.byref(cal), lut_ptr, planes.ctypes.data, C.byref(out))
    if rc != 0:
        raise ValueError(f"slo_reconstruct failed rc={rc}")
    return res


def decode_gray(cfg: SloConfig, gray_planes: np.ndarray, gray_lut=None):
    H, W = cfg.height, cfg.width
    gray_planes = np.ascontiguousarray(gray_planes, dtype=np.uint8)
    assert gray_planes.shape == (2 * cfg.gray_digits, H, W)
    val = np.empty((H, W), np.float64)
    kbin = np.empty((H, W), np.int16)
    lut_ptr = None
    if gray_lut is not None:
        gray_lut = np.ascontiguousarray(gray_lut, dtype=np.int16)
        lut_ptr = gray_lut.ctypes.data
    rc = lib().slo_decode_gray(C.byref(cfg), lut_ptr, gray_planes.ctypes.data, val.ctypes.data, kbin.ctypes.data)
    if rc != 0:
        raise ValueError(f"slo_decode_gray failed rc={rc}")
    return val, kbin


def decode_phase(cfg: SloConfig, phase_planes: np.ndarray):
    H, W = cfg.height, cfg.width
    phase_planes = np.ascontiguousarray(phase_planes, dtype=np.uint8)
    assert phase_planes.shape == (cfg.phase_steps, H, W)
    pix = np.empty((H, W), np.float64)
    mod = np.empty((H, W), np.uint8)
    rc = lib().slo_decode_phase(C.byref(cfg), phase_planes.ctypes.data, pix.ctypes.data, mod.ctypes.data)
    if rc != 0:
        raise ValueError(f"slo_decode_phase failed rc={rc}")
    return pix, mod


def time_reconstruct(cfg: SloConfig, cal: SloCalib, planes: np.ndarray, reps: int) -> np.ndarray:
    planes = np.ascontiguousarray(planes, dtype=np.uint8)
    secs = np.zeros(reps, dtype=np.float64)
    rc = lib().slo_time_reconstruct(C.byref(cfg), C.byref(cal), planes.ctypes.data, reps, secs.ctypes.data)
    if rc != 0:
        raise ValueError(f"slo_time_reconstruct failed rc={rc}")
    return secs


def max_threads() -> int:
    return int(lib().slo_max_threads())


# ---- dynamic frames --------------------------------------------------------
def strip_regression(cfg: SloConfig, image: np.ndarray, window: int = 21):
    image = np.ascontiguousarray(image, dtype=np.uint8)
    assert image.shape == (cfg.height, cfg.width)
    B = np.empty(image.shape, np.float32)
    W = np.empty(image.shape, np.float32)
    lib().slo_strip_regression(C.byref(cfg), window, image.ctypes.data, B.ctypes.data, W.ctypes.data)
    return B, W


def delta_p(cfg: SloConfig, B0, W0, B1, W1) -> np.ndarray:
    arrs = [np.ascontiguousarray(a, dtype=np.float32) for a in (B0, W0, B1, W1)]
    out = np.empty((cfg.height, cfg.width), np.float32)
    lib().slo_delta_p(C.byref(cfg), *[a.ctypes.data for a in arrs], out.ctypes.data)
    return out


def dyna_frame(cfg: SloConfig, cal: SloCalib, U0: np.ndarray, dP: np.ndarray, z0: np.ndarray) -> dict:
    H, W = cfg.height, cfg.width
    U0 = np.ascontiguousarray(U0, dtype=np.float64)
    dP = np.ascontiguousarray(dP, dtype=np.float32)
    z0 = np.ascontiguousarray(z0, dtype=np.float64)
    res = {k: np.empty((H, W), np.float64) for k in ("proj_u", "x", "y", "z", "delta_z")}
    res["mask"] = np.empty((H, W), np.uint8)
    lib().slo_dyna_frame(C.byref(cfg), C.byref(cal), U0.ctypes.data, dP.ctypes.data, z0.ctypes.data,
                         res["proj_u"].ctypes.data, res["x"].ctypes.data, res["y"].ctypes.data,
                         res["z"].ctypes.data, res["delta_z"].ctypes.data, res["mask"].ctypes.data)
    return res


def dyna_sequence(cfg: SloConfig, cal: SloCalib, U0: np.ndarray, z0: np.ndarray, frames: np.ndarray,
                  window: int = 21) -> list:
    """CalculateOther (CCalculation.cpp:221-317) over frames[1:], frames[0] being the image
    StripRegression(0) saw at the end of CalculateFirst (:201)."""
    B0, W0 = strip_regression(cfg, frames[0], window)
    out = []
    U, z = U0, z0
    for f in range(1, frames.shape[0]):
        B1, W1 = strip_regression(cfg, frames[f], window)
        dP = delta_p(cfg, B0, W0, B1, W1)
        r = dyna_frame(cfg, cal, U, dP, z)
        r.update(strip_b=B1, strip_w=W1, delta_p=dP)
        out.append(r)
        U, z, B0, W0 = r["proj_u"], r["z"], B1, W1
    return out


TEXT_CRLF, TEXT_EXP3 = 1, 2


def result_text(cfg: SloConfig, x: np.ndarray, y: np.ndarray, z: np.ndarray, flags: int = 0):
    """CCalculation::Result (CCalculation.cpp:323-357) on f64 x, y, z planes -> (bytes, n_points)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    z = np.ascontiguousarray(z, dtype=np.float64)
    cap = 44 * x.size + 16
    buf = np.empty(cap, dtype=np.uint8)
    pts = C.c_longlong(0)
    n = lib().slo_result_text(C.byref(cfg), x.ctypes.data, y.ctypes.data, z.ctypes.data, flags, buf.ctypes.data, cap,
                              C.byref(pts))
    return buf[:n].tobytes(), int(pts.value)


def format_g6(v: float, flags: int = 0) -> bytes:
    out = C.create_string_buffer(40)
    n = lib().slo_format_g6(float(v), flags, out)
    return out.raw[:n]


def triangulate_uv(cfg: SloConfig, cal: SloCalib, U: np.ndarray, V: np.ndarray) -> dict:
    """[EXT] least-squares z from the projector column U and row V."""
    U = np.ascontiguousarray(U, dtype=np.float64)
    V = np.ascontiguousarray(V, dtype=np.float64)
    shape = (cfg.height, cfg.width)
    out = {k: np.empty(shape, np.float64) for k in "xyz"}
    out["mask"] = np.empty(shape, np.uint8)
    lib().slo_triangulate_uv(C.byref(cfg), C.byref(cal), U.ctypes.data, V.ctypes.data, out["x"].ctypes.data,
                             out["y"].ctypes.data, out["z"].ctypes.data, out["mask"].ctypes.data)
    return out
