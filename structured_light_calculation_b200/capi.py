"""ctypes binding of libslcalc_b200.so (include/slcalc_b200.h).

Host plumbing only.  There is no CPU fallback: if the shared library is not
built, or no CUDA device is present, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import re

import numpy as np

from .configs import StackConfig

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "lib", "libslcalc_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_PKG), "include", "slcalc_b200.h")

SLC_OK = 0
SLC_ERR_INVALID_ARG = 1
SLC_ERR_NOT_INITIALISED = 2
SLC_ERR_CUDA = 3
SLC_ERR_NO_DEVICE = 4
SLC_ERR_OUT_OF_MEMORY = 5
SLC_ERR_STATE = 6

SLC_FLAG_Z_FP64 = 1 << 0
SLC_FLAG_SCALAR_KERNEL = 1 << 1

SLC_ORDER_ROW_MAJOR = 0
SLC_ORDER_REFERENCE = 1
SLC_TEXT_CRLF = 1 << 0
SLC_TEXT_EXP3 = 1 << 1

SLC_RESULT_XYZW = 0
SLC_RESULT_DEPTH = 1
SLC_RESULT_POINTS = 2

SLC_HOST_WRITE_COMBINED = 1 << 0
SLC_HOST_HUGE_PAGES = 1 << 1


class SlcError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"slcalc_b200 status {status}: {message}")
        self.status = status
        self.message = message


class SlcConfig(C.Structure):
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32), ("projector_width", C.c_int32),
        ("gray_digits", C.c_int32), ("phase_steps", C.c_int32),
        ("fov_min", C.c_double), ("fov_max", C.c_double),
        ("modulation_min", C.c_float), ("flags", C.c_uint32),
        ("device", C.c_int32), ("max_batch", C.c_int32), ("num_slots", C.c_int32),
    ]


class SlcInfo(C.Structure):
    _fields_ = [
        ("planes", C.c_int32), ("gray_period", C.c_int32), ("phase_period", C.c_int32), ("sm_count", C.c_int32),
        ("pixels", C.c_int64), ("stack_bytes", C.c_int64), ("xyzw_bytes", C.c_int64), ("mask_bytes", C.c_int64),
        ("kernel_variant", C.c_int32), ("kernel_regs", C.c_int32), ("kernel_block", C.c_int32),
        ("kernel_smem", C.c_int32),
    ]


class SlcBmpInfo(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("bits_per_pixel", C.c_int32), ("top_down", C.c_int32),
                ("row_stride", C.c_int32), ("palette_is_identity", C.c_int32), ("pixel_offset", C.c_int64),
                ("gray", C.c_uint8 * 256)]


class SlcDynaParity(C.Structure):
    _fields_ = [("strips", C.c_void_p), ("delta_p", C.c_void_p), ("proj_u", C.c_void_p)]


class SlcResult(C.Structure):
    _fields_ = [("format", C.c_int32), ("order", C.c_int32), ("xyzw", C.c_void_p), ("mask", C.c_void_p),
                ("depth", C.c_void_p), ("mask_bits", C.c_void_p), ("points", C.c_void_p),
                ("point_stride", C.c_int64), ("n_points", C.c_void_p)]


class SlcParityPlanes(C.Structure):
    _fields_ = [("kbin", C.c_void_p), ("corr", C.c_void_p), ("phase_pix", C.c_void_p), ("proj_u", C.c_void_p)]


def declared_symbols() -> list[str]:
    """Every function name include/slcalc_b200.h declares."""
    with open(HEADER_PATH, "r", encoding="utf-8") as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(slc_[a-z0-9_]+)\s*\(", text)))


_lib = None


def load_library():
    """dlopen the CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(
            f"{LIB_PATH} is missing: build it with __graft_entry__.build() or "
            "`make -C structured_light_calculation_b200` (there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i32 = C.c_void_p, C.c_int32
    L.slc_create.argtypes = [C.POINTER(SlcConfig), C.POINTER(vp)]
    L.slc_destroy.argtypes = [vp]
    L.slc_destroy.restype = None
    L.slc_last_error.argtypes = [vp]
    L.slc_last_error.restype = C.c_char_p
    L.slc_status_string.argtypes = [C.c_int]
    L.slc_status_string.restype = C.c_char_p
    L.slc_get_info.argtypes = [vp, C.POINTER(SlcInfo)]
    L.slc_set_calibration.argtypes = [vp, vp, vp, vp, vp]
    L.slc_set_gray_lut.argtypes = [vp, vp, i32]
    L.slc_host_alloc.argtypes = [C.c_size_t]
    L.slc_host_alloc.restype = vp
    L.slc_host_free.argtypes = [vp]
    L.slc_host_free.restype = None
    L.slc_host_alloc_ex.argtypes = [C.c_size_t, C.c_uint32]
    L.slc_host_alloc_ex.restype = vp
    L.slc_host_register.argtypes = [vp, C.c_size_t]
    L.slc_host_unregister.argtypes = [vp]
    L.slc_device_alloc.argtypes = [vp, C.c_size_t]
    L.slc_device_alloc.restype = vp
    L.slc_device_free.argtypes = [vp, vp]
    L.slc_device_free.restype = None
    L.slc_copy_to_device.argtypes = [vp, vp, vp, C.c_size_t]
    L.slc_copy_to_host.argtypes = [vp, vp, vp, C.c_size_t]
    L.slc_synchronize.argtypes = [vp]
    L.slc_reconstruct_device.argtypes = [vp, vp, i32, vp, vp, C.POINTER(SlcParityPlanes), vp]
    L.slc_reconstruct_host.argtypes = [vp, vp, i32, vp, vp, C.POINTER(SlcParityPlanes)]
    L.slc_submit_host.argtypes = [vp, i32, vp, i32, vp, vp]
    L.slc_wait.argtypes = [vp, i32]
    L.slc_reconstruct_device_ex.argtypes = [vp, vp, i32, C.POINTER(SlcResult), vp]
    L.slc_reconstruct_host_ex.argtypes = [vp, vp, i32, C.POINTER(SlcResult)]
    L.slc_pool_create.argtypes = [C.POINTER(SlcConfig), C.POINTER(i32), i32, C.POINTER(vp)]
    L.slc_pool_destroy.argtypes = [vp]
    L.slc_pool_destroy.restype = None
    L.slc_pool_last_error.argtypes = [vp]
    L.slc_pool_last_error.restype = C.c_char_p
    L.slc_pool_size.argtypes = [vp]
    L.slc_pool_context.argtypes = [vp, i32]
    L.slc_pool_context.restype = vp
    L.slc_pool_set_calibration.argtypes = [vp, vp, vp, vp, vp]
    L.slc_pool_set_gray_lut.argtypes = [vp, vp, i32]
    L.slc_shard_range.argtypes = [C.c_int64, i32, i32, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.slc_pool_reconstruct_host.argtypes = [vp, vp, i32, C.POINTER(SlcResult)]
    L.slc_pool_last_shares.argtypes = [vp, C.POINTER(i32), i32]
    L.slc_submit_host_ex.argtypes = [vp, i32, vp, i32, C.POINTER(SlcResult)]
    L.slc_pool_reconstruct_device.argtypes = [vp, C.POINTER(vp), C.POINTER(i32), C.POINTER(SlcResult)]
    L.slc_decode_gray_host.argtypes = [vp, vp, vp, vp]
    L.slc_decode_phase_host.argtypes = [vp, vp, vp, vp]
    L.slc_triangulate_host.argtypes = [vp, vp, vp, vp]
    L.slc_triangulate_uv_device.argtypes = [vp, vp, vp, vp, vp, vp]
    L.slc_triangulate_uv_host.argtypes = [vp, vp, vp, vp, vp]
    L.slc_dyna_track_device.argtypes = [vp, vp, i32, i32, vp, vp, vp, vp, C.POINTER(SlcDynaParity), vp]
    L.slc_dyna_track_host.argtypes = [vp, vp, i32, i32, vp, vp, vp, vp, C.POINTER(SlcDynaParity)]
    L.slc_dyna_track_host_ex.argtypes = [vp, vp, i32, i32, vp, C.POINTER(SlcResult)]
    L.slc_eval_phase_host.argtypes = [vp, vp, vp, C.c_int64, vp, vp]
    i64p = C.POINTER(C.c_int64)
    L.slc_bmp_parse.argtypes = [vp, C.c_int64, C.POINTER(SlcBmpInfo)]
    L.slc_bmp_unpack_device.argtypes = [vp, vp, C.POINTER(SlcBmpInfo), vp, vp]
    L.slc_bmp_unpack_batch_device.argtypes = [vp, C.POINTER(vp), C.POINTER(SlcBmpInfo), i32, vp, vp]
    L.slc_bmp_decode_host.argtypes = [vp, vp, C.c_int64, vp, i32, i32]
    L.slc_load_bmp_planes.argtypes = [vp, C.POINTER(C.c_char_p), i32, vp]
    L.slc_pointcloud_text_device.argtypes = [vp, vp, C.c_uint32, vp, C.c_int64, i64p, i64p, vp]
    L.slc_pointcloud_text_host.argtypes = [vp, vp, C.c_uint32, vp, C.c_int64, i64p, i64p]
    L.slc_pointcloud_compact_device.argtypes = [vp, vp, vp, i32, vp, C.c_int64, i64p, vp]
    L.slc_pointcloud_compact_host.argtypes = [vp, vp, vp, i32, vp, C.c_int64, i64p]
    L.slc_compact_points_device.argtypes = [vp, vp, vp, i32, i32, vp, C.c_int64, vp, vp, vp]
    L.slc_format_g6_host.argtypes = [vp, vp, C.c_int64, C.c_uint32, vp, vp]
    L.slc_time_reconstruct_device.argtypes = [vp, vp, i32, vp, vp, i32, C.POINTER(C.c_float)]
    L.slc_launch_count.argtypes = [vp]
    L.slc_launch_count.restype = C.c_int64
    L.slc_set_pixels_per_thread.argtypes = [vp, i32]
    _lib = L
    return L


def bmp_parse(file_bytes: bytes) -> SlcBmpInfo:
    """Header + palette of a BMP (host only; raises SlcError for unsupported flavours)."""
    info = SlcBmpInfo()
    st = load_library().slc_bmp_parse(file_bytes, len(file_bytes), C.byref(info))
    if st != SLC_OK:
        raise SlcError(st, "not an uncompressed 8/24/32-bit BMP (or truncated)")
    return info


class PinnedArray:
    """numpy view over pinned host memory from slc_host_alloc."""

    def __init__(self, shape, dtype, flags: int = 0):
        self._lib = load_library()
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        self.ptr = self._lib.slc_host_alloc_ex(max(nbytes, 1), flags)
        if not self.ptr:
            raise MemoryError(f"slc_host_alloc_ex({nbytes}, {flags}) failed")
        buf = (C.c_uint8 * max(nbytes, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def free(self):
        if self.ptr:
            self.array = None
            self._lib.slc_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _nbytes(a) -> int:
    if isinstance(a, PinnedArray):
        return int(np.prod(a.shape)) * a.dtype.itemsize
    return int(a.nbytes)


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, PinnedArray):
        return a.ptr
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data
    return int(a)


def shard_range(n_items: int, index: int, n_shards: int) -> tuple[int, int]:
    """slc_shard_range: the contiguous block of frame sets member `index` of `n_shards` owns."""
    lo, hi = C.c_int64(0), C.c_int64(0)
    st = load_library().slc_shard_range(n_items, index, n_shards, C.byref(lo), C.byref(hi))
    if st != SLC_OK:
        raise SlcError(st, f"bad shard request n={n_items} index={index} shards={n_shards}")
    return int(lo.value), int(hi.value)


def bits_bytes(npx: int) -> int:
    return (npx + 7) // 8


def make_result(fmt: int, order: int = SLC_ORDER_ROW_MAJOR, xyzw=None, mask=None, depth=None, mask_bits=None,
                points=None, point_stride: int = 0, n_points=None) -> SlcResult:
    """slc_result from numpy arrays / PinnedArrays / raw addresses (None = not wanted)."""
    return SlcResult(fmt, order, _ptr(xyzw), _ptr(mask), _ptr(depth), _ptr(mask_bits), _ptr(points), point_stride,
                     _ptr(n_points))


def alloc_result(cfg: StackConfig, n: int, fmt: int, order: int = SLC_ORDER_ROW_MAJOR, pinned: bool = False,
                 point_stride: int | None = None, full_maps: bool = False):
    """Host buffers for `n` frame sets in `fmt` -> (dict of arrays / PinnedArrays, SlcResult)."""
    H, W, npx = cfg.height, cfg.width, cfg.pixels
    mk = (lambda shape, dt: PinnedArray(shape, dt)) if pinned else (lambda shape, dt: np.empty(shape, dt))
    bufs = {}
    if fmt == SLC_RESULT_XYZW or full_maps:
        bufs["xyzw"] = mk((n, H, W, 4), np.float32)
        bufs["mask"] = mk((n, H, W), np.uint8)
    if fmt == SLC_RESULT_DEPTH:
        bufs["depth"] = mk((n, H, W), np.float32)
    if fmt in (SLC_RESULT_DEPTH, SLC_RESULT_POINTS):
        bufs["mask_bits"] = mk(((n * bits_bytes(npx) + 3) // 4 * 4,), np.uint8)
    stride = 0
    if fmt == SLC_RESULT_POINTS:
        stride = npx if point_stride is None else point_stride
        bufs["points"] = mk((n, stride, 3), np.float32)
        bufs["n_points"] = mk((n,), np.int64)
    res = make_result(fmt, order, bufs.get("xyzw"), bufs.get("mask"), bufs.get("depth"), bufs.get("mask_bits"),
                      bufs.get("points"), stride, bufs.get("n_points"))
    return bufs, res


def _arr(a):
    return a.array if isinstance(a, PinnedArray) else a


def unpack_mask_bits(bits, n: int, npx: int) -> np.ndarray:
    """mask_bits [n][(npx+7)/8] -> u8 [n][npx] of 0/1."""
    b = np.asarray(_arr(bits))[: n * bits_bytes(npx)].reshape(n, bits_bytes(npx))
    return np.unpackbits(b, axis=1, bitorder="little")[:, :npx]


class Reconstructor:
    """One slc_context: the fused decode -> unwrap -> triangulate path on one GPU."""

    def __init__(self, cfg: StackConfig, device: int = 0, max_batch: int = 1, num_slots: int = 2, flags: int = 0):
        self.lib = load_library()
        self.cfg = cfg
        self.max_batch = max_batch
        c = SlcConfig(cfg.width, cfg.height, cfg.projector_width, cfg.gray_digits, cfg.phase_steps,
                      float(cfg.fov_min), float(cfg.fov_max), float(cfg.modulation_min), flags, device,
                      max_batch, num_slots)
        h = C.c_void_p()
        st = self.lib.slc_create(C.byref(c), C.byref(h))
        if st != SLC_OK:
            raise SlcError(st, (self.lib.slc_last_error(None) or b"").decode())
        self.h = h

    # -- helpers ---------------------------------------------------------
    def _check(self, st: int):
        if st != SLC_OK:
            raise SlcError(st, (self.lib.slc_last_error(self.h) or b"").decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.slc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self) -> SlcInfo:
        out = SlcInfo()
        self._check(self.lib.slc_get_info(self.h, C.byref(out)))
        return out

    def set_calibration(self, cal):
        cam = np.ascontiguousarray(cal.cam, dtype=np.float64).reshape(9)
        pro = np.ascontiguousarray(cal.pro, dtype=np.float64).reshape(9)
        R = np.ascontiguousarray(cal.R, dtype=np.float64).reshape(9)
        T = np.ascontiguousarray(cal.T, dtype=np.float64).reshape(3)
        self._check(self.lib.slc_set_calibration(self.h, cam.ctypes.data, pro.ctypes.data, R.ctypes.data,
                                                 T.ctypes.data))

    def set_gray_lut(self, lut: np.ndarray):
        lut = np.ascontiguousarray(lut, dtype=np.int16)
        self._check(self.lib.slc_set_gray_lut(self.h, lut.ctypes.data, int(lut.size)))

    # -- device memory ---------------------------------------------------
    def device_alloc(self, nbytes: int) -> int:
        p = self.lib.slc_device_alloc(self.h, nbytes)
        if not p:
            raise SlcError(SLC_ERR_OUT_OF_MEMORY, (self.lib.slc_last_error(self.h) or b"").decode())
        return p

    def device_free(self, p: int):
        self.lib.slc_device_free(self.h, p)

    def to_device(self, dptr: int, host, nbytes: int | None = None):
        if nbytes is None:
            nbytes = host.array.nbytes if isinstance(host, PinnedArray) else host.nbytes
        self._check(self.lib.slc_copy_to_device(self.h, dptr, _ptr(host), nbytes))

    def to_host(self, host, dptr: int, nbytes: int | None = None):
        if nbytes is None:
            nbytes = host.array.nbytes if isinstance(host, PinnedArray) else host.nbytes
        self._check(self.lib.slc_copy_to_host(self.h, _ptr(host), dptr, nbytes))

    def synchronize(self):
        self._check(self.lib.slc_synchronize(self.h))

    # -- hot path --------------------------------------------------------
    def reconstruct_device(self, d_stack: int, n_stacks: int, d_xyzw: int, d_mask: int, d_parity=None,
                           stream: int | None = None):
        par = C.byref(d_parity) if d_parity is not None else None
        self._check(self.lib.slc_reconstruct_device(self.h, d_stack, n_stacks, d_xyzw, d_mask, par, stream))

    def reconstruct(self, stacks, parity: bool = False) -> dict:
        """Host buffers in, host buffers out (numpy or PinnedArray): the public call."""
        arr = stacks.array if isinstance(stacks, PinnedArray) else np.ascontiguousarray(stacks, dtype=np.uint8)
        cfg = self.cfg
        if arr.ndim == 3:
            arr = arr[None]
        n = arr.shape[0]
        assert arr.shape[1:] == (cfg.planes, cfg.height, cfg.width), arr.shape
        out = {
            "xyzw": np.empty((n, cfg.height, cfg.width, 4), np.float32),
            "mask": np.empty((n, cfg.height, cfg.width), np.uint8),
        }
        par = None
        if parity:
            out["kbin"] = np.empty((n, cfg.height, cfg.width), np.int16)
            out["corr"] = np.empty((n, cfg.height, cfg.width), np.int8)
            out["phase_pix"] = np.empty((n, cfg.height, cfg.width), np.float32)
            out["proj_u"] = np.empty((n, cfg.height, cfg.width), np.float64)
            par = SlcParityPlanes(out["kbin"].ctypes.data, out["corr"].ctypes.data,
                                  out["phase_pix"].ctypes.data, out["proj_u"].ctypes.data)
        self._check(self.lib.slc_reconstruct_host(self.h, arr.ctypes.data, n, out["xyzw"].ctypes.data,
                                                  out["mask"].ctypes.data, C.byref(par) if par else None))
        return out

    def reconstruct_into(self, h_stack, n_stacks: int, h_xyzw, h_mask):
        """Host path into caller-provided (ideally pinned) buffers."""
        self._check(self.lib.slc_reconstruct_host(self.h, _ptr(h_stack), n_stacks, _ptr(h_xyzw), _ptr(h_mask), None))

    def reconstruct_into_ex(self, h_stack, n_stacks: int, result: SlcResult):
        """Host path with a result format (slc_reconstruct_host_ex)."""
        self._check(self.lib.slc_reconstruct_host_ex(self.h, _ptr(h_stack), n_stacks, C.byref(result)))

    def reconstruct_device_ex(self, d_stack: int, n_stacks: int, result: SlcResult, stream: int | None = None):
        self._check(self.lib.slc_reconstruct_device_ex(self.h, d_stack, n_stacks, C.byref(result), stream))

    def reconstruct_ex(self, stacks, fmt: int, order: int = SLC_ORDER_ROW_MAJOR, full_maps: bool = False) -> dict:
        """numpy in, numpy out, in result format `fmt`."""
        arr = np.ascontiguousarray(_arr(stacks), dtype=np.uint8)
        if arr.ndim == 3:
            arr = arr[None]
        bufs, res = alloc_result(self.cfg, arr.shape[0], fmt, order, full_maps=full_maps)
        self.reconstruct_into_ex(arr, arr.shape[0], res)
        return bufs

    def set_pixels_per_thread(self, pxt: int):
        self._check(self.lib.slc_set_pixels_per_thread(self.h, pxt))

    def submit(self, slot: int, h_stack, n_stacks: int, h_xyzw, h_mask):
        self._check(self.lib.slc_submit_host(self.h, slot, _ptr(h_stack), n_stacks, _ptr(h_xyzw), _ptr(h_mask)))

    def submit_ex(self, slot: int, h_stack, n_stacks: int, result: SlcResult):
        self._check(self.lib.slc_submit_host_ex(self.h, slot, _ptr(h_stack), n_stacks, C.byref(result)))

    def wait(self, slot: int):
        self._check(self.lib.slc_wait(self.h, slot))

    # -- decoder objects ---------------------------------------------------
    def decode_gray(self, gray_planes: np.ndarray):
        cfg = self.cfg
        gray_planes = np.ascontiguousarray(gray_planes, dtype=np.uint8)
        assert gray_planes.shape == (2 * cfg.gray_digits, cfg.height, cfg.width)
        val = np.empty((cfg.height, cfg.width), np.float64)
        kbin = np.empty((cfg.height, cfg.width), np.int16)
        self._check(self.lib.slc_decode_gray_host(self.h, gray_planes.ctypes.data, val.ctypes.data, kbin.ctypes.data))
        return val, kbin

    def decode_phase(self, phase_planes: np.ndarray):
        cfg = self.cfg
        phase_planes = np.ascontiguousarray(phase_planes, dtype=np.uint8)
        assert phase_planes.shape == (cfg.phase_steps, cfg.height, cfg.width)
        pix = np.empty((cfg.height, cfg.width), np.float64)
        mod = np.empty((cfg.height, cfg.width), np.uint8)
        self._check(self.lib.slc_decode_phase_host(self.h, phase_planes.ctypes.data, pix.ctypes.data, mod.ctypes.data))
        return pix, mod

    def triangulate(self, proj_u: np.ndarray):
        cfg = self.cfg
        proj_u = np.ascontiguousarray(proj_u, dtype=np.float64)
        assert proj_u.shape == (cfg.height, cfg.width)
        xyzw = np.empty((cfg.height, cfg.width, 4), np.float32)
        mask = np.empty((cfg.height, cfg.width), np.uint8)
        self._check(self.lib.slc_triangulate_host(self.h, proj_u.ctypes.data, xyzw.ctypes.data, mask.ctypes.data))
        return xyzw, mask

    def triangulate_uv(self, proj_u: np.ndarray, proj_v: np.ndarray):
        """[EXT] least-squares triangulation from the projector column and row planes (f64)."""
        cfg = self.cfg
        proj_u = np.ascontiguousarray(proj_u, dtype=np.float64)
        proj_v = np.ascontiguousarray(proj_v, dtype=np.float64)
        assert proj_u.shape == (cfg.height, cfg.width) and proj_v.shape == proj_u.shape
        xyzw = np.empty((cfg.height, cfg.width, 4), np.float32)
        mask = np.empty((cfg.height, cfg.width), np.uint8)
        self._check(self.lib.slc_triangulate_uv_host(self.h, proj_u.ctypes.data, proj_v.ctypes.data, xyzw.ctypes.data,
                                                     mask.ctypes.data))
        return xyzw, mask

    # -- dynamic frames -----------------------------------------------------
    def dyna_track(self, frames: np.ndarray, u0: np.ndarray, window: int = 21, parity: bool = False) -> dict:
        """CalculateOther over frames[1:] (frames[0] = the image StripRegression(0) saw)."""
        cfg = self.cfg
        frames = np.ascontiguousarray(frames, dtype=np.uint8)
        u0 = np.ascontiguousarray(u0, dtype=np.float64)
        n = frames.shape[0]
        assert frames.shape[1:] == (cfg.height, cfg.width) and u0.shape == (cfg.height, cfg.width)
        out = {
            "xyzw": np.empty((n - 1, cfg.height, cfg.width, 4), np.float32),
            "mask": np.empty((n - 1, cfg.height, cfg.width), np.uint8),
            "delta_z": np.empty((n - 1, cfg.height, cfg.width), np.float32),
        }
        par = None
        if parity:
            out["strips"] = np.empty((n, cfg.height, cfg.width, 2), np.int8)
            out["delta_p"] = np.empty((n - 1, cfg.height, cfg.width), np.float32)
            out["proj_u"] = np.empty((n - 1, cfg.height, cfg.width), np.float64)
            par = SlcDynaParity(out["strips"].ctypes.data, out["delta_p"].ctypes.data, out["proj_u"].ctypes.data)
        self._check(self.lib.slc_dyna_track_host(self.h, frames.ctypes.data, n, window, u0.ctypes.data,
                                                 out["xyzw"].ctypes.data, out["mask"].ctypes.data,
                                                 out["delta_z"].ctypes.data, C.byref(par) if par else None))
        return out

    def dyna_track_into(self, h_frames, n_frames: int, h_u0, h_xyzw, h_mask, h_delta_z=None, window: int = 21):
        """CalculateOther with caller-owned (ideally pinned) host buffers: frames u8 [n][H][W], u0 f64 [H][W] in;
        xyzw f32 [n-1][H][W][4], mask u8 [n-1][H][W], delta_z f32 [n-1][H][W] out."""
        self._check(self.lib.slc_dyna_track_host(self.h, _ptr(h_frames), n_frames, window, _ptr(h_u0), _ptr(h_xyzw),
                                                 _ptr(h_mask), _ptr(h_delta_z), None))

    def dyna_track_into_ex(self, h_frames, n_frames: int, h_u0, result: SlcResult, window: int = 21):
        """CalculateOther with a result format for the n_frames - 1 maps (slc_dyna_track_host_ex)."""
        self._check(self.lib.slc_dyna_track_host_ex(self.h, _ptr(h_frames), n_frames, window, _ptr(h_u0), C.byref(result)))

    def dyna_track_device(self, d_frames: int, n_frames: int, d_u0: int, d_xyzw: int, d_mask: int,
                          d_delta_z: int | None = None, window: int = 21, stream: int | None = None):
        self._check(self.lib.slc_dyna_track_device(self.h, d_frames, n_frames, window, d_u0, d_xyzw, d_mask,
                                                   d_delta_z, None, stream))

    # -- input ingest -----------------------------------------------------------
    def bmp_decode(self, file_bytes: bytes, expect_shape=None) -> np.ndarray:
        """imread(..., GRAYSCALE) of a BMP file image, pixel work on the device."""
        info = bmp_parse(file_bytes)
        out = np.empty((info.height, info.width), np.uint8)
        eh, ew = expect_shape if expect_shape else (0, 0)
        self._check(self.lib.slc_bmp_decode_host(self.h, file_bytes, len(file_bytes), out.ctypes.data, ew, eh))
        return out

    def bmp_unpack_batch_device(self, d_pixels, infos, d_stack: int, stream: int = 0):
        """Pixel arrays already on the device (d_pixels: device addresses, infos: SlcBmpInfo of each file)
        -> planes 0..n-1 of the u8 [n][H][W] stack at d_stack, one launch (slc_bmp_unpack_batch_device)."""
        n = len(d_pixels)
        ptrs = (C.c_void_p * n)(*[int(p) for p in d_pixels])
        arr = (SlcBmpInfo * n)(*infos)
        self._check(self.lib.slc_bmp_unpack_batch_device(self.h, ptrs, arr, n, d_stack, stream))

    def load_bmp_planes(self, paths, d_stack: int):
        arr = (C.c_char_p * len(paths))(*[os.fsencode(p) for p in paths])
        self._check(self.lib.slc_load_bmp_planes(self.h, arr, len(paths), d_stack))

    # -- point-cloud output ----------------------------------------------------
    def pointcloud_text(self, proj_u: np.ndarray, flags: int = 0, capacity: int | None = None):
        """CCalculation::Result's text for one f64 ProjectorU plane -> (bytes, n_points)."""
        cfg = self.cfg
        proj_u = np.ascontiguousarray(proj_u, dtype=np.float64)
        assert proj_u.shape == (cfg.height, cfg.width)
        cap = 43 * proj_u.size + 16 if capacity is None else capacity
        buf = np.empty(max(cap, 1), np.uint8)
        nb, npts = self.pointcloud_text_into(proj_u, buf, flags, cap)
        return buf[:nb].tobytes(), npts

    def pointcloud_text_into(self, proj_u, text_buf, flags: int = 0, capacity: int | None = None):
        """Same, into a caller-owned (ideally pinned) byte buffer -> (n_bytes, n_points)."""
        cap = int(_nbytes(text_buf)) if capacity is None else capacity
        nb, npts = C.c_int64(0), C.c_int64(0)
        self._check(self.lib.slc_pointcloud_text_host(self.h, _ptr(proj_u), flags, _ptr(text_buf), cap,
                                                      C.byref(nb), C.byref(npts)))
        return int(nb.value), int(npts.value)

    def pointcloud_text_device(self, d_proj_u: int, d_text: int, capacity: int, flags: int = 0,
                               stream: int | None = None):
        nb, npts = C.c_int64(0), C.c_int64(0)
        self._check(self.lib.slc_pointcloud_text_device(self.h, d_proj_u, flags, d_text, capacity, C.byref(nb),
                                                        C.byref(npts), stream))
        return int(nb.value), int(npts.value)

    def pointcloud_compact(self, xyzw: np.ndarray, mask: np.ndarray, order: int = SLC_ORDER_ROW_MAJOR) -> np.ndarray:
        """float32 [n_points][3] of the valid pixels of one (xyzw, mask) map."""
        cfg = self.cfg
        xyzw = np.ascontiguousarray(xyzw, dtype=np.float32)
        mask = np.ascontiguousarray(mask, dtype=np.uint8)
        assert xyzw.shape == (cfg.height, cfg.width, 4) and mask.shape == (cfg.height, cfg.width)
        out = np.empty((mask.size, 3), np.float32)
        n = C.c_int64(0)
        self._check(self.lib.slc_pointcloud_compact_host(self.h, xyzw.ctypes.data, mask.ctypes.data, order,
                                                         out.ctypes.data, mask.size, C.byref(n)))
        return out[: n.value].copy()

    def pointcloud_compact_device(self, d_xyzw: int, d_mask: int, d_xyz: int, capacity_points: int,
                                  order: int = SLC_ORDER_ROW_MAJOR, stream: int | None = None) -> int:
        n = C.c_int64(0)
        self._check(self.lib.slc_pointcloud_compact_device(self.h, d_xyzw, d_mask, order, d_xyz, capacity_points,
                                                           C.byref(n), stream))
        return int(n.value)

    def compact_points_device(self, d_xyzw: int, d_mask: int, n_maps: int, d_points: int, point_stride: int,
                              d_n_points: int, order: int = SLC_ORDER_ROW_MAJOR, d_mask_bits: int | None = None,
                              stream: int | None = None):
        """n_maps maps -> packed point lists in one asynchronous launch; everything stays on the device."""
        self._check(self.lib.slc_compact_points_device(self.h, d_xyzw, d_mask, n_maps, order, d_points, point_stride,
                                                       d_mask_bits, d_n_points, stream))

    def format_g6(self, values: np.ndarray, flags: int = 0) -> list[bytes]:
        """Device number formatting (printf "%g") of each value."""
        v = np.ascontiguousarray(values, dtype=np.float64).reshape(-1)
        text = np.zeros((v.size, 16), np.uint8)
        ln = np.zeros(v.size, np.uint8)
        self._check(self.lib.slc_format_g6_host(self.h, v.ctypes.data, v.size, flags, text.ctypes.data, ln.ctypes.data))
        return [text[i, : ln[i]].tobytes() for i in range(v.size)]

    def eval_phase(self, s: np.ndarray, c: np.ndarray):
        """Device cvFastArctan + in-period offset for arbitrary (sin, cos) sums."""
        s = np.ascontiguousarray(s, dtype=np.float32).reshape(-1)
        c = np.ascontiguousarray(c, dtype=np.float32).reshape(-1)
        assert s.size == c.size
        deg = np.empty(s.size, np.float32)
        pix = np.empty(s.size, np.float32)
        self._check(self.lib.slc_eval_phase_host(self.h, s.ctypes.data, c.ctypes.data, s.size, deg.ctypes.data,
                                                 pix.ctypes.data))
        return deg, pix

    # -- measurement -------------------------------------------------------
    def time_device(self, d_stack: int, n_stacks: int, d_xyzw: int, d_mask: int, iters: int) -> float:
        ms = C.c_float()
        self._check(self.lib.slc_time_reconstruct_device(self.h, d_stack, n_stacks, d_xyzw, d_mask, iters,
                                                         C.byref(ms)))
        return float(ms.value)

    def launch_count(self) -> int:
        return int(self.lib.slc_launch_count(self.h))


class Pool:
    """slc_pool: one context + one feeder thread per listed GPU behind one call."""

    def __init__(self, cfg: StackConfig, devices, max_batch: int = 2, num_slots: int = 4, flags: int = 0):
        self.lib = load_library()
        self.cfg = cfg
        c = SlcConfig(cfg.width, cfg.height, cfg.projector_width, cfg.gray_digits, cfg.phase_steps,
                      float(cfg.fov_min), float(cfg.fov_max), float(cfg.modulation_min), flags, 0, max_batch, num_slots)
        devs = (C.c_int32 * len(devices))(*devices)
        h = C.c_void_p()
        st = self.lib.slc_pool_create(C.byref(c), devs, len(devices), C.byref(h))
        if st != SLC_OK:
            raise SlcError(st, (self.lib.slc_pool_last_error(None) or b"").decode())
        self.h = h
        self.n = len(devices)

    def _check(self, st: int):
        if st != SLC_OK:
            raise SlcError(st, (self.lib.slc_pool_last_error(self.h) or b"").decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.slc_pool_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_calibration(self, cal):
        cam = np.ascontiguousarray(cal.cam, dtype=np.float64).reshape(9)
        pro = np.ascontiguousarray(cal.pro, dtype=np.float64).reshape(9)
        R = np.ascontiguousarray(cal.R, dtype=np.float64).reshape(9)
        T = np.ascontiguousarray(cal.T, dtype=np.float64).reshape(3)
        self._check(self.lib.slc_pool_set_calibration(self.h, cam.ctypes.data, pro.ctypes.data, R.ctypes.data,
                                                      T.ctypes.data))

    def context(self, member: int) -> int:
        return self.lib.slc_pool_context(self.h, member)

    def launch_count(self) -> int:
        return sum(int(self.lib.slc_launch_count(self.context(i))) for i in range(self.n))

    def reconstruct_into_ex(self, h_stack, n_stacks: int, result: SlcResult):
        self._check(self.lib.slc_pool_reconstruct_host(self.h, _ptr(h_stack), n_stacks, C.byref(result)))

    def last_shares(self) -> list[int]:
        """Frame sets each member took in the last reconstruct_into_ex call (handed out on demand)."""
        out = (C.c_int32 * self.n)()
        self._check(self.lib.slc_pool_last_shares(self.h, out, self.n))
        return list(out)

    def reconstruct_device(self, d_stacks, n_stacks, results):
        """Per-member device shards: d_stacks[i] / results[i] live on member i's GPU."""
        ptrs = (C.c_void_p * self.n)(*[int(p) for p in d_stacks])
        ns = (C.c_int32 * self.n)(*n_stacks)
        res = (SlcResult * self.n)(*results)
        self._check(self.lib.slc_pool_reconstruct_device(self.h, ptrs, ns, res))
