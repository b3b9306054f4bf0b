"""vkresample_b200 -- Python host binding of libb2resample.so (ctypes over the C ABI in
include/b2resample.h).

This mirrors the call sequence of the reference's ``launchResample`` (VkResample.cpp:1280-1780):
``Plan(...)`` = configuration + buffer allocation + plan build, ``upload`` =
``transferDataFromCPU`` (:1688), ``execute`` = ``performVulkanUpscale`` (:1692), ``download`` =
``transferDataToCPU`` (:1697-1700).  All arithmetic happens in the CUDA library; there is no CPU
fallback -- importing works anywhere, but every compute call raises ``B2RError`` when the
library or a CUDA device is missing.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import numpy as np

__all__ = ["Plan", "B2RError", "library_path", "load_library", "device_count", "device_name", "EXPORTS"]

_HERE = os.path.dirname(os.path.abspath(__file__))
PRECISION_FP32 = 0
PRECISION_FP64 = 1
PRECISION_FP16 = 2
FLAG_NO_GRAPH = 1
FLAG_NO_SHARPEN_LITERAL_ROUNDING = 2
FLAG_C2C_PARITY = 4
FLAG_JIT = 8
FLAG_NO_JIT = 16
FLAG_FAST_SHARPEN = 32   # round-1 opt-in, now a no-op on its own (the default sharpen is the tolerance-bound one)
FLAG_EXACT_SHARPEN = 64  # bit-exact (oracle-identical) sharpen kernels, see b2resample.h
FLAG_SEPARATE_SHARPEN = 128  # C2R rows and sharpen as two kernels (default: one fused kernel where it applies)

# every symbol include/b2resample.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "b2r_device_count", "b2r_device_name", "b2r_plan_create", "b2r_plan_destroy", "b2r_plan_input_bytes",
    "b2r_plan_output_bytes", "b2r_plan_get_info", "b2r_upload", "b2r_execute", "b2r_download",
    "b2r_upscale_host", "b2r_device_input", "b2r_device_output", "b2r_download_pre_sharpen",
    "b2r_plan_pre_sharpen_bytes", "b2r_sharpen_host", "b2r_synchronize", "b2r_plan_stream",
    "b2r_plan_launch_count", "b2r_last_error", "b2r_version", "b2r_enqueue_device", "b2r_timer_start",
    "b2r_timer_stop", "b2r_profile_kernels", "b2r_enqueue_host", "b2r_plan_set_lanes", "b2r_plan_lanes",
    "b2r_plan_input_u8_bytes", "b2r_plan_output_u8_bytes", "b2r_upload_u8", "b2r_download_u8", "b2r_enqueue_host_u8",
    "b2r_plan_last_ticket", "b2r_wait_ticket", "b2r_host_alloc", "b2r_host_free",
]


class B2RError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"b2resample error {code}: {msg}")
        self.code = code


class PlanInfo(ctypes.Structure):
    _fields_ = [
        ("w", ctypes.c_uint32), ("h", ctypes.c_uint32), ("up_w", ctypes.c_uint32), ("up_h", ctypes.c_uint32),
        ("precision", ctypes.c_uint32), ("upscale", ctypes.c_float), ("sharpen", ctypes.c_float),
        ("zeropad_lo_y", ctypes.c_uint32), ("zeropad_hi_y", ctypes.c_uint32),
        ("spectrum_row_stride", ctypes.c_uint32),
        ("input_bytes", ctypes.c_size_t), ("output_bytes", ctypes.c_size_t), ("device_bytes", ctypes.c_size_t),
        ("n_stages", ctypes.c_uint32 * 4), ("radices", (ctypes.c_uint32 * 8) * 4), ("threads", ctypes.c_uint32 * 4),
        ("column_tile", ctypes.c_uint32), ("kernels_per_frame", ctypes.c_uint32), ("static_kernels", ctypes.c_uint32),
        ("jit_kernels", ctypes.c_uint32), ("jit_note", ctypes.c_char * 128),
        ("c2c_mode", ctypes.c_uint32), ("pre_sharpen_plane_stride", ctypes.c_size_t),
        ("fused_strips_per_plane", ctypes.c_uint32), ("sharpen_mode", ctypes.c_uint32),
    ]


def library_path() -> str:
    return os.environ.get("B2R_LIBRARY", os.path.join(_HERE, "lib", "libb2resample.so"))


_LIB = None


def load_library():
    """dlopen libb2resample.so and declare the prototypes.  Raises B2RError if it is not built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise B2RError(-2, f"{path} not found: build it with `make` (nvcc, sm_100a); there is no CPU fallback")
    L = ctypes.CDLL(path)
    vp, u32, f32 = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_float
    L.b2r_device_count.restype = ctypes.c_int
    L.b2r_device_name.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t]
    L.b2r_plan_create.argtypes = [ctypes.POINTER(vp), ctypes.c_int, u32, u32, f32, u32, f32, u32]
    L.b2r_plan_destroy.argtypes = [vp]
    L.b2r_plan_destroy.restype = None
    for name in ("b2r_plan_input_bytes", "b2r_plan_output_bytes", "b2r_plan_pre_sharpen_bytes",
                 "b2r_plan_input_u8_bytes", "b2r_plan_output_u8_bytes"):
        getattr(L, name).argtypes = [vp]
        getattr(L, name).restype = ctypes.c_size_t
    L.b2r_plan_get_info.argtypes = [vp, ctypes.POINTER(PlanInfo)]
    L.b2r_upload.argtypes = [vp, vp]
    L.b2r_execute.argtypes = [vp, u32, ctypes.POINTER(ctypes.c_double)]
    L.b2r_download.argtypes = [vp, vp]
    L.b2r_upscale_host.argtypes = [vp, vp, vp, ctypes.POINTER(ctypes.c_double)]
    for name in ("b2r_device_input", "b2r_device_output", "b2r_plan_stream"):
        getattr(L, name).argtypes = [vp]
        getattr(L, name).restype = vp
    L.b2r_download_pre_sharpen.argtypes = [vp, vp]
    L.b2r_sharpen_host.argtypes = [vp, vp, vp]
    L.b2r_synchronize.argtypes = [vp]
    L.b2r_enqueue_device.argtypes = [vp, vp, vp]
    L.b2r_enqueue_host.argtypes = [vp, vp, vp]
    L.b2r_upload_u8.argtypes = [vp, vp]
    L.b2r_download_u8.argtypes = [vp, vp]
    L.b2r_enqueue_host_u8.argtypes = [vp, vp, vp]
    L.b2r_plan_last_ticket.argtypes = [vp]
    L.b2r_plan_last_ticket.restype = ctypes.c_uint64
    L.b2r_wait_ticket.argtypes = [vp, ctypes.c_uint64]
    L.b2r_host_alloc.argtypes = [ctypes.c_size_t]
    L.b2r_host_alloc.restype = vp
    L.b2r_host_free.argtypes = [vp]
    L.b2r_host_free.restype = None
    L.b2r_plan_set_lanes.argtypes = [vp, u32]
    L.b2r_plan_lanes.argtypes = [vp]
    L.b2r_plan_lanes.restype = u32
    L.b2r_timer_start.argtypes = [vp]
    L.b2r_timer_stop.argtypes = [vp, ctypes.POINTER(ctypes.c_double)]
    L.b2r_profile_kernels.argtypes = [vp, u32, ctypes.POINTER(ctypes.c_double)]
    L.b2r_plan_launch_count.argtypes = [vp]
    L.b2r_plan_launch_count.restype = ctypes.c_uint64
    L.b2r_last_error.restype = ctypes.c_char_p
    L.b2r_version.restype = ctypes.c_char_p
    _LIB = L
    return L


def _check(rc: int):
    if rc != 0:
        raise B2RError(rc, load_library().b2r_last_error().decode())


def device_count() -> int:
    return int(load_library().b2r_device_count())


def device_name(device: int = 0) -> str:
    buf = ctypes.create_string_buffer(256)
    _check(load_library().b2r_device_name(device, buf, 256))
    return buf.value.decode()


def _ptr(a) -> int:
    """host pointer of a numpy array or a raw integer address"""
    if isinstance(a, np.ndarray):
        if not a.flags["C_CONTIGUOUS"]:
            raise ValueError("host buffers must be C-contiguous")
        return a.ctypes.data
    return int(a)


class Plan:
    """One upscale plan (fixed W, H, factor, precision, sharpen) bound to one GPU.

    Reusable for every frame of that size, like the per-thread VkFFT applications the reference
    builds once in launchResample (VkResample.cpp:1506-1617)."""

    def __init__(self, w: int, h: int, upscale: float = 2.0, precision: int = 0, sharpen: float = 0.2,
                 device: int = 0, flags: int = 0):
        self._lib = load_library()
        self._h = ctypes.c_void_p()
        _check(self._lib.b2r_plan_create(ctypes.byref(self._h), device, w, h, upscale, precision, sharpen, flags))
        info = PlanInfo()
        _check(self._lib.b2r_plan_get_info(self._h, ctypes.byref(info)))
        self.info = info
        self.w, self.h, self.up_w, self.up_h = info.w, info.h, info.up_w, info.up_h
        self.precision = info.precision
        self.dtype = {0: np.float32, 1: np.float64, 2: np.float16}[int(info.precision)]
        self.input_bytes, self.output_bytes = info.input_bytes, info.output_bytes
        self.device = device

    def get_info(self) -> "PlanInfo":
        """a fresh b2r_plan_get_info snapshot (``self.info`` is the one taken at creation)"""
        info = PlanInfo()
        _check(self._lib.b2r_plan_get_info(self._h, ctypes.byref(info)))
        return info

    # -- layouts -----------------------------------------------------------------------------
    @property
    def in_plane_stride(self) -> int:
        return (self.w + 2) * self.h

    @property
    def pre_plane_stride(self) -> int:
        return int(self.info.pre_sharpen_plane_stride)

    def pack_input(self, x: np.ndarray, out: Optional[np.ndarray] = None) -> np.ndarray:
        """[3,H,W] array -> the reference's input buffer layout (plane stride (W+2)*H)."""
        assert x.shape == (3, self.h, self.w), x.shape
        buf = out if out is not None else np.zeros(3 * self.in_plane_stride, self.dtype)
        view = buf[: 3 * self.in_plane_stride].reshape(3, self.in_plane_stride)
        view[:, : self.w * self.h] = x.reshape(3, -1)
        return buf

    def radix_schedule(self):
        names = ["W", "H", "upH", "upW"]
        return {names[i]: [int(self.info.radices[i][s]) for s in range(self.info.n_stages[i])] for i in range(4)}

    # -- reference call sequence ---------------------------------------------------------------
    def upload(self, host_in):
        _check(self._lib.b2r_upload(self._h, _ptr(host_in)))

    def execute(self, num_iter: int = 1) -> float:
        ms = ctypes.c_double()
        _check(self._lib.b2r_execute(self._h, num_iter, ctypes.byref(ms)))
        return ms.value

    def download(self, host_out=None) -> np.ndarray:
        if host_out is None:
            host_out = np.empty((3, self.up_h, self.up_w), self.dtype)
        _check(self._lib.b2r_download(self._h, _ptr(host_out)))
        return host_out

    def upscale_host(self, host_in, host_out) -> float:
        ms = ctypes.c_double()
        _check(self._lib.b2r_upscale_host(self._h, _ptr(host_in), _ptr(host_out), ctypes.byref(ms)))
        return ms.value

    def upscale(self, x: np.ndarray) -> np.ndarray:
        """[3,H,W] in -> [3,upH,upW] out (upload + execute + download)."""
        self.upload(self.pack_input(np.ascontiguousarray(x, self.dtype)))
        self.execute(1)
        return self.download()

    # -- stage access for parity tests -----------------------------------------------------------
    def download_pre_sharpen(self) -> np.ndarray:
        buf = np.empty(3 * self.pre_plane_stride, self.dtype)
        _check(self._lib.b2r_download_pre_sharpen(self._h, buf.ctypes.data))
        n = self.up_w * self.up_h
        return np.stack([buf[c * self.pre_plane_stride: c * self.pre_plane_stride + n].reshape(self.up_h, self.up_w)
                         for c in range(3)])

    def sharpen_host(self, pre_planes: np.ndarray) -> np.ndarray:
        """run only the sharpen kernel on [3,upH,upW] planes (pad regions zero)"""
        buf = np.zeros(3 * self.pre_plane_stride, self.dtype)
        n = self.up_w * self.up_h
        for c in range(3):
            buf[c * self.pre_plane_stride: c * self.pre_plane_stride + n] = pre_planes[c].ravel()
        out = np.empty((3, self.up_h, self.up_w), self.dtype)
        _check(self._lib.b2r_sharpen_host(self._h, buf.ctypes.data, out.ctypes.data))
        return out

    # -- device-resident access --------------------------------------------------------------------
    @property
    def device_input(self) -> int:
        return int(self._lib.b2r_device_input(self._h))

    @property
    def device_output(self) -> int:
        return int(self._lib.b2r_device_output(self._h))

    @property
    def stream(self) -> int:
        return int(self._lib.b2r_plan_stream(self._h) or 0)

    @property
    def launch_count(self) -> int:
        return int(self._lib.b2r_plan_launch_count(self._h))

    def enqueue_device(self, device_in: int, device_out: int):
        """asynchronously process one device-resident frame (raw CUDA device pointers)"""
        _check(self._lib.b2r_enqueue_device(self._h, int(device_in), int(device_out)))

    def enqueue_host(self, host_in, host_out):
        """asynchronously upload + process + download one frame (pinned host buffers)"""
        _check(self._lib.b2r_enqueue_host(self._h, _ptr(host_in), _ptr(host_out)))

    def upscale_u8(self, rgb: np.ndarray) -> np.ndarray:
        """[H,W,3] uint8 in -> [upH,upW,3] uint8 out; both pixel conversions run on the GPU"""
        rgb = np.ascontiguousarray(rgb, np.uint8)
        assert rgb.shape == (self.h, self.w, 3), rgb.shape
        out = np.empty((self.up_h, self.up_w, 3), np.uint8)
        _check(self._lib.b2r_upload_u8(self._h, rgb.ctypes.data))
        self.execute(1)
        _check(self._lib.b2r_download_u8(self._h, out.ctypes.data))
        return out

    def enqueue_host_u8(self, host_in, host_out):
        _check(self._lib.b2r_enqueue_host_u8(self._h, _ptr(host_in), _ptr(host_out)))

    @property
    def last_ticket(self) -> int:
        return int(self._lib.b2r_plan_last_ticket(self._h))

    def wait_ticket(self, ticket: int):
        """block until the frame that took `ticket` (enqueue_host / enqueue_host_u8) has landed in its host buffer"""
        _check(self._lib.b2r_wait_ticket(self._h, ticket))

    def set_lanes(self, lanes: int):
        _check(self._lib.b2r_plan_set_lanes(self._h, lanes))

    @property
    def lanes(self) -> int:
        return int(self._lib.b2r_plan_lanes(self._h))

    def timer_start(self):
        _check(self._lib.b2r_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = ctypes.c_double()
        _check(self._lib.b2r_timer_stop(self._h, ctypes.byref(ms)))
        return ms.value

    def profile_kernels(self, num_iter: int = 10):
        """average device ms of (r2c_rows, cols, c2r_rows, sharpen)"""
        ms = (ctypes.c_double * 4)()
        _check(self._lib.b2r_profile_kernels(self._h, num_iter, ms))
        return dict(zip(("r2c_rows", "cols", "c2r_rows", "sharpen"), [float(v) for v in ms]))

    def synchronize(self):
        _check(self._lib.b2r_synchronize(self._h))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.b2r_plan_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
