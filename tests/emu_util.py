"""ctypes driver for tests/emu/libb2r_emu.so (CPU thread emulation of the CUDA kernels; test
infrastructure only -- see tests/emu/emu.cpp)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(HERE, "emu", "libb2r_emu.so")
        srcs = [os.path.join(HERE, "emu", "emu.cpp")] + [
            os.path.join(HERE, "..", "vkresample_b200", "csrc", f)
            for f in ("b2r_fft.cuh", "b2r_kernels.cuh", "b2r_cas.cuh", "b2r_fused.cuh", "b2r_plan.cpp", "b2r_plan.h", "b2r_common.cuh",
                      "b2r_static_sizes.h")]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            subprocess.check_call([os.path.join(HERE, "emu", "build.sh")])
        L = ctypes.CDLL(so)
        L.b2r_emu_fft.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        L.b2r_emu_schedule.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        L.b2r_emu_frame.argtypes = ([ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_float,
                                     ctypes.c_float, ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 7)
        L.b2r_emu_sharpen.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_float,
                                      ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p]
        L.b2r_emu_sharpen_fast.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_float,
                                           ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_void_p, ctypes.c_void_p]
        L.b2r_emu_set_fused.argtypes = [ctypes.c_int]
        L.b2r_emu_set_r2c_bulk.argtypes = [ctypes.c_int]
        L.b2r_emu_set_cols_staged.argtypes = [ctypes.c_int]
        L.b2r_emu_set_cols_2x.argtypes = [ctypes.c_int]
        L.b2r_emu_set_sharpen_fast.argtypes = [ctypes.c_int]
        L.b2r_emu_u8_to_planar.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        L.b2r_emu_planar_to_u8.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        _LIB = L
    return _LIB


def schedule(n):
    rad = (ctypes.c_int * 8)()
    th = ctypes.c_int()
    ns = lib().b2r_emu_schedule(n, rad, ctypes.byref(th))
    return (list(rad)[:ns], th.value) if ns > 0 else (None, 0)


def fft(x, direction, use_static=False):
    x = np.ascontiguousarray(x, np.complex64)
    out = np.zeros_like(x)
    rc = lib().b2r_emu_fft(x.size, direction, int(use_static), x.ctypes.data, out.ctypes.data)
    assert rc >= 0
    return out, rc


def pack_input(x, dtype):
    """[3,H,W] -> the reference's input buffer (plane stride (W+2)*H, VkResample.cpp:1644)."""
    c, h, w = x.shape
    buf = np.zeros(c * (w + 2) * h, dtype)
    for ch in range(c):
        buf[ch * (w + 2) * h: ch * (w + 2) * h + w * h] = x[ch].astype(dtype).ravel()
    return buf


def unpack_pre(pre, plan):
    ps = plan.pre_plane_stride
    return np.stack([pre[c * ps: c * ps + plan.up_w * plan.up_h].reshape(plan.up_h, plan.up_w) for c in range(3)])


def frame(x, up, precision, sharpen, plan, cc=4, use_static=True):
    dt = np.float16 if precision == 2 else np.float32
    c, h, w = x.shape
    buf = pack_input(x, dt)
    out = np.zeros((3, plan.up_h, plan.up_w), dt)
    ss = ((w // 2 + 1 + 15) // 16) * 16
    spec1 = np.zeros((3, h, ss), np.complex64)
    spec2 = np.zeros((3, plan.up_h, ss), np.complex64)
    pre = np.zeros(3 * plan.pre_plane_stride + plan.up_w + 8, dt)
    stride = ctypes.c_int()
    used = ctypes.c_int()
    rc = lib().b2r_emu_frame(w, h, up, precision, sharpen, plan.up2, cc, int(use_static), buf.ctypes.data,
                             out.ctypes.data, spec1.ctypes.data, spec2.ctypes.data, pre.ctypes.data,
                             ctypes.byref(stride), ctypes.byref(used))
    assert rc == 0, rc
    assert stride.value == ss
    return dict(out=out, spec1=spec1[:, :, :w // 2 + 1], spec2=spec2[:, :, :w // 2 + 1],
                pre=unpack_pre(pre, plan), pre_flat=pre, used_static=used.value)
