"""CPU oracle for the VkResample hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

This module restates, in numpy, the arithmetic the reference (DTolm/VkResample,
``/root/reference``) performs per frame.  It exists only so that ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs can *check* (or time, as the CPU baseline) the CUDA path.  The
product (``vkresample_b200`` / ``libb2resample.so``) never imports it and has no
CPU fallback.

Parity pinning
--------------
* PINNED by the reference's own golden image pairs
  ``samples/no_upscaling.png -> samples/FFT_upscaled.png`` and
  ``samples/no_upscaling_2.png -> samples/FFT_upscaled_2.png`` (README.md:55,
  ``-u 2``, fp32, sharpen 0.2): ``oracle/pin_goldens.py`` re-runs this oracle on
  the golden inputs and records max |diff| = 1 LSB, >= 99.993 % of bytes identical,
  in ``tests/golden/pin_record.json``.
* fp16 mode (``-p 2``), non-2x factors and white-noise behaviour are pinned by
  restatement only (the goldens do not cover them; the reference cannot be
  executed here: no Vulkan loader / lavapipe in the image).

Every function cites the reference region it follows (paths relative to
``/root/reference``).  Conventions: numpy FFT sign convention (forward e^{-i..});
the reference uses e^{+} forward / e^{-} inverse (vkFFT.h:4544-4545) which, for
real input, yields the conjugate spectrum and the same real output -- except for
the one place where the reference's processing is not conjugate-symmetric (the
complex DC bin of the C2R pack, see ``inverse_plane``), which is translated
explicitly.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

try:  # pocketfft with threads; numpy.fft fallback keeps the oracle importable anywhere
    import scipy.fft as _fft
    _HAVE_SCIPY = True
except Exception:  # pragma: no cover
    import numpy.fft as _fft
    _HAVE_SCIPY = False

CHANNELS = 3  # coordinateFeatures = 3, VkResample.cpp:1425


def _kw(workers):
    return {"workers": workers} if (_HAVE_SCIPY and workers) else {}


@dataclass(frozen=True)
class FramePlan:
    """Geometry of one plan -- VkResample.cpp:1409-1503 (configuration fill)."""
    w: int
    h: int
    upscale: float
    up_w: int
    up_h: int
    # inverse zero-pad ranges (R2C branch), VkResample.cpp:1490-1496
    zp_left_x: int   # = W/2      : stored columns >= this are never read by the inverse
    zp_right_x: int  # = upW/2
    zp_left_y: int   # = upH/(2 up)   : first zero row
    zp_right_y: int  # = (2up-1) upH/(2up) : first non-zero row of the negative block
    up2: float       # appSharpen.upscale = up*up as float, VkResample.cpp:1615

    @property
    def in_plane_stride(self) -> int:
        """Elements between channel planes of the input buffer: (W+2)*H, VkResample.cpp:1644."""
        return (self.w + 2) * self.h

    @property
    def pre_plane_stride(self) -> int:
        """Plane stride of the C2R output / sharpen input: (upW+2)*upH, VkResample.cpp:1596."""
        return (self.up_w + 2) * self.up_h

    @property
    def out_plane_stride(self) -> int:
        """Compact plane stride of the sharpen output: upW*upH, VkResample.cpp:1599-1601."""
        return self.up_w * self.up_h


def make_plan(w: int, h: int, upscale: float = 2.0) -> FramePlan:
    """Sizes exactly as the reference derives them (float math then uint32 truncation).

    ``bufferStride[i] = config.upscale * size[i]`` with ``float upscale``
    (VkResample.cpp:1417-1418), zero-pad bounds VkResample.cpp:1491-1496.
    """
    up = np.float32(upscale)
    up_w = int(np.float32(up * np.float32(w)))
    up_h = int(np.float32(up * np.float32(h)))
    # uint32 / float -> float, truncated on store to uint32
    zl_y = int(np.float32(up_h) / np.float32(np.float32(2.0) * up))
    zr_y = int(np.float32(np.float32(np.float32(2.0) * up - np.float32(1.0)) * np.float32(up_h))
               / np.float32(np.float32(2.0) * up))
    return FramePlan(w=w, h=h, upscale=float(up), up_w=up_w, up_h=up_h,
                     zp_left_x=w // 2, zp_right_x=up_w // 2,
                     zp_left_y=zl_y, zp_right_y=zr_y,
                     up2=float(np.float32(up * up)))


# --------------------------------------------------------------------------- a2
def fill_input(u8_hwc: np.ndarray, precision: int = 0) -> np.ndarray:
    """u8 interleaved HWC -> planar [3,H,W] in [0,1].  VkResample.cpp:1636-1685.

    fp32: ``(float)u8 / 255.0`` -- the division is carried out in double and
    narrowed to float on store (:1644).  fp16: ``(half)u8 / 255.0`` narrowed to
    half, round-to-nearest (half.hpp:373-374).  The returned array is *compact*
    ([3,H,W]); the reference's buffer has 2*H unused pad elements per plane.
    """
    x = u8_hwc[..., :CHANNELS].astype(np.float64) / 255.0
    x = np.ascontiguousarray(np.moveaxis(x, -1, 0))
    if precision == 2:
        return x.astype(np.float16)
    if precision == 1:
        return x                      # (double)u8 / 255.0, VkResample.cpp:1659
    return x.astype(np.float32)


# ---------------------------------------------------------------------- a3 + a4
def forward_spectrum(x: np.ndarray, dtype=np.float64, workers=None) -> np.ndarray:
    """Forward 2-D R2C, unnormalised.  vkFFT.h:7639-7680 (axis 0, R2C rows,
    two-rows-per-complex trick + split :4274-4376) and :7740-7789 (axis 1 and the
    DC "support" column).  Returns F[c, ky, kx] for kx = 0..W/2 (all W/2+1 bins are
    kept by the reference: DC in the extra column, Nyquist inside the main block)."""
    x = np.asarray(x, dtype=dtype)
    f = _fft.rfft(x, axis=-1, **_kw(workers))
    return _fft.fft(f, axis=-2, **_kw(workers))


# --------------------------------------------------------------------------- a5
def shift_zero_pad(f: np.ndarray, plan: FramePlan) -> np.ndarray:
    """fftshift-style relocation + implicit zero padding.

    Shift (R2C branch, VkResample.cpp:514-526): rows ky in [H/2, H) move to
    [upH-H/2, upH); rows [0, H/2) stay.  The inverse treats rows
    [zp_left_y, zp_right_y) and stored columns [W/2, upW/2) (bins kx > W/2) as zero
    (vkFFT.h:1277-1332, :1656-1717, :2077-2094; bounds VkResample.cpp:1491-1496).
    """
    c, h, nx = f.shape
    assert h == plan.h and nx == plan.w // 2 + 1
    b = np.zeros((c, plan.up_h, plan.up_w // 2 + 1), dtype=f.dtype)
    half = plan.h // 2
    b[:, :half, :nx] = f[:, :half]
    b[:, plan.up_h - (plan.h - half):, :nx] = f[:, half:]
    # rows the inverse reads as zero regardless of content (only differs from the
    # copy above for non-integer factors where the float bounds truncate)
    b[:, plan.zp_left_y:plan.zp_right_y, :] = 0
    return b


# ---------------------------------------------------------------------- a6 + a7
def inverse_plane(b: np.ndarray, plan: FramePlan, workers=None) -> np.ndarray:
    """Inverse along y (upH, 1/upH) then C2R along x (upW, 1/upW).

    vkFFT.h:8187-8242 (support + axis 1) and :8246-8288 (axis 0 C2R); per-stage
    1/radix normalisation :2917-2965 == numpy's 1/n.  The result is interp/up^2 -- the
    x up^2 lives in the sharpen shader.

    C2R DC quirk (vkFFT.h:2108-2131, the zero-pad branch VkResample takes; twin :2167-2190): the C2R packs spectrum
    rows 2j (A) and 2j+1 (B) into one complex sequence, Z[k] = A[k] + i B[k],
    Z[N-k] = conj A[k] + i conj B[k], and for the DC bin uses the *full complex* values
    ``sdata[0] = (A0.x - B0.y, A0.y + B0.x)``.  A true C2R would drop Im(A0), Im(B0); the
    reference instead leaks them into the partner row.  After the asymmetric placement
    of the y-Nyquist row (shift_zero_pad) the DC column is no longer Hermitian in ky, so
    Im(G[y][0]) != 0.  Translated from the reference's e^{+}-forward convention into
    numpy's:   row 2j   = irfft(A) + Im(B0)/upW,    row 2j+1 = irfft(B) - Im(A0)/upW.
    Pinned by the goldens: with this term 99.993 % / 99.995 % of the golden bytes match
    exactly (98.2 % / 97.6 % without it, 96.4 % / 95.3 % with the opposite sign).
    """
    g = _fft.ifft(b, axis=-2, **_kw(workers))
    o = _fft.irfft(g, n=plan.up_w, axis=-1, **_kw(workers))
    dc_im = g[..., 0].imag / plan.up_w            # [c, upH]
    o[:, 0::2, :] += dc_im[:, 1::2, None]
    o[:, 1::2, :] -= dc_im[:, 0::2, None]
    return o


def store_pre_sharpen(o: np.ndarray, precision: int) -> np.ndarray:
    """C2R store type: float, or float16_t round-to-nearest in half-memory mode
    (vkFFT.h:3525-3527, :7280-7293)."""
    return o.astype({0: np.float32, 1: np.float64, 2: np.float16}[precision])


def pre_sharpen(x: np.ndarray, plan: FramePlan, precision: int = 0, dtype=np.float64,
                workers=None) -> np.ndarray:
    """a3..a7: planar input -> stored C2R plane [3, upH, upW] (= interp / up^2)."""
    f = forward_spectrum(np.asarray(x, dtype=np.float32 if (dtype == np.float32 and precision != 1) else np.float64),
                         dtype=dtype, workers=workers)
    o = inverse_plane(shift_zero_pad(f, plan), plan, workers=workers)
    return store_pre_sharpen(o, precision)


# --------------------------------------------------------------------------- a8
def _flat_with_pad(pre: np.ndarray, plan: FramePlan) -> np.ndarray:
    """Lay the planes out as the reference's tempBuffer: plane stride (upW+2)*upH with
    a zero pad region (fp32 mode: nothing ever writes there; VkResample.cpp:1596) and
    zero slack after the last plane (reads past the buffer end are out of bounds in the
    reference; defined as 0 here and in the CUDA path)."""
    c = pre.shape[0]
    ps = plan.pre_plane_stride
    flat = np.zeros(c * ps + plan.up_w + 2, dtype=pre.dtype)
    for ch in range(c):
        flat[ch * ps: ch * ps + plan.up_w * plan.up_h] = pre[ch].ravel()
    return flat


def _literal(v) -> np.float32:
    """The value of a float pasted into the GLSL source with "%f" (6 decimals) and parsed back as
    a float literal (VkResample.cpp:893-920: ``app->upscale``, ``app->sharpenCoeff``)."""
    return np.float32(float("%f" % float(np.float32(v))))


def _sharpen_rows(t_all, base, up_w, up_h, y0, y1, s, dt, out_ch):
    """rows [y0, y1) of one channel; t_all = clamped magnitudes of the whole flat buffer"""
    ya = max(y0 - 1, 0)
    n_rows = (y1 + 1) - ya                                   # rows ya .. y1 inclusive
    e0 = t_all[base + ya * up_w: base + (ya + n_rows) * up_w].reshape(n_rows, up_w)           # X = x
    ep = t_all[base + ya * up_w + 1: base + (ya + n_rows) * up_w + 1].reshape(n_rows, up_w)   # X = x+1 (flat)
    em = np.empty_like(e0)
    em[:, 1:] = e0[:, :-1]
    em[:, 0] = e0[:, 0]                                                                         # X = max(x-1,0)
    ys = np.arange(y0, y1)
    r_m = np.maximum(ys - 1, 0) - ya
    r_0 = ys - ya
    r_p = ys + 1 - ya
    l0, l1, l2 = em[r_m], e0[r_m], ep[r_m]
    l3, l4, l5 = em[r_0], e0[r_0], ep[r_0]
    l6, l7, l8 = em[r_p], e0[r_p], ep[r_p]
    mn0 = np.minimum(l1, np.minimum(l3, np.minimum(l4, np.minimum(l5, l7))))
    mn1 = np.minimum(mn0, np.minimum(l0, np.minimum(l2, np.minimum(l6, l8))))
    mx0 = np.maximum(l1, np.maximum(l3, np.maximum(l4, np.maximum(l5, l7))))
    mx1 = np.maximum(mx0, np.maximum(l0, np.maximum(l2, np.maximum(l6, l8))))
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        minlen = dt(0.5) * (mn0 + mn1)
        maxlen = dt(0.5) * (mx0 + mx1)
        minlen = minlen / (dt(1.0) - minlen)
        maxlen = (dt(1.0) - maxlen) / maxlen
        scale = np.where(minlen < maxlen, minlen, maxlen)
        scale = -s * np.sqrt(scale)
        out_ch[y0:y1] = (l4 + scale * (((l1 + l3) + l5) + l7)) / (dt(1.0) + scale * dt(4.0))


def sharpen(pre: np.ndarray, plan: FramePlan, sharpen_const: float = 0.2,
            precision: int = 0, dtype=None, workers=None) -> np.ndarray:
    """FidelityFX-CAS-like 3x3 sharpen, R2C branch.  VkResample.cpp:849-923, strides
    :1564-1617, launch :1202-1219.

    Neighbour rule (:888-892): left/up clamp at 0; right/down are NOT clamped
    (``x < size`` is always true) and the index is flat ``X + Y*upW`` inside a plane of
    stride (upW+2)*upH -- so the right neighbour of the last column is the first pixel
    of the next row and the row below the last row is the (zero) pad region.

    ``dtype``: arithmetic type.  Default: float16 for precision 2 (the shader is
    generated with float16_t and HF literals, :823-827), else float32.  Pass
    np.float64 for the high-precision oracle.  ``workers`` > 1 splits the rows over
    threads (identical results; used by the timed CPU baseline).
    """
    if dtype is None:
        dtype = {0: np.float32, 1: np.float64, 2: np.float16}[precision]
    dt = np.dtype(dtype).type
    up_w, up_h, ps = plan.up_w, plan.up_h, plan.pre_plane_stride
    flat = _flat_with_pad(pre, plan).astype(dtype)
    # the "%f" texts are float literals in the fp32 / fp16 shaders and double literals in the -p 1 shader
    lit = (lambda v: float("%f" % float(np.float32(v)))) if precision == 1 else _literal
    # tex = up2 * in ; len = |tex| clamped to [0,1]     (:893-907)
    t_all = np.abs(dt(lit(plan.up2)) * flat)
    t_all = np.minimum(t_all, dt(1.0))
    t_all = np.maximum(t_all, dt(0.0))
    s = dt(lit(sharpen_const))
    out = np.empty((pre.shape[0], up_h, up_w), dtype=dtype)
    nw = int(workers) if workers else 1
    jobs = []
    for ch in range(pre.shape[0]):
        if nw <= 1:
            jobs.append((ch, 0, up_h))
        else:
            step = max(16, -(-up_h // nw))
            jobs += [(ch, y0, min(y0 + step, up_h)) for y0 in range(0, up_h, step)]
    if nw <= 1:
        for ch, y0, y1 in jobs:
            _sharpen_rows(t_all, ch * ps, up_w, up_h, y0, y1, s, dt, out[ch])
    else:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(nw) as ex:
            list(ex.map(lambda j: _sharpen_rows(t_all, j[0] * ps, up_w, up_h, j[1], j[2], s, dt, out[j[0]]), jobs))
    return out


# --------------------------------------------------------------------------- a9
def quantise(out: np.ndarray) -> np.ndarray:
    """planar [3,upH,upW] -> u8 HWC with the reference's truncating cast
    ``(uchar)(255.0 * v)`` (VkResample.cpp:1708-1748): multiply in double, truncate
    toward zero, wrap modulo 256 (x86 behaviour of the out-of-range conversion)."""
    v = 255.0 * out.astype(np.float64)
    v = np.nan_to_num(v, nan=0.0, posinf=0.0, neginf=0.0)
    q = np.trunc(v).astype(np.int64) & 0xFF
    return np.ascontiguousarray(np.moveaxis(q.astype(np.uint8), 0, -1))


# ------------------------------------------------------------------ whole frame
def upscale_frame(x: np.ndarray, upscale: float = 2.0, sharpen_const: float = 0.2,
                  precision: int = 0, dtype=np.float64, workers=None,
                  return_pre: bool = False):
    """performVulkanUpscale restated (VkResample.cpp:1249-1279): fwd FFT -> shift ->
    inverse FFT (zero padded) -> sharpen, on a planar [3,H,W] frame.

    dtype=float64: authoritative oracle (FFT and sharpen in double; in fp16 mode the
    input/C2R-store roundings to half are still applied and the sharpen is evaluated in
    double on the half-rounded plane).  dtype=float32: mimics the device arithmetic
    (complex64 FFTs; sharpen in float32, or in float16 when precision == 2).
    """
    c, h, w = x.shape
    plan = make_plan(w, h, upscale)
    pre = pre_sharpen(x, plan, precision=precision, dtype=dtype, workers=workers)
    if dtype == np.float64:
        sh_dtype = np.float64
    else:
        sh_dtype = None
    out = sharpen(pre, plan, sharpen_const, precision, dtype=sh_dtype, workers=workers)
    if precision == 2 and dtype != np.float64:
        out = out.astype(np.float16)
    elif dtype != np.float64 and precision != 1:
        out = out.astype(np.float32)
    if return_pre:
        return out, pre
    return out


def upscale_u8(u8_hwc: np.ndarray, upscale: float = 2.0, sharpen_const: float = 0.2,
               precision: int = 0, dtype=np.float64, workers=None) -> np.ndarray:
    """launchResample's per-frame body, PNG pixels in -> PNG pixels out
    (VkResample.cpp:1636-1748)."""
    x = fill_input(u8_hwc, precision)
    out = upscale_frame(x, upscale, sharpen_const, precision, dtype=dtype, workers=workers)
    return quantise(out)


# ------------------------------------------------------ the reference's other path
def reference_uses_r2c(up_w: int, max_shared_bytes: int = 49152, intel: bool = False) -> bool:
    """performR2C predicate (VkResample.cpp:1423-1424): the reference silently switches to its C2C
    path when upW > maxComputeSharedMemorySize / 8 (/4 more on Intel).  48 KB (NVIDIA Vulkan) ->
    threshold 6144: BASELINE config 5 (upW = 7680) would run C2C there.  The CUDA library keeps R2C
    semantics at every size; this predicate and ``upscale_frame_c2c`` exist so that the difference can
    be stated (tests/test_oracle.py::test_c2c_path_differs)."""
    return not (up_w > max_shared_bytes // 8 // (4 if intel else 1))


def upscale_frame_c2c(x: np.ndarray, upscale: float = 2.0, sharpen_const: float = 0.2, workers=None) -> np.ndarray:
    """The reference's C2C branch restated (float64): complex forward FFT of the real frame, 3-quadrant
    shift (VkResample.cpp:527-546: both Nyquist lines go to the negative side only), complex inverse
    with zero ranges [W/2, upW-W/2) x [H/2, upH-H/2) (:1498-1501), CAS on length(vec2) (:884-904) with
    plane stride upW*upH -- no pad rows: the row below the last row is the next channel's first row
    (:1598).  NOT pinned by any golden vector (the goldens were produced on the R2C path)."""
    c, h, w = x.shape
    plan = make_plan(w, h, upscale)
    up_w, up_h = plan.up_w, plan.up_h
    f = _fft.fft2(np.asarray(x, np.float64), axes=(-2, -1), **_kw(workers))
    b = np.zeros((c, up_h, up_w), np.complex128)
    hy, hx = h // 2, w // 2
    b[:, :hy, :hx] = f[:, :hy, :hx]
    b[:, :hy, up_w - (w - hx):] = f[:, :hy, hx:]
    b[:, up_h - (h - hy):, :hx] = f[:, hy:, :hx]
    b[:, up_h - (h - hy):, up_w - (w - hx):] = f[:, hy:, hx:]
    z = _fft.ifft2(b, axes=(-2, -1), **_kw(workers))
    t = np.minimum(np.abs(np.float64(_literal(plan.up2)) * z), 1.0)
    n = up_w * up_h
    flat = np.concatenate([t.reshape(-1), np.zeros(up_w + 2)])   # beyond the last channel: out of bounds -> 0
    s = np.float64(_literal(sharpen_const))
    out = np.empty((c, up_h, up_w))
    for ch in range(c):
        _sharpen_rows(flat, ch * n, up_w, up_h, 0, up_h, s, np.float64, out[ch])
    return out


# ------------------------------------------------------------ synthetic frames
def synthetic_frame(kind: str, w: int, h: int, seed: int = 1234) -> np.ndarray:
    """Deterministic synthetic inputs: 'noise' uniform [0,1), 'smooth'
    0.5+0.4 sin(x/37) cos(y/53+c), 'u8' PNG-like k/255 values.  float32 [3,H,W]."""
    if kind == "noise":
        return np.random.default_rng(seed).random((CHANNELS, h, w), dtype=np.float32)
    if kind == "u8":
        q = np.random.default_rng(seed).integers(0, 256, (CHANNELS, h, w))
        return (q.astype(np.float64) / 255.0).astype(np.float32)
    if kind == "smooth":
        xx = np.arange(w, dtype=np.float64)[None, None, :]
        yy = np.arange(h, dtype=np.float64)[None, :, None]
        cc = np.arange(CHANNELS, dtype=np.float64)[:, None, None]
        return (0.5 + 0.4 * np.sin(xx / 37.0) * np.cos(yy / 53.0 + cc)).astype(np.float32)
    raise ValueError(kind)
