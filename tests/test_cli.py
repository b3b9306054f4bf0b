"""The b2resample CLI: PNG codec self-test (CPU) and the VkResample command line end to end (GPU)."""
import os
import subprocess

import numpy as np
import pytest

import vkresample_b200 as vb
from oracle import vkresample_oracle as vo

CLI = os.path.join(os.path.dirname(vb.library_path()), "b2resample")


def _cli_or_skip():
    if not os.path.exists(CLI):
        pytest.skip("CLI not built (run `make`)")


@pytest.mark.parametrize("mode", ["RGB", "RGBA", "L", "P", "LA"])
def test_png_codec_roundtrip(tmp_path, mode):
    """decode (all colour types stbi_load(...,3) accepts) + encode must preserve the RGB pixels"""
    _cli_or_skip()
    from PIL import Image
    rng = np.random.default_rng(3)
    rgb = rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)
    rgb[5:20, 7:30] = (rgb[5:20, 7:30] // 32) * 32  # some flat-ish areas for the filters
    img = Image.fromarray(rgb, "RGB")
    if mode == "RGBA":
        img.putalpha(Image.fromarray(rng.integers(0, 256, (37, 53), dtype=np.uint8)))
    elif mode in ("L", "LA", "P"):
        img = img.convert(mode)
    src, dst = str(tmp_path / "a.png"), str(tmp_path / "b.png")
    img.save(src)
    expect = np.asarray(Image.open(src).convert("RGB") if mode != "RGBA" else np.asarray(Image.open(src))[..., :3])
    subprocess.check_call([CLI, "-pngcopy", src, "-o", dst])
    got = np.asarray(Image.open(dst))
    assert got.shape == expect.shape and np.array_equal(got, expect)


def test_cli_help_and_argument_errors():
    _cli_or_skip()
    out = subprocess.run([CLI, "-h"], capture_output=True, text=True)
    assert out.returncode == 0
    for flag in ("-i", "-o", "-ifolder", "-ofolder", "-u", "-p", "-s", "-n", "-numfiles", "-numthreads", "-d", "-devices"):
        assert flag in out.stdout
    out = subprocess.run([CLI, "-u", "2"], capture_output=True, text=True)
    assert out.returncode != 0 and "No input file is selected with -i flag" in out.stdout


def _lsb_circular(a, b):
    d = np.abs(a.astype(np.int16) - b.astype(np.int16))
    return np.minimum(d, 256 - d)


@pytest.mark.gpu
def test_cli_single_image_matches_oracle(tmp_path):
    """VkResample -i in.png -u 2 -o out.png (README.md:55) -> within 1 LSB of the oracle's PNG bytes"""
    _cli_or_skip()
    from PIL import Image
    rng = np.random.default_rng(11)
    yy, xx = np.mgrid[0:120, 0:160]
    base = 127 + 100 * np.sin(xx / 9.0)[..., None] * np.cos(yy[..., None] / 7.0 + np.arange(3))
    img = np.clip(base + rng.integers(-20, 20, (120, 160, 3)), 0, 255).astype(np.uint8)
    src, dst = str(tmp_path / "in.png"), str(tmp_path / "out.png")
    Image.fromarray(img, "RGB").save(src)
    r = subprocess.run([CLI, "-i", src, "-o", dst, "-u", "2", "-n", "3"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "VkResample 2.0x upscale: 160x120 to 320x240 Time:" in r.stdout and "Total time:" in r.stdout
    got = np.asarray(Image.open(dst))
    ref = vo.upscale_u8(img, 2.0, 0.2, 0, dtype=np.float64)
    d = _lsb_circular(got, ref)
    assert got.shape == (240, 320, 3) and d.max() <= 1 and (d == 0).mean() > 0.995


@pytest.mark.gpu
def test_cli_batch_folder_threads(tmp_path):
    """-ifolder/-ofolder/-numfiles/-numthreads: file f is handled by thread (f-1) % numthreads and
    the result does not depend on the thread count (VkResample.cpp:1622-1629)"""
    _cli_or_skip()
    from PIL import Image
    rng = np.random.default_rng(5)
    ind, out1, out3 = tmp_path / "in", tmp_path / "o1", tmp_path / "o3"
    for d in (ind, out1, out3):
        d.mkdir()
    frames = []
    for f in range(1, 6):
        img = rng.integers(0, 256, (64, 96, 3), dtype=np.uint8)
        frames.append(img)
        Image.fromarray(img, "RGB").save(str(ind / f"{f:06d}.png"))
    for nthreads, od in ((1, out1), (3, out3)):
        r = subprocess.run([CLI, "-ifolder", str(ind), "-ofolder", str(od), "-numfiles", "5", "-numthreads", str(nthreads),
                            "-u", "2", "-p", "2", "-s", "0.1"], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
    for f in range(1, 6):
        a = np.asarray(Image.open(str(out1 / f"{f:06d}.png")))
        b = np.asarray(Image.open(str(out3 / f"{f:06d}.png")))
        assert np.array_equal(a, b)
        ref = vo.quantise(vo.upscale_frame(vo.fill_input(frames[f - 1], 2), 2.0, 0.1, 2, dtype=np.float32))
        d = _lsb_circular(a, ref)
        assert d[:-1].max() <= 3   # fp16 storage: a few LSB; the last row depends on stale memory in the reference
