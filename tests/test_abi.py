"""The C-ABI library loads and exports every symbol include/b2resample.h declares; argument
validation that needs no GPU.  (Compute calls are exercised by the -m gpu tests.)"""
import ctypes
import os
import re

import pytest

import vkresample_b200 as vb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib_or_skip():
    if not os.path.exists(vb.library_path()):
        pytest.skip("libb2resample.so not built (run `make` or __graft_entry__.build())")
    return vb.load_library()


def test_header_symbols_all_exported():
    lib = _lib_or_skip()
    hdr = open(os.path.join(ROOT, "include", "b2resample.h")).read()
    declared = set(re.findall(r"\b(b2r_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"b2r_plan_info"}
    assert declared == set(vb.EXPORTS), declared ^ set(vb.EXPORTS)
    for name in declared:
        assert getattr(lib, name) is not None


def test_no_torch_types_in_header():
    hdr = open(os.path.join(ROOT, "include", "b2resample.h")).read()
    assert "torch" not in hdr and "at::" not in hdr and 'extern "C"' in hdr


def test_invalid_arguments_fail_loudly():
    lib = _lib_or_skip()
    h = ctypes.c_void_p()
    assert lib.b2r_plan_create(ctypes.byref(h), 0, 256, 128, 2.0, 3, 0.2, 0) == -1   # precision must be 0 / 1 / 2
    assert b"precision" in lib.b2r_last_error()
    assert lib.b2r_plan_create(ctypes.byref(h), 0, 255, 128, 2.0, 0, 0.2, 0) == -1   # odd width
    assert lib.b2r_plan_create(ctypes.byref(h), 0, 256, 128, 0.5, 0, 0.2, 0) == -1   # factor < 1
    assert lib.b2r_plan_create(ctypes.byref(h), 0, 2 * 11, 128, 2.0, 0, 0.2, 0) == -3  # prime factor 11
    assert lib.b2r_plan_create(None, 0, 256, 128, 2.0, 0, 0.2, 0) == -1
    assert lib.b2r_execute(None, 1, None) == -1
    assert lib.b2r_upload(None, None) == -1
    assert lib.b2r_version().startswith(b"b2resample")


def test_no_cpu_fallback_without_device():
    """on a box without a GPU plan creation must fail with B2R_ERR_CUDA, never compute on the CPU"""
    lib = _lib_or_skip()
    if lib.b2r_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(vb.B2RError) as e:
        vb.Plan(256, 128)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "vkresample_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "oracle/" not in txt, f


def test_cached_cubin_images_are_checked_before_loading(tmp_path):
    """cuModuleLoadData takes no length, so a truncated file in the JIT cache must never reach it: the check the
    cache applies (ELF64 header, section / program header tables and section contents inside the image) accepts a
    real sm_100a cubin and rejects every truncation of it and garbage"""
    import ctypes
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("no nvcc to make a cubin with")
    src = tmp_path / "k.cu"
    src.write_text("__global__ void k(float* p) { p[threadIdx.x] *= 2.f; }\n")
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-cubin", "-o", str(tmp_path / "k.cubin"), str(src)],
                   check=True, timeout=300)
    img = (tmp_path / "k.cubin").read_bytes()
    lib = vb.load_library()
    ok = lib.b2r_debug_cubin_image_ok
    ok.argtypes = [ctypes.c_char_p, ctypes.c_size_t]
    ok.restype = ctypes.c_int
    assert ok(img, len(img)) == 1
    for n in (0, 1, 63, 64, 1000, len(img) // 2, len(img) - 1):
        assert ok(img[:n], n) == 0, n
    assert ok(b"x" * 5000, 5000) == 0
    assert ok(b"\x7fELF" + b"\xff" * 4096, 4100) == 0
