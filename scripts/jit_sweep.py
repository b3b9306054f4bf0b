"""Sweep thread counts / tile widths / radix orders of the FFT kernels through the plan-time JIT
(B2R_FORCE_JIT=1 + B2R_TUNE_*), no rebuild needed (development aid).
usage: python scripts/jit_sweep.py [w h prec]   -- prints per-kernel us for every variant"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
args = sys.argv[1:4] if len(sys.argv) >= 3 else ["2048", "1024", "0"]
CODE = r'''
import sys, os
sys.path.insert(0, %r)
import numpy as np, vkresample_b200 as vb
w, h, prec = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 0
x = np.random.default_rng(0).random((3, h, w), dtype=np.float32)
with vb.Plan(w, h, 2.0, prec, 0.2) as p:
    p.upload(p.pack_input(x.astype(p.dtype))); p.execute(5)
    ms = min(p.execute(50) for _ in range(3))
    pk = p.profile_kernels(30)
    print(f"{ms*1e3:7.1f} us/frame jit={p.info.jit_kernels} cc={p.info.column_tile}", {k: round(v * 1e3, 1) for k, v in pk.items()}, p.radix_schedule(), p.info.jit_note.decode())
''' % ROOT

VARIANTS = {
    "c2": [
        {},
        {"B2R_TUNE_TW": "64", "B2R_TUNE_PPBW": "4"},
        {"B2R_TUNE_TW": "64", "B2R_TUNE_PPBW": "2"},
        {"B2R_TUNE_TW": "128", "B2R_TUNE_PPBW": "1"},
        {"B2R_TUNE_TW": "256", "B2R_TUNE_PPBW": "1"},
        {"B2R_TUNE_RW": "8,16,16"},
        {"B2R_TUNE_RW": "16,8,16"},
        {"B2R_TUNE_TH": "64", "B2R_TUNE_TUH": "64", "B2R_TUNE_CC": "4"},
        {"B2R_TUNE_TH": "64", "B2R_TUNE_TUH": "64", "B2R_TUNE_CC": "8"},
        {"B2R_TUNE_TH": "256", "B2R_TUNE_TUH": "256", "B2R_TUNE_CC": "2"},
        {"B2R_TUNE_RH": "4,16,16"},
        {"B2R_TUNE_RH": "16,4,16"},
        {"B2R_TUNE_RUH": "8,16,16"},
        {"B2R_TUNE_RUH": "16,8,16"},
        {"B2R_TUNE_RH": "4,16,16", "B2R_TUNE_RUH": "8,16,16"},
        {"B2R_TUNE_TUW": "128"},
        {"B2R_TUNE_TUW": "512"},
    ],
}
key = "c2"
variants = VARIANTS[key]
if len(sys.argv) > 4:   # extra variants from the command line: "A=1,B=2" ...
    variants = [{}] + [dict(kv.split("=") for kv in v.split(";")) for v in sys.argv[4:]]
for v in variants:
    env = dict(os.environ, B2R_FORCE_JIT="1", B2R_CACHE_DIR="/tmp/b2r_sweep_cache", **v)
    r = subprocess.run([sys.executable, "-c", CODE] + args, env=env, capture_output=True, text=True)
    print(v or "baseline (static table's schedule through the JIT)")
    print("   ", (r.stdout.strip() or r.stderr.strip()[-300:]))
