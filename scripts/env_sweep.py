"""Sweep environment switches on the GPU box (development aid): every argument is one variant,
"VAR=a,VAR2=b" (or "-" for the defaults); each runs scripts/quick_time.py in a fresh process.
  python scripts/env_sweep.py [--size W H PREC] - B2R_SHARPEN_RY=12 B2R_SHARPEN_RY=48,B2R_SHARPEN_REVERSE=0"""
import os, subprocess, sys
args = sys.argv[1:]
size = []
if args and args[0] == "--size":
    size, args = args[1:4], args[4:]
for variant in args:
    env = dict(os.environ)
    if variant != "-":
        for kv in variant.split(","):
            k, v = kv.split("=")
            env[k] = v
    r = subprocess.run([sys.executable, "scripts/quick_time.py"] + size, env=env, capture_output=True, text=True)
    print("##", variant, "\n" + r.stdout.strip() + ("\n" + r.stderr.strip()[-400:] if r.returncode else ""), flush=True)
