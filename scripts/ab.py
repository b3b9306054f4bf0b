"""A/B timing of an environment switch on the GPU box (development aid): python scripts/ab.py VAR=1 [w h prec]"""
import os, subprocess, sys
var = sys.argv[1]
rest = sys.argv[2:]
for setting in ("", var):
    env = dict(os.environ)
    if setting:
        k, v = setting.split("=")
        env[k] = v
    print("##", setting or "default")
    print(subprocess.run([sys.executable, "scripts/quick_time.py"] + rest, env=env, capture_output=True, text=True).stdout.strip())
