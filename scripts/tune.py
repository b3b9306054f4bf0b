"""Tuning runs on the GPU box: column-tile width and multi-plan concurrency (development aid)."""
import os, sys, time, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

def one(cc):
    env = dict(os.environ)
    if cc: env["B2R_COLS_CC"] = str(cc)
    code = ("import sys; sys.path.insert(0,'.'); import numpy as np, vkresample_b200 as vb\n"
            "p=vb.Plan(2048,1024); x=np.random.default_rng(0).random((3,1024,2048),dtype=np.float32)\n"
            "p.upload(p.pack_input(x)); p.execute(5); ms=min(p.execute(50) for _ in range(3)); pk=p.profile_kernels(20)\n"
            "print('cc',p.info.column_tile,'frame us',round(ms*1e3,1),{k:round(v*1e3,1) for k,v in pk.items()})\n")
    print(subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True).stdout.strip())

def concurrency(nplans, frames=240):
    import vkresample_b200 as vb
    rng = np.random.default_rng(0)
    plans = [vb.Plan(2048, 1024) for _ in range(nplans)]
    for p in plans:
        p.upload(p.pack_input(rng.random((3, 1024, 2048), dtype=np.float32)))
        p.execute(3)
    lib = vb.load_library()
    import ctypes
    def run(n):
        for i in range(n):
            p = plans[i % nplans]
            lib.b2r_enqueue_device(p._h, ctypes.c_void_p(p.device_input), ctypes.c_void_p(p.device_output))
        for p in plans: p.synchronize()
    run(nplans * 4)
    t0 = time.perf_counter(); run(frames); dt = time.perf_counter() - t0
    print(f"{nplans} concurrent plan(s): {frames/dt:.0f} frames/s ({dt/frames*1e6:.1f} us/frame)")
    for p in plans: p.close()

if __name__ == "__main__":
    for cc in (0, 2, 8):
        one(cc)
    for n in (1, 2, 3):
        concurrency(n)
