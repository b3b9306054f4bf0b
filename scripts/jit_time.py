"""dynamic vs plan-time-JIT kernels for a size without an ahead-of-time schedule (development aid)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vkresample_b200 as vb
w, h, prec = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 0
os.environ["B2R_CACHE_DIR"] = "/tmp/b2r_cache_test"
x = np.random.default_rng(0).random((3, h, w), dtype=np.float32)
for flags in (0, vb.FLAG_JIT):
    t0 = time.perf_counter()
    with vb.Plan(w, h, 2.0, prec, 0.2, flags=flags) as p:
        t_plan = time.perf_counter() - t0
        p.upload(p.pack_input(x.astype(p.dtype))); p.execute(5)
        ms = min(p.execute(50) for _ in range(3))
        print(f"{w}x{h} p={prec} flags={flags}: plan {t_plan:.2f} s, {ms*1e3:.1f} us/frame, static={p.info.static_kernels} jit={p.info.jit_kernels} note='{p.info.jit_note.decode()}'",
              {k: round(v * 1e3, 1) for k, v in p.profile_kernels(20).items()})
    if flags == 0:
        import shutil; shutil.rmtree("/tmp/b2r_cache_test", ignore_errors=True)
