"""Quick device timing of one config: per-kernel ms and frames/s (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vkresample_b200 as vb

def main():
    w, h, prec = 2048, 1024, 0
    if len(sys.argv) > 2: w, h = int(sys.argv[1]), int(sys.argv[2])
    if len(sys.argv) > 3: prec = int(sys.argv[3])
    rng = np.random.default_rng(0)
    x = rng.random((3, h, w), dtype=np.float32)
    with vb.Plan(w, h, 2.0, prec, 0.2) as p:
        p.upload(p.pack_input(x.astype(p.dtype)))
        p.execute(5)
        ms = min(p.execute(50) for _ in range(3))
        pk = p.profile_kernels(20)
        print(f"{w}x{h} p={prec}: {ms*1e3:.1f} us/frame = {1e3/ms:.0f} frames/s; static={p.info.static_kernels} "
              f"cc={p.info.column_tile} fused_nsp={p.info.fused_strips_per_plane} sched={p.radix_schedule()}")
        print("  per-kernel us:", {k: round(v * 1e3, 1) for k, v in pk.items()}, "sum", round(sum(pk.values()) * 1e3, 1))

if __name__ == "__main__":
    main()
