"""Randomised parity sweep on the GPU box (development aid): random 2^a 3^b 5^c 7^d frame sizes, factors and
precisions through the default path (plan-time JIT or ahead-of-time kernels) against the oracle, with the same
bars as tests/test_gpu_parity.py.  usage: python scripts/fuzz_parity.py [n_cases] [seed]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vkresample_b200 as vb
from oracle import vkresample_oracle as vo

def smooth_sizes(limit):
    out = []
    for a in range(0, 12):
        for b in range(0, 8):
            for c in range(0, 6):
                for d in range(0, 5):
                    n = 2 ** a * 3 ** b * 5 ** c * 7 ** d
                    if 4 <= n <= limit and n % 2 == 0:
                        out.append(n)
    return sorted(set(out))

def ok_size(n):
    m = n
    for p in (2, 3, 5, 7):
        while m % p == 0:
            m //= p
    return m == 1 and n % 2 == 0 and n >= 4

def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rng = np.random.default_rng(seed)
    sizes = smooth_sizes(1600)
    bad = 0
    done = 0
    t0 = time.time()
    while done < n_cases:
        w, h = int(rng.choice(sizes)), int(rng.choice(sizes))
        up = float(rng.choice([1.0, 1.5, 2.0, 2.0, 2.0, 2.5, 3.0, 4.0]))
        prec = int(rng.choice([0, 0, 2, 1]))
        s = float(rng.choice([0.2, 0.2, 0.0, 0.1, 0.24]))
        uw, uh = int(np.float32(up) * np.float32(w)), int(np.float32(up) * np.float32(h))
        if not (ok_size(uw) and ok_size(uh)) or uw * uh > 6_000_000 or w * h < 64:
            continue
        done += 1
        kind = str(rng.choice(["noise", "u8", "smooth"]))
        x = vo.synthetic_frame(kind, w, h, int(rng.integers(1 << 30)))
        dt = {0: np.float32, 1: np.float64, 2: np.float16}[prec]
        try:
            with vb.Plan(w, h, up, prec, s, flags=vb.FLAG_EXACT_SHARPEN) as p:   # the bit-level statement
                out = p.upscale(x.astype(dt)).copy()
                pre = p.download_pre_sharpen()
                info = (p.info.static_kernels, p.info.jit_kernels, p.info.column_tile, p.radix_schedule())
            with vb.Plan(w, h, up, prec, s) as p:                                # the default (tolerance-bound, maybe fused) path
                out_d = p.upscale(x.astype(dt)).copy()
                info = info + (int(p.info.sharpen_mode), int(p.info.fused_strips_per_plane))
            plan_o = vo.make_plan(w, h, up)
            pre_o = vo.pre_sharpen(x.astype(dt), plan_o, precision=prec, dtype=np.float64, workers=os.cpu_count())
            e_pre = float(np.abs(pre.astype(np.float64) - pre_o).max() * plan_o.up2)
            sh_o = vo.sharpen(pre, plan_o, s, prec)
            bits = {2: np.uint16, 4: np.uint32, 8: np.uint64}[out.dtype.itemsize]
            exact = bool(np.all((sh_o.view(bits) == out.view(bits)) | (np.isnan(sh_o) & np.isnan(out))))
            tol = {0: 1e-5, 1: 1e-12, 2: 2e-3}[prec]
            fin = np.isfinite(sh_o.astype(np.float64))
            e_def = float(np.abs(out_d.astype(np.float64) - sh_o.astype(np.float64))[fin].max()) if fin.any() else 0.0
            good = exact and e_pre <= tol and e_def <= {0: 1e-5, 1: 0.0, 2: 1e-2}[prec]
        except Exception as e:   # noqa
            good, e_pre, exact, info, e_def = False, float("nan"), False, repr(e), float("nan")
        bad += not good
        print(f"{'ok ' if good else 'BAD'} {w}x{h} x{up} p={prec} s={s} {kind}: pre {e_pre:.2e} sharpen-exact {exact} default-vs-oracle {e_def:.2e} {info}", flush=True)
    print(f"{done} cases, {bad} failures, {time.time() - t0:.0f} s")
    return 1 if bad else 0

if __name__ == "__main__":
    sys.exit(main())
