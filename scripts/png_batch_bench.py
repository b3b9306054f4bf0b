#!/usr/bin/env python
"""PNG-fed batch throughput of the b2resample CLI: a folder of 2048x1024 PNG frames through
`-ifolder/-ofolder` in the default pipelined engine and in the reference-style synchronous loop (-sync).
  python scripts/png_batch_bench.py --frames 64 --gpus 1 --threads 16 [--out profiles/r2_png_batch_1gpu.json]
Prints one JSON line.  The codec (zlib inflate / deflate + PNG filters on the host cores) is inside the clock --
that is the point of the figure; `frames/s` here is bounded by PNG encoding, not by the GPU."""
import argparse, json, os, re, shutil, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--size", type=int, nargs=2, default=[2048, 1024])
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 8, help="codec threads per GPU (pipeline) / worker threads (sync)")
    ap.add_argument("--pnglevel", type=int, default=1)
    ap.add_argument("--precision", type=int, default=0)
    ap.add_argument("--skip-sync", action="store_true")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    from PIL import Image
    import vkresample_b200 as vb
    cli = os.path.join(os.path.dirname(vb.library_path()), "b2resample")
    w, h = args.size
    tmp = tempfile.mkdtemp(prefix="b2r_png_")
    ind = os.path.join(tmp, "in"); os.makedirs(ind)
    yy, xx = np.mgrid[0:h, 0:w]
    rng = np.random.default_rng(1)
    for f in range(1, args.frames + 1):   # natural-ish content: smooth structure + mild noise (compresses like a photo)
        base = 127 + 90 * np.sin((xx + 3 * f) / 37.0)[..., None] * np.cos(yy[..., None] / 53.0 + np.arange(3))
        img = np.clip(base + rng.integers(-6, 7, (h, w, 3)), 0, 255).astype(np.uint8)
        Image.fromarray(img, "RGB").save(os.path.join(ind, f"{f:06d}.png"), compress_level=1)
    res = {"frames": args.frames, "size": [w, h], "gpus": args.gpus, "threads": args.threads, "pnglevel": args.pnglevel,
           "precision": args.precision, "host_cores": os.cpu_count()}
    modes = [("pipeline", [])] + ([] if args.skip_sync else [("sync", ["-sync"])])
    for mode, extra in modes:
        od = os.path.join(tmp, mode); os.makedirs(od)
        nthreads = args.threads if mode == "pipeline" else args.threads * args.gpus
        cmd = [cli, "-ifolder", ind, "-ofolder", od, "-numfiles", str(args.frames), "-u", "2", "-p", str(args.precision),
               "-numthreads", str(nthreads), "-gpus", str(args.gpus), "-pnglevel", str(args.pnglevel)] + extra
        t0 = time.perf_counter()
        r = subprocess.run(cmd, capture_output=True, text=True)
        dt = time.perf_counter() - t0
        if r.returncode != 0:
            res[mode] = {"error": (r.stdout + r.stderr)[-400:]}
            continue
        m = re.search(r"Total time: ([0-9.]+) s", r.stdout)
        total = float(m.group(1)) if m else dt
        res[mode] = {"frames_per_s": args.frames / total, "total_s": total, "wall_s": dt,
                     "stdout_tail": [l for l in r.stdout.splitlines() if "finished" in l or "Pipelined" in l or "timeline" in l][-17:]}
        n_out = len([f for f in os.listdir(od) if f.endswith(".png")])
        res[mode]["files_written"] = n_out
    shutil.rmtree(tmp, ignore_errors=True)
    if args.out:
        json.dump(res, open(args.out, "w"), indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
