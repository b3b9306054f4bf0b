#!/usr/bin/env python
"""BASELINE config 4 as stated: a 256-frame 2048x1024 -> 4096x2048 fp16 stream, whole frames sharded over the
GPUs of one box (frame f, 1-based, -> rank (f-1) mod world: the reference's striding, VkResample.cpp:1622-1629;
vkresample_b200/sharding.py), no collective on the data path.

  python scripts/c4_stream.py --out profiles/r2_c4_stream_n1.json                                  # 1 GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \\
         --master-port 29511 scripts/c4_stream.py --out profiles/r2_c4_stream_n8.json               # 8 GPUs
  python scripts/c4_stream.py --compare profiles/r2_c4_stream_n1.json profiles/r2_c4_stream_n8.json

Pass 1 (timed, max over ranks): every rank pushes its frames through b2r_enqueue_host (pinned host in, pinned
host out, copies inside the timed region) and, separately, through b2r_enqueue_device (inputs resident).
Pass 2 (untimed): the same frames again, every output hashed (blake2b of the raw fp16 bytes); rank 0 gathers the
digests and writes them with the timings.  --compare asserts that two runs (different GPU counts) produced
byte-identical output for every frame.
"""
import argparse
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def frame_of(f, w, h):
    """frame f of the synthetic stream (deterministic, independent of who processes it)"""
    return np.random.default_rng(100000 + f).random((3, h, w), dtype=np.float32).astype(np.float16)


def compare(a, b):
    da, db = json.load(open(a)), json.load(open(b))
    assert da["frames"] == db["frames"] and da["config"].split(", frame f")[0] == db["config"].split(", frame f")[0], "different streams"
    bad = [f for f in da["digests"] if da["digests"][f] != db["digests"].get(f)]
    print(json.dumps({"compare": [a, b], "frames": da["frames"], "world": [da["world"], db["world"]],
                      "byte_identical_frames": da["frames"] - len(bad), "mismatching_frames": bad[:8]}))
    return 1 if bad else 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--size", type=int, nargs=2, default=[2048, 1024])
    ap.add_argument("--precision", type=int, default=2)
    ap.add_argument("--lanes", type=int, default=3)
    ap.add_argument("--out", default=None)
    ap.add_argument("--compare", nargs=2, default=None)
    args = ap.parse_args()
    if args.compare:
        return compare(*args.compare)

    import torch
    import vkresample_b200 as vb
    from vkresample_b200 import sharding
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    w, h = args.size
    mine = sharding.frames_for_worker(args.frames, world, rank)
    plan = vb.Plan(w, h, 2.0, args.precision, 0.2, device=local)
    plan.set_lanes(args.lanes)
    np_dt = np.float16 if args.precision == 2 else np.float32
    pool = 2 * args.lanes
    h_in = [torch.from_numpy(plan.pack_input(frame_of(f, w, h).astype(np_dt)).view(np.uint8)).pin_memory() for f in mine]
    h_out = [torch.empty(plan.output_bytes, dtype=torch.uint8).pin_memory() for _ in range(pool)]

    def allmax(v):
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up
    for i in range(min(pool, len(mine))):
        plan.enqueue_host(h_in[i].data_ptr(), h_out[i % pool].data_ptr())
    plan.synchronize()
    # pass 1a: host-fed stream (buffer i mod pool is always used by lane i mod lanes: stream order keeps reuse safe)
    barrier()
    t0 = time.perf_counter()
    for i in range(len(mine)):
        plan.enqueue_host(h_in[i].data_ptr(), h_out[i % pool].data_ptr())
    plan.synchronize()
    t_host = allmax(time.perf_counter() - t0)
    # pass 1b: device-resident stream
    d_in = [t.to(dev) for t in h_in]
    d_out = [torch.empty(plan.output_bytes, dtype=torch.uint8, device=dev) for _ in range(pool)]
    torch.cuda.synchronize()
    barrier()
    plan.timer_start()
    for i in range(len(mine)):
        plan.enqueue_device(d_in[i].data_ptr(), d_out[i % pool].data_ptr())
    t_dev = allmax(plan.timer_stop() * 1e-3)
    # pass 2: digests
    digests = {}
    for i0 in range(0, len(mine), pool):
        chunk = list(range(i0, min(i0 + pool, len(mine))))
        for i in chunk:
            plan.enqueue_host(h_in[i].data_ptr(), h_out[i % pool].data_ptr())
        plan.synchronize()
        for i in chunk:
            digests[str(mine[i])] = hashlib.blake2b(h_out[i % pool].numpy().tobytes(), digest_size=16).hexdigest()
    gathered = [digests]
    if dist is not None:
        gathered = [None] * world
        dist.all_gather_object(gathered, digests)
    if rank == 0:
        allf = {}
        for part in gathered:
            allf.update(part)
        assert sorted(int(k) for k in allf) == list(range(1, args.frames + 1)), "every frame exactly once"
        res = {"config": f"c4: {args.frames}-frame {w}x{h}->{plan.up_w}x{plan.up_h} {'fp16' if args.precision == 2 else 'fp32'} stream, "
                         f"frame f -> rank (f-1) mod {world}", "frames": args.frames, "world": world, "lanes": args.lanes,
               "frames_per_s_host_fed": args.frames / t_host, "frames_per_s_device_resident": args.frames / t_dev,
               "seconds_host_fed": t_host, "seconds_device_resident": t_dev,
               "h2d_bytes_per_frame": plan.input_bytes, "d2h_bytes_per_frame": plan.output_bytes,
               "digests": dict(sorted(allf.items(), key=lambda kv: int(kv[0])))}
        if args.out:
            os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
            json.dump(res, open(args.out, "w"), indent=0)
        print(json.dumps({k: v for k, v in res.items() if k != "digests"}))
    plan.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
