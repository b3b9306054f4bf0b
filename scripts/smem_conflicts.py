#!/usr/bin/env python
"""Shared-memory wavefront model of the exact-2x column kernel (k_cols2x, b2r_kernels.cuh).

Counts, for one frame, the 64-bit shared-memory wavefronts the kernel's LDS/STS issue under a padding rule,
and how many of them are bank-conflict replays.  Model: a warp's 64-bit access is served per half-warp; a
half-warp needs as many wavefronts as the most loaded of the sixteen 8-byte bank pairs (distinct addresses
only).  Used to choose the padding rule per (schedule, columns-per-CTA) without GPU time; the totals are
compared with ncu's l1tex__data_pipe_lsu_wavefronts_mem_shared / l1tex__data_bank_conflicts_pipe_lsu_mem_shared
in profiles/README.md.

    python scripts/smem_conflicts.py                # the built-in BASELINE schedules, every rule
"""
import sys
from collections import Counter


def pad16(i):      # smem_pad of b2r_fft.cuh: one element of padding after every 16
    return i + (i >> 4)


def nopad(i):
    return i


def make_pad(shift):
    return lambda i: i + (i >> shift)


def halfwarp_wavefronts(addrs):
    """addrs: element (8-byte) addresses of up to 16 active lanes."""
    if not addrs:
        return 0
    c = Counter(a % 16 for a in set(addrs))
    return max(c.values())


def warp_access(lane_addrs):
    """lane_addrs: list of 32 entries (address or None) -> (wavefronts, ideal)"""
    w = 0
    ideal = 0
    for h in (0, 16):
        a = [x for x in lane_addrs[h:h + 16] if x is not None]
        w += halfwarp_wavefronts(a)
        ideal += 1 if a else 0
    return w, ideal


def stages_of(n, radices):
    out, s = [], 1
    for r in radices:
        out.append((r, n // r, s))
        s *= r
    return out


def cols2x_frame(n, t, radices, cc, pad, nx=1025, channels=3):
    """wavefronts of one frame; returns (total, ideal)"""
    stages = stages_of(n, radices)
    threads = t * cc
    total = ideal = 0

    def access(fn):
        # fn(tid, c) -> list of addresses (one per unrolled access slot) or None when the thread is idle
        nonlocal total, ideal
        per_thread = [fn(x // cc, x % cc) for x in range(threads)]
        nslots = max(len(p) for p in per_thread if p is not None)
        for w0 in range(0, threads, 32):
            for s in range(nslots):
                lanes = []
                for x in range(w0, min(w0 + 32, threads)):
                    p = per_thread[x]
                    lanes.append(p[s] if p is not None and s < len(p) else None)
                lanes += [None] * (32 - len(lanes))
                w, i = warp_access(lanes)
                total += w
                ideal += i

    def store(r, nb, s):
        nbt = -(-nb // t)

        def fn(tid, c):
            out = []
            for b in range(nbt):
                j = tid + b * t
                if j >= nb:
                    continue
                q, p = divmod(j, s)
                base = q * s * r + p
                out += [pad((base + k * s) * cc + c) for k in range(r)]
            return out or None
        access(fn)

    def load(r, nb, s):
        nbt = -(-nb // t)

        def fn(tid, c):
            out = []
            for b in range(nbt):
                j = tid + b * t
                if j >= nb:
                    continue
                out += [pad((j + i * nb) * cc + c) for i in range(r)]
            return out or None
        access(fn)

    # forward: stage 0 from global -> store; stages 1.. load + store
    # inverse: stage 0 loads (in place) -> store; middle stages load + store; last stage load only
    r0, nb0, s0 = stages[0]
    store(r0, nb0, s0)
    for (r, nb, s) in stages[1:]:
        load(r, nb, s)
        store(r, nb, s)
    load(r0, nb0, s0)
    if len(stages) > 1:
        store(r0, nb0, s0)
        for (r, nb, s) in stages[1:-1]:
            load(r, nb, s)
            store(r, nb, s)
        r, nb, s = stages[-1]
        load(r, nb, s)
    tiles = -(-nx // cc) * channels
    return total * tiles, ideal * tiles


CASES = {
    "c2 cols2x<1024: 64 x (16,16,4), CC 8>": (1024, 64, (16, 16, 4), 8, 1025),
    "c3/c4 cols2x<1080: 90 x (15,12,6), CC 8>": (1080, 90, (15, 12, 6), 8, 961),
    "c5 cols2x<2160: 144 x (16,9,15), CC 4>": (2160, 144, (16, 9, 15), 4, 1921),
    "c1 cols2x<512>: 32 x (16,8,4), CC 4": (512, 32, (16, 8, 4), 4, 257),
}

if __name__ == "__main__":
    rules = {"pad16 (i + i/16)": pad16, "no padding": nopad, "i + i/32": make_pad(5), "i + i/8": make_pad(3),
             "i + i/64": make_pad(6), "i + i/128": make_pad(7)}
    for name, (n, t, rad, cc, nx) in CASES.items():
        print(name)
        for rn, rule in rules.items():
            tot, ideal = cols2x_frame(n, t, rad, cc, rule, nx=nx)
            print(f"    {rn:18s} wavefronts {tot/1e6:7.3f} M   conflict replays {(tot-ideal)/1e6:7.3f} M ({(tot-ideal)/tot:5.1%})")
