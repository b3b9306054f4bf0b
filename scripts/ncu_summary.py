"""Summarise an ncu report (--set full) into profiles/: per-kernel key metrics as markdown + the
dram traffic per launch as JSON (read by bench.py for roofline.traffic).
usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_ncu_c2 c2
"""
import csv, io, json, os, subprocess, sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__occupancy_limit_registers", "CTA/SM limit (regs)"),
    ("launch__occupancy_limit_shared_mem", "CTA/SM limit (smem)"),
    ("launch__waves_per_multiprocessor", "waves/SM"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
]
SHORT = {"k_r2c_rows": "r2c_rows", "k_cols": "cols", "k_c2r_sharpen": "c2r_sharpen", "k_c2r_rows": "c2r_rows",
         "k_sharpen_fix": "sharpen_fix", "k_sharpen": "sharpen"}   # first match wins (dict order)


def main():
    rep, out, cfg = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "c2")
    if rep.endswith(".csv.gz"):      # the raw page exported on the GPU box (scripts/collect_ncu.sh)
        import gzip
        txt = gzip.open(rep, "rt").read()
    else:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    seen, md, traffic = {}, [], {}
    for r in rows[2:]:
        name = r[ki]
        short = next((v for k, v in SHORT.items() if k in name), name[:30])
        if short in seen:
            continue
        seen[short] = True
        md.append(f"### {short}\n\n`{name[:150]}`\n\n| metric | value |\n|---|---|")
        vals = {}
        for m, label in WANT:
            if m in hdr:
                i = hdr.index(m)
                md.append(f"| {label} (`{m}`) | {r[i]} {units[i]} |")
                vals[m] = (r[i], units[i])
        md.append("")
        try:
            def to_bytes(v, u):
                f = float(v.replace(",", ""))
                return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            traffic[short] = int(to_bytes(*vals["dram__bytes_read.sum"]) + to_bytes(*vals["dram__bytes_write.sum"]))
        except Exception:
            pass
    with open(out + ".md", "w") as f:
        f.write(f"# ncu --set full summary ({os.path.basename(rep)}, config {cfg})\n\n"
                "Captured under `gpurun` with `--clock-control none`; cold-cache, serialised replays -- use the\n"
                "shares and per-launch byte counts, not the absolute times.\n\n" + "\n".join(md))
    tj_path = os.path.join(os.path.dirname(out), "ncu_traffic.json")
    tj = json.load(open(tj_path)) if os.path.exists(tj_path) else {}
    tj[cfg] = traffic
    json.dump(tj, open(tj_path, "w"), indent=1)
    print(open(out + ".md").read()[:600]); print(traffic)


if __name__ == "__main__":
    main()
