"""print the warp stall breakdown (cycles stalled per issued instruction) of every kernel in an ncu report"""
import csv, io, subprocess, sys
if sys.argv[1].endswith(".csv.gz"):
    import gzip
    txt = gzip.open(sys.argv[1], "rt").read()
else:
    txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[0]
ki = hdr.index("Kernel Name")
cols = [i for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
seen = set()
for r in rows[2:]:
    name = r[ki][:48]
    if name in seen:
        continue
    seen.add(name)
    vals = sorted([(float(r[i]), hdr[i][len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for i in cols if r[i]], reverse=True)[:6]
    print(name)
    print("    " + "  ".join(f"{n}={v:.2f}" for v, n in vals))
