#!/bin/bash
# Run on the GPU box (gpurun): one `ncu --set full` capture of ONE frame (its 4 kernel launches) per BASELINE
# config + the launch list of the bench command.  Reports land in gpurun_out/; summarise them here with
# scripts/ncu_summary.py into profiles/.   usage: scripts/collect_ncu.sh <tag>   (e.g. r2)
tag=${1:-r2}
mkdir -p gpurun_out
run() {  # name w h prec [env...]
  name=$1; w=$2; h=$3; p=$4; shift 4
  env "$@" ncu --set full --clock-control none -s 20 -c 4 -o /tmp/prof_${tag}_${name} -f \
      python scripts/quick_time.py $w $h $p > gpurun_out/ncu_${tag}_${name}.log 2>&1
  # only the raw metric table travels back (gpurun merges at most 64 MiB; a --set full report is ~15 MB)
  ncu -i /tmp/prof_${tag}_${name}.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/prof_${tag}_${name}.raw.csv.gz
}
run c2 2048 1024 0
run c2_separate 2048 1024 0 B2R_FUSED=0
run c3 1920 1080 2
run c4 2048 1024 2
run c5 3840 2160 0
# launch list of the bench command (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches_c2.csv \
    python bench.py --steps 2 --warmup 3 --frames-per-step 8 --no-exact-leg --no-cpu-baseline --min-seconds-e2e 0.02 > gpurun_out/ncu_${tag}_bench.log 2>&1
