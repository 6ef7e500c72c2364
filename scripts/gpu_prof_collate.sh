#!/bin/bash
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:n1_gather -s 2 -c 1 -o gpurun_out/prof_n1_gather -f python scripts/profile_collate.py > gpurun_out/prof_n1.log 2>&1
echo "ncu exit $?"
ncu -i gpurun_out/prof_n1_gather.ncu-rep --page raw --csv 2>/dev/null | python - <<'PY'
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr, vals = rows[0], rows[-1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__occupancy_limit_registers", "smsp__cycles_active.avg",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
for h, v in zip(hdr, vals):
    if h in want or "stall" in h and "pct" in h or "issue_stalled" in h and "ratio" in h:
        print(h, v)
PY
