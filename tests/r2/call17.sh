#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chol_factor_flow -c 1 -o $O/c17_ncu_wide python tests/prof_run.py sphere300 1 > $O/c17_ncu.log 2>&1
tail -3 $O/c17_ncu.log
ncu -i $O/c17_ncu_wide.ncu-rep --page raw --csv > $O/c17_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open("gpurun_out/r2/c17_raw.csv")))
h=rows[0]; u=rows[1]; v=rows[2]
want=["gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","lts__t_sector_hit_rate.pct","lts__t_bytes.sum","l1tex__t_bytes.sum","sm__inst_executed_pipe_fp64.sum","sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active","smsp__issue_active.avg.pct","sm__warps_active.avg.pct_of_peak_sustained_active","dram__throughput.avg.pct_of_peak_sustained_elapsed","lts__t_sectors_srcunit_tex_op_read.sum","lts__t_sectors_srcunit_tex_lookup_miss.sum","l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum","smsp__inst_executed_pipe_fp64_op_dmma.sum"]
for i,n in enumerate(h):
    if n in want or "dmma" in n.lower() or "fp64" in n.lower() or ("lts__t_sector" in n and "hit" in n): print(n,u[i],v[i])
PY
