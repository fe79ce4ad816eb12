#!/bin/bash
# round 2, session 3, call 2: lane-per-observation linearisation + side-stream overlap: parity (full GPU suite), A/B bench, ncu
out=gpurun_out/r2b
mkdir -p $out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -30 > $out/c41_pytest_gpu.txt
tail -3 $out/c41_pytest_gpu.txt
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $out/c41_bench_new.json 2> $out/c41_bench_new.err
G2O_B200_OVERLAP=0 timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $out/c41_bench_nooverlap.json 2> $out/c41_bench_nooverlap.err
G2O_B200_LIN_PACKETS=0 G2O_B200_OVERLAP=0 timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $out/c41_bench_old.json 2> $out/c41_bench_old.err
python - <<'PY'
import json
for n in ("new", "nooverlap", "old"):
    try:
        d = json.loads(open("gpurun_out/r2b/c41_bench_%s.json" % n).read().strip().splitlines()[-1])
        print(n, round(d["value"], 1), d["ms_per_step"], round(d["e2e"]["value"], 1), d["kernel_groups_ms_per_10_iterations"])
    except Exception as e:
        print(n, "failed", e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ba_linearize_packets -c 1 -o $out/c41_ncu_linearize python tests/prof_run.py venice 1 > $out/c41_ncu.log 2>&1
ncu -i $out/c41_ncu_linearize.ncu-rep --page raw --csv > $out/c41_ncu_linearize_raw.csv 2>/dev/null
tail -2 $out/c41_ncu.log
