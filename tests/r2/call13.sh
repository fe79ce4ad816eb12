#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -q -x -k "tail_chain" > $O/c13_sanitizer_racecheck.txt 2>&1; tail -3 $O/c13_sanitizer_racecheck.txt
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -q -x -k "tail_chain" > $O/c13_sanitizer_memcheck.txt 2>&1; tail -3 $O/c13_sanitizer_memcheck.txt
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 > $O/c13_pytest_gpu.txt; cat $O/c13_pytest_gpu.txt
for w in ba10k venice; do G2O_B200_LIB=openslam_g2o_b200/libg2o_b200_timing.so timeout 600 python tests/chain_timing.py $w | grep -v "^ba10k\|^venice"; done 2>&1 | tee $O/c13_chain_timing.txt
for wl in venice ba10k; do
    timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 --no-extras --no-cpu-baseline --no-parallel-ordering > $O/c13_${wl}.json 2> $O/c13_${wl}.err
    python - <<PY
import json
d=json.loads([l for l in open("$O/c13_${wl}.json") if l.startswith("{")][-1])
print("$wl value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "chi2", d["chi2_first_run"][-1])
print({k: round(v["ms_total"],3) for k,v in d[[k for k in d if k.startswith("kernel_groups")][0]].items()})
PY
done
