#!/bin/bash
# round 2, GPU call 5: tail chain v2 (descriptor ring, front-resident backward sweep, stores off the critical path)
O=gpurun_out/r2; mkdir -p $O
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -q -x -k "tail_chain" > $O/c5_sanitizer_memcheck.txt 2>&1; tail -3 $O/c5_sanitizer_memcheck.txt
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -q -x -k "tail_chain" > $O/c5_sanitizer_racecheck.txt 2>&1; tail -3 $O/c5_sanitizer_racecheck.txt
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 > $O/c5_pytest_gpu.txt; cat $O/c5_pytest_gpu.txt
for wl in venice ba10k; do
  for ch in 1; do
    G2O_B200_CHAIN=$ch timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 --no-extras --no-cpu-baseline --no-parallel-ordering > $O/c5_${wl}_chain$ch.json 2> $O/c5_${wl}_chain$ch.err
    python - <<PY
import json
d=json.loads([l for l in open("$O/c5_${wl}_chain$ch.json") if l.startswith("{")][-1])
print("$wl chain=$ch value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "chi2", d["chi2_first_run"][-1])
print({k: round(v["ms_total"],3) for k,v in d[[k for k in d if k.startswith("kernel_groups")][0]].items()})
PY
  done
done
