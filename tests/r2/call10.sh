#!/bin/bash
# round 2, GPU call 10: chi2 pass skipped when the state's chi2 is known; fused trial tail (A/B with G2O_B200_FUSE)
O=gpurun_out/r2; mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 > $O/c10_pytest_gpu.txt; cat $O/c10_pytest_gpu.txt
for wl in venice ba10k; do
  for f in 1 0; do
    G2O_B200_FUSE=$f timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 --no-extras --no-cpu-baseline --no-parallel-ordering > $O/c10_${wl}_fuse$f.json 2> $O/c10_${wl}_fuse$f.err
    python - <<PY
import json
d=json.loads([l for l in open("$O/c10_${wl}_fuse$f.json") if l.startswith("{")][-1])
print("$wl fuse=$f value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "chi2", d["chi2_first_run"][-1], "launches", d["gpu_launches"])
print({k: round(v["ms_total"],3) for k,v in d[[k for k in d if k.startswith("kernel_groups")][0]].items()})
PY
  done
done
