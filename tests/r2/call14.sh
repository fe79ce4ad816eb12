#!/bin/bash
# every command under a SHORT timeout: a mis-counted named barrier hangs the kernel
O=gpurun_out/r2; mkdir -p $O
timeout 120 python -m pytest tests/test_gpu_parity.py -q -x -k "tail_chain" 2>&1 | tail -3 || exit 1
timeout 200 python -m pytest tests/test_gpu_parity.py -q -x -k "bundle or venice" 2>&1 | tail -3
for w in ba10k venice; do G2O_B200_LIB=openslam_g2o_b200/libg2o_b200_timing.so timeout 150 python tests/chain_timing.py $w | grep -v "^ba10k\|^venice\|backward"; done 2>&1 | tee $O/c14_chain_timing.txt
for wl in venice ba10k; do
    timeout 200 python bench.py --workload $wl --steps 20 --warmup 5 --no-extras --no-cpu-baseline --no-parallel-ordering > $O/c14_${wl}.json 2> $O/c14_${wl}.err
    python - <<PY
import json
d=json.loads([l for l in open("$O/c14_${wl}.json") if l.startswith("{")][-1])
print("$wl value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "chi2", d["chi2_first_run"][-1])
print({k: round(v["ms_total"],3) for k,v in d[[k for k in d if k.startswith("kernel_groups")][0]].items() if 'chol' in k or k=='factor'})
PY
done
