#!/bin/bash
# config 4 (10k cameras / 2M points / 20M observations) landmark-sharded over all GPUs of the box
O=gpurun_out/r2; mkdir -p $O
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload ba10k --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $O/c25_ba10k_n$N.json 2> $O/c25_ba10k_n$N.err
tail -3 $O/c25_ba10k_n$N.err
python - <<PY
import json
d=json.loads([l for l in open("$O/c25_ba10k_n$N.json") if l.startswith("{")][-1])
print("ba10k N=$N value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "chi2", d.get("chi2_first_run", [None])[-1])
print({k: round(v["ms_total"],3) for k,v in d[[k for k in d if k.startswith("kernel_groups")][0]].items()})
print(d.get("parallel_ordering"))
PY
