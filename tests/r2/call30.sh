#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
timeout 1500 python bench.py > $O/c30_bench_default.json 2> $O/c30_bench_default.err
tail -3 $O/c30_bench_default.err
python - <<PY
import json
d=json.loads([l for l in open("$O/c30_bench_default.json") if l.startswith("{")][-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "parity", d["parity"], "cpu", d["cpu_baseline"]["value"] if d.get("cpu_baseline") else None)
print("roofline", {k: d["roofline"][k] for k in ("bound","kernel","achieved","peak","frac","share_of_step")})
for k, v in (d.get("configs") or {}).items():
    print(k, v.get("value"), v.get("ms_per_step"), (v.get("e2e") or {}).get("value"), (v.get("parity") or {}).get("ok") if isinstance(v.get("parity"), dict) else v.get("parity"), v.get("error"))
PY
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > $O/c30_bench_reference.json 2> $O/c30_bench_reference.err; tail -c 600 $O/c30_bench_reference.json
