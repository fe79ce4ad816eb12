#!/bin/bash
# round 2, session 3, call 5: register budgets of the two BA linearisation kernels (CTAs per SM): A/B through the bench's kernel-group times
out=gpurun_out/r2b
mkdir -p $out
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras > $out/c44_$name.json 2> $out/c44_$name.err
}
run base G2O_B200_LIN_MINB=4
run lin5 G2O_B200_LIN_MINB=5
run lin6 G2O_B200_LIN_MINB=6
run cams4 G2O_B200_CAMS_MINB=4
run cams6 G2O_B200_CAMS_MINB=6
python - <<'PY'
import json
for n in ("base", "lin5", "lin6", "cams4", "cams6"):
    try:
        d = json.loads(open("gpurun_out/r2b/c44_%s.json" % n).read().strip().splitlines()[-1])
        kg = d["kernel_groups_ms_per_10_iterations"]
        print(n, round(d["value"], 1), round(d["ms_per_step"], 4), "lin", round(kg["linearize"]["ms_total"] / 10, 4), "cams", round(kg["linearize_cams"]["ms_total"] / 10, 4), "chi2", d["chi2_first_run"]["chi2"][-1] if "chi2_first_run" in d else None)
    except Exception as e:
        print(n, "failed", e)
PY
