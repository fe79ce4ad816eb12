#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
timeout 600 python tests/config5_probe.py 300 500 2>&1 | tee $O/c20_config5_wide.jsonl | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['poses'], d.get('iteration_s'), d.get('phases_ms', {}).get('chol_factor_flow'), d.get('factor_tflops'), d.get('error'))
"
