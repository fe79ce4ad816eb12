#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3 | tee $O/c23_pytest_gpu.txt
timeout 900 python tests/config5_probe.py 200 700 1000 2>&1 | tee $O/c23_config5.jsonl | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['poses'], d.get('build_structure_s'), d.get('iteration_s'), d.get('phases_ms', {}), d.get('chi2'), d.get('error'))
"
nvidia-smi --query-gpu=memory.used --format=csv
