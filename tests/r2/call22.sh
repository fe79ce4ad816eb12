#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
for cfg in "0 2e7" "0 5e6" "0 2e6" "0 5e5" "1 2e6"; do
set -- $cfg
G2O_B200_GROUP_SLACK=$1 G2O_B200_SUBTREE_MAX_FLOPS=$2 timeout 600 python tests/config5_probe.py 300 500 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('slack $1 submax $2', d['poses'], d.get('iteration_s'), d.get('phases_ms', {}).get('chol_factor_flow'), d['factor']['flow_tasks'], d.get('error'))
"
done 2>&1 | tee -a $O/c22_slack.txt
