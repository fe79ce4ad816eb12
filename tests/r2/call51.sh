#!/bin/bash
# round 2, session 3: compute-sanitizer racecheck (shared-memory hazards) over smoke(): packets linearisation, sparse-inverse row / diagonal kernels
out=gpurun_out/r2b
mkdir -p $out
timeout 280 compute-sanitizer --tool racecheck --print-limit 5 python -c "import __graft_entry__ as g; g.smoke()" > $out/c51_racecheck_smoke.txt 2>&1
tail -5 $out/c51_racecheck_smoke.txt
