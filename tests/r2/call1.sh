#!/bin/bash
# round 2, GPU call 1: box probes, FP64 peak, config 4 on one GPU, config 5 sizes (existing round-1 code)
O=gpurun_out/r2; mkdir -p $O
{ echo "nproc: $(nproc)"; free -g | head -2; nvidia-smi -L;
  echo "--- Eigen / SuiteSparse / BLAS probe (find /)";
  find / \( -path /proc -o -path /sys \) -prune -o \( -path '*Eigen/Core' -o -name 'cholmod.h' -o -name 'libcholmod*' -o -name 'libamd.*' -o -name 'amd.h' -o -name 'libsuitesparse*' -o -name 'libopenblas*' -o -name 'liblapack*' -o -name 'libblas.*' \) -print 2>/dev/null | head -40;
  echo "--- end"; } > $O/box_probe.txt 2>&1
python tests/fp64_peak.py > $O/fp64_peak.json 2> $O/fp64_peak.err
timeout 900 python bench.py --workload ba10k --steps 20 --warmup 3 > $O/c1_ba10k_n1.json 2> $O/c1_ba10k_n1.err
timeout 1500 python tests/config5_probe.py 200 300 500 700 1000 > $O/c1_config5.jsonl 2> $O/c1_config5.err
tail -c 1500 $O/fp64_peak.json $O/c1_config5.jsonl
