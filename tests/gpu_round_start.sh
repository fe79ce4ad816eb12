#!/bin/bash
# One gpurun call that re-establishes the measured state of the repository on a fresh B200 box (start of a round, or
# after host-side changes).  Everything lands in gpurun_out/; copy what is to be judged into profiles/.
#
#   gpurun --timeout 900 -- 'bash tests/gpu_round_start.sh r02'
#
# ~6 GPU-minutes: full GPU suite, the four bench workloads (Venice carries the CPU leg), loader thread scaling,
# one ncu launch list per headline workload.
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
python -m pytest tests -q -m gpu -x 2>&1 | tail -5 | tee $out/${tag}_pytest_gpu.txt
python bench.py --steps 100 --warmup 10 > $out/${tag}_bench_venice.json 2> $out/${tag}_bench_venice.err
python bench.py --workload sphere2500 --steps 100 --warmup 10 --no-cpu-baseline > $out/${tag}_bench_sphere.json 2> $out/${tag}_bench_sphere.err
python bench.py --workload ba10k --steps 20 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_ba10k.json 2> $out/${tag}_bench_ba10k.err
python bench.py --workload sphere40k --steps 20 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_sphere40k.json 2> $out/${tag}_bench_sphere40k.err
python tests/summarize_bench.py $out/${tag}_bench_*.json | tee $out/${tag}_bench_summary.txt
G2O_B200_LOADER_VERBOSE=1 python tests/loader_scaling.py 120000 > $out/${tag}_loader_scaling.txt 2>&1
for wl in venice sphere2500; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches_${wl}.csv \
      python tests/prof_run.py $wl 3 > $out/${tag}_ncu_${wl}.log 2>&1
done
tail -3 $out/${tag}_loader_scaling.txt
