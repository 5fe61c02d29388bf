#!/bin/bash
# Round-end measurement on one B200 (run through gpurun): GPU tests, bench line, ncu launch list of the same command,
# one ncu --set full capture of the discretization kernel (a window of the overlapped pass and the full-batch launch).
# usage: scripts/final_gpu_run.sh <tag>
tag=${1:-r1x}
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$tag.log
timeout 200 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
timeout 120 python bench.py --no-overlap --no-cpu-baseline > gpurun_out/bench_${tag}_b2b.json 2>> gpurun_out/bench_$tag.err; echo "bench b2b rc=$?"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_l_$tag.log 2>&1; echo "ncu list rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:discretize_pair_kernel --launch-skip 78 --launch-count 3 \
    -f -o gpurun_out/prof_$tag python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_f_$tag.log 2>&1; echo "ncu full rc=$?"
python -c "
import json,sys
for f in ('gpurun_out/bench_$tag.json','gpurun_out/bench_${tag}_b2b.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['kernel']['discretize_ms'], d['kernel']['propagate_ms'], d['roofline']['frac'], d['roofline']['in_step']['frac'], d['gpu_launches'], d['clocks'])
"
