#!/bin/bash
# Round-end measurement on one B200 (final build of round 2): GPU tests, smoke, bench line (+ reference arm), ncu launch list of
# the same command, ncu --set full of the full-batch discretize_pair_kernel and of discretize_default_kernel, probes.
tag=${1:-r02zzz}
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_$tag.log
python -c "import __graft_entry__ as g; g.smoke()"
fi
timeout 300 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${tag}_ref.json 2>> gpurun_out/bench_$tag.err; echo "ref rc=$?"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_l_$tag.log 2>&1; echo "ncu list rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:discretize_pair_kernel --launch-skip 80 --launch-count 1 \
    --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum \
    -f -o gpurun_out/prof_pair_$tag python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/ncu_f_$tag.log 2>&1; echo "ncu pair rc=$?"
python scripts/ncu_summary.py gpurun_out/prof_pair_$tag.ncu-rep gpurun_out/prof_pair_$tag.txt > /dev/null 2>&1; rm -f gpurun_out/prof_pair_$tag.ncu-rep
timeout 200 ncu --set full --clock-control none --import-source on -k regex:discretize_default_kernel --launch-skip 1 --launch-count 1 \
    --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum \
    -f -o gpurun_out/prof_default_$tag python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/ncu_d_$tag.log 2>&1; echo "ncu default rc=$?"
python scripts/ncu_summary.py gpurun_out/prof_default_$tag.ncu-rep gpurun_out/prof_default_$tag.txt > /dev/null 2>&1; rm -f gpurun_out/prof_default_$tag.ncu-rep
python scripts/r02_probe_default_quick.py 2>&1 | tail -1 | tee gpurun_out/probe_default_$tag.txt
python scripts/r02_probe_prop_quick.py 2>&1 | tee gpurun_out/probe_prop_$tag.txt | tail -3
python scripts/r02_probe_sequence_quick.py 2>&1 | tee gpurun_out/probe_seq_$tag.txt | tail -4
