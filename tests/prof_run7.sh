# Evidence run, round 2 final state (one B200): GPU tests, smoke, both bench arms, ncu launch lists + full captures of the
# kernels bench.py reports rooflines for, of the learner kernels and of the float32-accurate tensor-core network kernel.
# bash tests/prof_run7.sh r02aw
R=${1:-r02aw}
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${R}_tests.log 2>&1; echo tests_rc=$?; tail -4 gpurun_out/${R}_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err; echo bench_rc=$?; tail -3 gpurun_out/${R}_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_ref.json 2> gpurun_out/${R}_ref.err; echo ref_rc=$?
NOX="--no-cpu-baseline --no-sweep --no-conv --no-f32 --no-selfplay --no-concurrent"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 2 --warmup 1 $NOX > gpurun_out/${R}_ncu_bench.log 2>&1; echo ncu_rc=$?
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fc_search_kernel -s 3 -c 1 -f -o gpurun_out/${R}_fcsearch python bench.py --steps 1 --warmup 1 --no-graph $NOX > gpurun_out/${R}_ncu_fs.log 2>&1; echo ncu_rc=$?
timeout 200 ncu --set full --clock-control none --import-source on -k regex:build_targets_tma -s 74 -c 1 -f -o gpurun_out/${R}_targets_c3bulk python tests/targets_bench.py > gpurun_out/${R}_ncu_t.log 2>&1; echo ncu_rc=$?
timeout 200 ncu --set full --clock-control none --import-source on -k "regex:build_targets_kernel" -s 3 -c 1 -f -o gpurun_out/${R}_targets_c2bulk python tests/targets_bench.py > gpurun_out/${R}_ncu_t2.log 2>&1; echo ncu_rc=$?
bash tests/learner_launches.sh ${R} 2>&1 | grep -E "chain|heads|pack|loss|adam|counter"
for k in chain_fwd chain_bwd heads_bwd heads_fwd; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$k -s 5 -c 1 -f -o gpurun_out/${R}_$k python /tmp/ll.py > /dev/null 2>&1
done
timeout 300 python tests/tf32_probe.py > gpurun_out/${R}_tf32_probe.log 2>&1; tail -6 gpurun_out/${R}_tf32_probe.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fc_recurrent_tf32x3 -s 70 -c 1 -f -o gpurun_out/${R}_tf32x3 python tests/tf32_probe.py > /dev/null 2>&1; echo ncu_rc=$?
timeout 300 python tests/fused_probe.py 4 1 > gpurun_out/${R}_fused_probe.log 2>&1; tail -12 gpurun_out/${R}_fused_probe.log
ls -la gpurun_out/${R}_*
