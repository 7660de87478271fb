# Final-state evidence run (one B200): GPU tests, smoke, both bench arms, ncu launch list of the bench command,
# ncu --set full captures of the target kernel (bulk launch) and the network kernel.
R=${1:-r01s}
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${R}_tests.log 2>&1; echo tests_rc=$?; tail -3 gpurun_out/${R}_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err; echo bench_rc=$?; tail -3 gpurun_out/${R}_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_ref.json 2> gpurun_out/${R}_ref.err; echo ref_rc=$?
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-sweep --conv-games 1024 > gpurun_out/${R}_ncu_bench.log 2>&1; echo ncu_rc=$?
timeout 200 ncu --set full --clock-control none --import-source on -k regex:build_targets_rows -s 74 -c 1 -o gpurun_out/${R}_targets_c3bulk python tests/targets_bench.py > gpurun_out/${R}_ncu_t.log 2>&1; echo ncu_rc=$?
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fc_recurrent_tc -s 45 -c 1 -o gpurun_out/${R}_fc python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph --no-conv --no-sweep --streams 1 > gpurun_out/${R}_ncu_fc.log 2>&1; echo ncu_rc=$?
