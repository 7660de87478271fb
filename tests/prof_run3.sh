# Evidence run (one B200): GPU tests, smoke, both bench arms, ncu launch list + full captures.  bash tests/prof_run3.sh r02l
R=${1:-r02l}
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${R}_tests.log 2>&1; echo tests_rc=$?; tail -4 gpurun_out/${R}_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 700 python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err; echo bench_rc=$?; tail -3 gpurun_out/${R}_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_ref.json 2> gpurun_out/${R}_ref.err; echo ref_rc=$?
timeout 300 python tests/fused_sweep.py C4 C1 C2 C3 > gpurun_out/${R}_sweep.log 2>&1; tail -20 gpurun_out/${R}_sweep.log
NOX="--no-cpu-baseline --no-sweep --no-conv --no-f32"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 2 --warmup 1 $NOX > gpurun_out/${R}_ncu_bench.log 2>&1; echo ncu_rc=$?
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fc_search_kernel -s 3 -c 1 -o gpurun_out/${R}_fcsearch python bench.py --steps 1 --warmup 1 --no-graph $NOX > gpurun_out/${R}_ncu_fs.log 2>&1; echo ncu_rc=$?
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tree_step_w32 -s 60 -c 1 -o gpurun_out/${R}_tree python bench.py --steps 1 --warmup 1 --no-graph --fused 0 $NOX > gpurun_out/${R}_ncu_tree.log 2>&1; echo ncu_rc=$?
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fc_recurrent_tc -s 60 -c 1 -o gpurun_out/${R}_fc python bench.py --steps 1 --warmup 1 --no-graph --fused 0 $NOX > gpurun_out/${R}_ncu_fc.log 2>&1; echo ncu_rc=$?
timeout 200 ncu --set full --clock-control none --import-source on -k regex:build_targets_rows -s 74 -c 1 -o gpurun_out/${R}_targets_c3bulk python tests/targets_bench.py > gpurun_out/${R}_ncu_t.log 2>&1; echo ncu_rc=$?
