set -x
timeout 300 python -m pytest tests/test_gpu_targets.py tests/test_gpu_search.py -x -q 2>&1 | tail -8
timeout 120 python tests/targets_bench.py 2>&1 | grep -v "^ *\"\(kernel\|bound\|unit\|traffic\|peak\|peak_source\|workload\|algorithmic\)" | tail -40
timeout 120 python tests/tree_probe.py 1024 2>&1 | grep "games 1024"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tree_step_w32 -s 40 -c 1 -o gpurun_out/r01s_tree python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph --no-conv --no-sweep --streams 1 --games 1024 > gpurun_out/r01s_ncu_tree.log 2>&1; echo ncu_rc=$?
