set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r01p_tests.log 2>&1; echo tests_rc=$?; tail -3 gpurun_out/r01p_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/r01p_bench.json 2> gpurun_out/r01p_bench.err; echo bench_rc=$?
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01p_ref.json 2> gpurun_out/r01p_ref.err; echo ref_rc=$?
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r01p_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-sweep --conv-games 1024 > gpurun_out/r01p_ncu_bench.log 2>&1; echo ncu_rc=$?
ls -la gpurun_out/ | tail -8
