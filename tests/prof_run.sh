# Final-state evidence run (one B200): GPU tests, smoke, both bench arms, ncu launch list of the bench command.
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r01r_tests.log 2>&1; echo tests_rc=$?; tail -3 gpurun_out/r01r_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r01r_bench.json 2> gpurun_out/r01r_bench.err; echo bench_rc=$?; tail -3 gpurun_out/r01r_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01r_ref.json 2> gpurun_out/r01r_ref.err; echo ref_rc=$?
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r01r_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-sweep --conv-games 1024 > gpurun_out/r01r_ncu_bench.log 2>&1; echo ncu_rc=$?
