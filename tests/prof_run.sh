set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r01o_tests.log 2>&1; echo tests_rc=$?; tail -3 gpurun_out/r01o_tests.log
timeout 400 python bench.py > gpurun_out/r01o_bench.json 2> gpurun_out/r01o_bench.err; echo bench_rc=$?
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_pair_tc -s 30 -c 1 -o gpurun_out/r01o_conv_pair python tests/conv_bench.py 4096 > gpurun_out/r01o_ncu_conv.log 2>&1; echo ncu1_rc=$?
timeout 300 ncu --set full --clock-control none --import-source on -k regex:unroll_loss_kernel -s 5 -c 1 -o gpurun_out/r01o_unroll_loss python bench.py --no-conv --no-sweep --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/r01o_ncu_loss.log 2>&1; echo ncu2_rc=$?
ls -la gpurun_out/
