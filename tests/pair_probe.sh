timeout 300 python -m pytest tests/test_gpu_muzero.py -m gpu -x -q 2>&1 | tail -4
for m in 0 2; do echo "--- mode $m"; MZ_CONV_PAIR=$m timeout 120 python tests/conv_bench.py 4096 2>&1 | tail -2; done
echo "--- mode 2 1024"; MZ_CONV_PAIR=2 timeout 120 python tests/conv_bench.py 1024 2>&1 | tail -2
