timeout 300 python -m pytest tests/test_gpu_muzero.py -m gpu -x -q -k "conv_pair or conv2d" 2>&1 | tail -3
for m in 2; do echo "--- mode $m"; MZ_CONV_PAIR=$m timeout 120 python tests/conv_bench.py 4096 2>&1 | tail -2; done
