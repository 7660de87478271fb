timeout 300 python -m pytest tests/test_gpu_muzero.py -m gpu -x -q 2>&1 | tail -3
echo "--- staged epilogue"; timeout 120 python tests/conv_bench.py 4096 2>&1 | tail -2; timeout 120 python tests/conv_bench.py 1024 2>&1 | tail -2
