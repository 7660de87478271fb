mkdir -p gpurun_out
PYTORCH_NO_CUDA_MEMORY_CACHING=1 timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_fused_learner.py -m gpu -q -x -k "tc_head_kernels_ragged or chain_kernels" > gpurun_out/r02af_memcheck2.log 2>&1; echo rc=$?
grep -E "ERROR SUMMARY|Invalid|passed|failed" gpurun_out/r02af_memcheck2.log | head -20
