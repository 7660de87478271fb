timeout 600 python -m pytest tests/test_gpu_fcnet.py tests/test_gpu_search.py -m gpu -x -q 2>&1 | tail -2; echo
for i in 1 2; do timeout 300 python bench.py --no-conv --no-sweep --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['kernel_share']['fc_recurrent_us'], d['kernel_share']['tree_step_us'])"; done
timeout 200 python tests/tc_trace.py 0 2>&1 | grep -E "A1 gather|epi end|IN-SEARCH" | head -6
