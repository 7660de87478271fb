mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fcnet.py tests/test_gpu_search.py tests/test_selfplay.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python bench.py --no-cpu-baseline --no-sweep --no-conv --no-f32 --no-selfplay --no-concurrent > gpurun_out/r02ah_bench.json 2> gpurun_out/r02ah_bench.err; echo rc=$?
python - <<'P'
import json
d=json.load(open('gpurun_out/r02ah_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['kernel_share'])
P
