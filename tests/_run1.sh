set -x
timeout 120 python tests/tc_trace.py 0 2>&1 | grep "d2_full\|A1 ready\|chunk  [0147]" | head -12
for st in 4 3 6 8; do
timeout 200 python bench.py --no-conv --no-cpu-baseline --no-sweep --streams $st --steps 30 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('streams', d['streams'], 'value %.1fM'%(d['value']/1e6), 'ms %.4f'%d['ms_per_step'], 'e2e %.1fM'%(d['e2e']['value']/1e6), 'fc %.2f tree %.2f'%(d['kernel_share']['fc_recurrent_us'], d['kernel_share']['tree_step_us']), d['clocks'])"
done
bash tests/prof_run.sh r01s
