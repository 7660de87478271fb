set -x
timeout 600 python -m pytest tests/test_gpu_search.py tests/test_gpu_fcnet.py -x -q 2>&1 | tail -3
for lib in ab/lib_r152.so ab/lib_r144.so ab/lib_oldtree.so ab/lib_old.so ab/lib_r152.so; do
MZB200_LIB=$lib timeout 200 python bench.py --no-conv --no-cpu-baseline --steps 30 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['games_sweep']['16384']; print('lib $lib', 'value %.1fM'%(d['value']/1e6), 'e2e %.1fM'%(d['e2e']['value']/1e6), 'sweep16k %.1fM'%(k['expansions_per_s']/1e6), k['kernels']['tree_step_us'], k['kernels']['fc_recurrent_us'], 'C1 %.1fM C2 %.1fM C3 %.1fM'%(d['other_configs']['C1_tictactoe']['expansions_per_s']/1e6, d['other_configs']['C2_lunarlander']['expansions_per_s']/1e6, d['other_configs']['C3_breakout_ram']['expansions_per_s']/1e6), 'fc %.2f tree %.2f'%(d['kernel_share']['fc_recurrent_us'], d['kernel_share']['tree_step_us']))"
done
