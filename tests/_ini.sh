mkdir -p gpurun_out
NOX="--no-cpu-baseline --no-sweep --no-conv --no-f32 --no-selfplay --no-concurrent"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fc_recurrent_tc -s 0 -c 3 -f -o gpurun_out/r02ag_initial python bench.py --steps 1 --warmup 1 --no-graph $NOX > gpurun_out/r02ag_ncu.log 2>&1; echo rc=$?
