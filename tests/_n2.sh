mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-conv > gpurun_out/r02ae_bench_n2.json 2> gpurun_out/r02ae_bench_n2.err; echo rc=$?; tail -3 gpurun_out/r02ae_bench_n2.err; head -c 300 gpurun_out/r02ae_bench_n2.json
