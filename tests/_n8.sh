mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --no-conv > gpurun_out/r02ae_bench_n8.json 2> gpurun_out/r02ae_bench_n8.err; echo rc=$?; tail -3 gpurun_out/r02ae_bench_n8.err; head -c 300 gpurun_out/r02ae_bench_n8.json
