# Multi-GPU bench lines (one box, N GPUs; gpurun --gpus N -- 'bash tests/prof_multi.sh N r02ae').
N=${1:-2}; R=${2:-r02ae}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --no-conv > gpurun_out/${R}_bench_n$N.json 2> gpurun_out/${R}_bench_n$N.err; echo rc=$?
tail -3 gpurun_out/${R}_bench_n$N.err; head -c 300 gpurun_out/${R}_bench_n$N.json
