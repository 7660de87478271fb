# Checkpoint run (one B200): GPU tests, smoke, both bench arms.  bash tests/prof_run4.sh r02v
R=${1:-r02v}
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${R}_tests.log 2>&1; echo tests_rc=$?; tail -4 gpurun_out/${R}_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err; echo bench_rc=$?; tail -3 gpurun_out/${R}_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_ref.json 2> gpurun_out/${R}_ref.err; echo ref_rc=$?
