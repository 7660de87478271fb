# Learner iteration (one B200): fused-learner tests + steps/s of every learner variant.  bash tests/prof_run5.sh r02y
R=${1:-r02y}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fused_learner.py -m gpu -q -s > gpurun_out/${R}_tests.log 2>&1; echo tests_rc=$?; tail -25 gpurun_out/${R}_tests.log
timeout 300 python tests/learner_fused_one.py 2>&1 | tail -6
