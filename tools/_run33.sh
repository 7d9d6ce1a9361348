export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
run() { echo "== $*"; (env "$@" timeout -k 5 100 python tools/stress.py) 2>&1 | grep "OK\|STALL\|Error\|error\|total" | cut -c1-260; }
run STRESS_MODE=pipe STRESS_CALLS=1500
run STRESS_MODE=pipe STRESS_HOST=1 STRESS_CALLS=1500
run STRESS_MODE=pipe STRESS_HOST=1 RNNT_PIPE_PRIO=0 STRESS_CALLS=600
run STRESS_MODE=pipe RNNT_PIPE_PRIO=0 STRESS_CALLS=600
run STRESS_MODE=pipe STRESS_HOST=1 RNNT_PIPE_SAME=1 STRESS_CALLS=600
(timeout -k 5 300 python -m pytest tests/test_gpu_parity_r2.py -m gpu -q -x -k "pipeline or soak") > gpurun_out/r3b_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r3b_pytest.log | cut -c1-300
