export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
(timeout -k 5 500 python -m pytest tests -m gpu -q -x) > gpurun_out/r3c_pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -3 gpurun_out/r3c_pytest_all.log | cut -c1-300
(BENCH_TRACE=1 timeout 500 python bench.py --steps 20 --warmup 3) > gpurun_out/r3c_bench.json 2> gpurun_out/r3c_bench.err; echo "bench rc=$?"; grep "pipelined" gpurun_out/r3c_bench.err | cut -c1-200
python - <<PY
import json
d=json.loads(open('gpurun_out/r3c_bench.json').read().strip().split('\n')[-1])
print(d['value'], d['ms_per_step'], d['config']['batches_in_flight'], d['config']['sequential_ms_per_step'], d['e2e'], d['stage_ms'], d['clocks'])
PY
