mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
(timeout -k 5 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_r2.py -m gpu -q -x --timeout 60 -k "cfg2 or bench_workload or host_api or facade or reset or modules") > gpurun_out/r2l_pytest_dec.log 2>&1; echo "pytest subset rc=$?"; tail -3 gpurun_out/r2l_pytest_dec.log | cut -c1-300
for i in 1 2; do
(RNNT_DEC_DBG=1 timeout 120 python bench.py --steps 20 --warmup 3 --no-extra --cpu-budget 1) > gpurun_out/r2l_bench$i.json 2> gpurun_out/r2l_bench$i.err; echo "bench $i rc=$?"; tail -1 gpurun_out/r2l_bench$i.err | cut -c1-400
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2l_bench$i.json').read().strip().split('\n')[-1])
    print(d['value'], d['ms_per_step'], d['e2e']['value'], d['stage_ms'])
except Exception as e: print('no bench', e)
PY
done
(timeout -k 5 400 python -m pytest tests -m gpu -q --timeout 120) > gpurun_out/r2l_pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -4 gpurun_out/r2l_pytest_all.log | cut -c1-300
(timeout 120 python __graft_entry__.py smoke) > gpurun_out/r2l_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2l_smoke.log | cut -c1-300
