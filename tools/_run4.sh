mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
(RNNT_LSTM_V=2 RNNT_LSTM_DBG=1 timeout 90 python tools/lstm_check.py --state) > gpurun_out/r2d_lstm.log 2>&1; echo "lstm rc=$?"; tail -3 gpurun_out/r2d_lstm.log | cut -c1-400
(timeout -k 5 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_r2.py -m gpu -q -x --timeout 60 -k "cfg2 or bench_workload or host_api or facade or reset") > gpurun_out/r2d_pytest_dec.log 2>&1; echo "pytest subset rc=$?"; tail -4 gpurun_out/r2d_pytest_dec.log
(RNNT_DEC_DBG=1 timeout 120 python bench.py --steps 10 --warmup 3 --no-extra --cpu-budget 1) > gpurun_out/r2d_bench_dbg.json 2> gpurun_out/r2d_bench_dbg.err; echo "bench dbg rc=$?"; tail -2 gpurun_out/r2d_bench_dbg.err | cut -c1-400
(timeout -k 5 200 python -m pytest tests/test_gpu_beam.py -m gpu -q -x --timeout 90) > gpurun_out/r2d_pytest_beam.log 2>&1; echo "pytest beam rc=$?"; tail -12 gpurun_out/r2d_pytest_beam.log | cut -c1-300
for occ in 1 2; do
(RNNT_FE_OCC=$occ timeout 120 python bench.py --steps 20 --warmup 3 --no-extra --cpu-budget 1) > gpurun_out/r2d_bench_fe$occ.json 2> gpurun_out/r2d_bench_fe$occ.err; echo "bench fe=$occ rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2d_bench_fe$occ.json').read().strip().split('\n')[-1])
    print(d['value'], d['ms_per_step'], d['e2e']['value'], d['stage_ms'])
except Exception as e: print('no bench', e)
PY
done
(timeout -k 5 400 python -m pytest tests -m gpu -q --timeout 120) > gpurun_out/r2d_pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -8 gpurun_out/r2d_pytest_all.log | cut -c1-300
