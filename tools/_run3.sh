mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for m in 1 0; do
(RNNT_DSM_ASYNC=$m RNNT_LSTM_V=2 RNNT_LSTM_DBG=1 timeout 150 python tools/lstm_check.py --state) > gpurun_out/r2c_lstm_dsm$m.log 2>&1; echo "lstm dsm=$m rc=$?"; tail -2 gpurun_out/r2c_lstm_dsm$m.log
done
(timeout -k 5 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_r2.py -m gpu -q -x --timeout 150 -k "cfg2 or bench_workload or host_api or facade or reset") > gpurun_out/r2c_pytest_dec.log 2>&1; echo "pytest subset rc=$?"; tail -4 gpurun_out/r2c_pytest_dec.log
for m in 1 0; do
(RNNT_DSM_ASYNC=$m RNNT_DEC_DBG=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-extra --cpu-budget 1) > gpurun_out/r2c_bench_dsm$m.json 2> gpurun_out/r2c_bench_dsm$m.err; echo "bench dsm=$m rc=$?"; tail -2 gpurun_out/r2c_bench_dsm$m.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2c_bench_dsm$m.json').read().strip().split('\n')[-1])
    print(d['value'], d['ms_per_step'], d['e2e']['value'], d['stage_ms'])
except Exception as e: print('no bench', e)
PY
done
(timeout 300 python tools/batch_sweep.py --batches 32 64 128 256) > gpurun_out/r2c_sweep.json 2>&1; tail -1 gpurun_out/r2c_sweep.json
(RNNT_SUB32=1 timeout 300 python tools/batch_sweep.py --batches 64 128 256) > gpurun_out/r2c_sweep_sub32.json 2>&1; tail -1 gpurun_out/r2c_sweep_sub32.json
