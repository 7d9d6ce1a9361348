mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
(RNNT_LSTM_V=2 RNNT_LSTM_DBG=1 timeout 150 python tools/lstm_check.py --state) > gpurun_out/r2b_lstm_v2.log 2>&1; echo "lstm v2 rc=$?"; tail -2 gpurun_out/r2b_lstm_v2.log
(timeout -k 5 500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 150 -k "cfg2 or bench_workload or host_api or facade") > gpurun_out/r2b_pytest_dec.log 2>&1; echo "pytest decode subset rc=$?"; tail -8 gpurun_out/r2b_pytest_dec.log
(RNNT_DEC_DBG=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-extra --cpu-budget 2) > gpurun_out/r2b_bench_dbg.json 2> gpurun_out/r2b_bench_dbg.err; echo "bench dbg rc=$?"; tail -3 gpurun_out/r2b_bench_dbg.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2b_bench_dbg.json').read().strip().split('\n')[-1])
    print(d['value'], d['ms_per_step'], d['e2e']['value'], d['stage_ms'])
except Exception as e: print('no bench', e)
PY
(timeout -k 5 900 python -m pytest tests -m gpu -q --timeout 300) > gpurun_out/r2b_pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -15 gpurun_out/r2b_pytest_all.log
(timeout 400 python bench.py --steps 20 --warmup 3) > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/r2b_bench.json
