mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
(RNNT_LSTM_V=2 RNNT_LSTM_DBG=1 timeout 90 python tools/lstm_check.py) > gpurun_out/r2i_lstm.log 2>&1; echo "lstm rc=$?"; grep "lstm_tc2\|lstm_v" gpurun_out/r2i_lstm.log | tail -5 | cut -c1-420
(RNNT_DEC_SPEC=3 timeout -k 5 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_r2.py -m gpu -q -x --timeout 60 -k "cfg2 or bench_workload or host_api or facade or reset or modules") > gpurun_out/r2i_pytest_spec3.log 2>&1; echo "pytest spec3 rc=$?"; tail -2 gpurun_out/r2i_pytest_spec3.log
for sp in 3 2; do
(RNNT_DEC_SPEC=$sp RNNT_DEC_DBG=1 timeout 120 python bench.py --steps 20 --warmup 3 --no-extra --cpu-budget 1) > gpurun_out/r2i_bench_spec$sp.json 2> gpurun_out/r2i_bench_spec$sp.err; echo "bench spec=$sp rc=$?"; tail -1 gpurun_out/r2i_bench_spec$sp.err | cut -c1-400
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2i_bench_spec$sp.json').read().strip().split('\n')[-1])
    print(d['value'], d['ms_per_step'], d['e2e']['value'], d['stage_ms'])
except Exception as e: print('no bench', e)
PY
done
(timeout 150 compute-sanitizer --tool memcheck --print-limit 20 python tools/lstm_check.py --B 8 --T 6 --oracle-rows 1) > gpurun_out/r2i_sanitizer_lstm.log 2>&1; echo "sanitizer lstm rc=$?"; grep -E "ERROR SUMMARY|Invalid|lstm_v" gpurun_out/r2i_sanitizer_lstm.log | head -8 | cut -c1-300
