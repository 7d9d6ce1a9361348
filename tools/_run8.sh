mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
(RNNT_LSTM_V=2 RNNT_LSTM_DBG=1 timeout 90 python tools/lstm_check.py --state) > gpurun_out/r2h_lstm.log 2>&1; echo "lstm rc=$?"; grep "lstm_tc2\|lstm_v" gpurun_out/r2h_lstm.log | tail -5 | cut -c1-500
(timeout -k 5 400 python -m pytest tests -m gpu -q -x --timeout 120) > gpurun_out/r2h_pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -5 gpurun_out/r2h_pytest_all.log | cut -c1-300
(RNNT_DEC_DBG=1 timeout 120 python bench.py --steps 20 --warmup 3 --no-extra --cpu-budget 1) > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench rc=$?"; tail -1 gpurun_out/r2h_bench.err | cut -c1-400
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2h_bench.json').read().strip().split('\n')[-1])
    print(d['value'], d['ms_per_step'], d['e2e']['value'], d['stage_ms'])
except Exception as e: print('no bench', e)
PY
export RNNT_NO_COOP=1
(timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --profile --steps 2 --warmup 1) > gpurun_out/r2h_ncu_launch.log 2>&1; echo "ncu launches rc=$?"; grep -c "tc2_kernel" gpurun_out/r2_launches.csv
for k in lstm_layer_tc2_kernel decode_tc2_kernel; do
(timeout 420 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/prof_r2_$k -f python bench.py --profile --steps 1 --warmup 1) > gpurun_out/r2h_ncu_$k.log 2>&1; echo "ncu $k rc=$?"; tail -3 gpurun_out/r2h_ncu_$k.log | cut -c1-200
done
ls -la gpurun_out/*.ncu-rep | tail -4
