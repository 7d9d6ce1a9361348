mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
(RNNT_LSTM_V=2 RNNT_LSTM_DBG=1 timeout 90 python tools/lstm_check.py) > gpurun_out/r2g_lstm.log 2>&1; echo "lstm rc=$?"; grep "lstm_tc2" gpurun_out/r2g_lstm.log | tail -4 | cut -c1-500
(timeout -k 5 120 python -m pytest tests/test_gpu_parity_r2.py -m gpu -q -x --timeout 60 -k "bundle or reset") > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2g_pytest.log
# launch list of the bench step (cold-cache, serialised: shares only)
(timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --profile --steps 2 --warmup 1) > gpurun_out/r2g_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
for k in lstm_layer_tc2_kernel decode_tc2_kernel; do
(timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/prof_r2_$k -f python bench.py --profile --steps 1 --warmup 1) > gpurun_out/r2g_ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
ls -la gpurun_out/*.ncu-rep | tail -4
