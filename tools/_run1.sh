mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
(RNNT_LSTM_V=2 RNNT_LSTM_DBG=1 timeout 150 python tools/lstm_check.py --out gpurun_out/enc_v2.npy --state) > gpurun_out/r2_lstm_v2.log 2>&1; echo "v2 rc=$?"
(RNNT_LSTM_V=1 RNNT_LSTM_DBG=1 timeout 150 python tools/lstm_check.py --out gpurun_out/enc_v1.npy --compare gpurun_out/enc_v2.npy) > gpurun_out/r2_lstm_v1.log 2>&1; echo "v1 rc=$?"
(RNNT_LSTM_V=2 timeout 150 python tools/lstm_check.py --B 20 --T 37 --ragged --oracle-rows 20) > gpurun_out/r2_lstm_v2_ragged.log 2>&1; echo "v2 ragged rc=$?"
tail -3 gpurun_out/r2_lstm_v2.log gpurun_out/r2_lstm_v1.log gpurun_out/r2_lstm_v2_ragged.log
(timeout 300 python bench.py --steps 10 --warmup 3) > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/r2_bench_a.json
(timeout 600 python -m pytest tests -m gpu -x -q) > gpurun_out/r2_pytest_a.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest_a.log
