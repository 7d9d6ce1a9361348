mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for tr in 1 3 5 7 8; do
(RNNT_TRIG=$tr RNNT_LSTM_V=2 RNNT_LSTM_DBG=1 timeout 90 python tools/lstm_check.py --oracle-rows 1) > gpurun_out/r2m_lstm_trig$tr.log 2>&1; echo "trig=$tr rc=$?"; grep "lstm_tc2 dbg\|lstm_tc2 loader" gpurun_out/r2m_lstm_trig$tr.log | tail -2 | cut -c1-330; grep -o '"encode_ms": [0-9.]*' gpurun_out/r2m_lstm_trig$tr.log
done
