mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
(timeout 500 python bench.py --steps 20 --warmup 3) > gpurun_out/r3_bench.json 2> gpurun_out/r3_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r3_bench.json').read().strip().split('\n')[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['stage_ms'], d['clocks'])
print(json.dumps(d['extra'])[:1800])
PY
(RNNT_NO_COOP=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3_launches.csv python bench.py --profile --steps 2 --warmup 1 --no-extra) > gpurun_out/r3_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
(RNNT_NO_COOP=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:decode_tc2_kernel -s 1 -c 1 -f -o gpurun_out/r3_ncu_decode python bench.py --profile --steps 1 --warmup 1 --no-extra) > gpurun_out/r3_ncu_decode.log 2>&1; echo "ncu decode rc=$?"
(RNNT_NO_COOP=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:lstm_layer_tc2_kernel -s 5 -c 1 -f -o gpurun_out/r3_ncu_lstm python bench.py --profile --steps 1 --warmup 1 --no-extra) > gpurun_out/r3_ncu_lstm.log 2>&1; echo "ncu lstm rc=$?"
for k in decode lstm; do ncu -i gpurun_out/r3_ncu_$k.ncu-rep --page raw --csv > gpurun_out/r3_ncu_full_$k.csv 2>/dev/null; ls -la gpurun_out/r3_ncu_$k.ncu-rep gpurun_out/r3_ncu_full_$k.csv; done
