mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
(timeout -k 5 200 python -m pytest tests/test_gpu_parity_r2.py -m gpu -q -x --timeout 100 -k "forward_lattice") > gpurun_out/r2o_pytest_fwd.log 2>&1; echo "pytest fwd rc=$?"; tail -12 gpurun_out/r2o_pytest_fwd.log | cut -c1-400
(timeout -k 5 400 python -m pytest tests -m gpu -q --timeout 120) > gpurun_out/r2o_pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -4 gpurun_out/r2o_pytest_all.log | cut -c1-300
(timeout 150 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 140 -k "test_offline_greedy_matches_reference_fixture and cfg2") > gpurun_out/r2o_sanitizer_decode.log 2>&1; echo "sanitizer decode rc=$?"; grep -E "ERROR SUMMARY|Invalid|passed|failed" gpurun_out/r2o_sanitizer_decode.log | head -6 | cut -c1-300
(timeout 500 python bench.py --steps 20 --warmup 3) > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2o_bench.json').read().strip().split('\n')[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['stage_ms'])
print(json.dumps(d['extra'])[:1500])
PY
