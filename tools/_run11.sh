mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
(timeout 500 python bench.py --steps 20 --warmup 3) > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2k_bench.err | cut -c1-300
python - <<PY
import json
d=json.loads(open('gpurun_out/r2k_bench.json').read().strip().split('\n')[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['stage_ms'], d['clocks'])
print(json.dumps(d['extra'])[:2500])
print(d['cpu_baseline'])
PY
(timeout 200 python bench.py --impl reference --steps 3 --warmup 1) > gpurun_out/r2k_bench_ref.json 2>&1; tail -1 gpurun_out/r2k_bench_ref.json | cut -c1-400
