mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
(timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -x --timeout 280) > gpurun_out/r2j_pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -3 gpurun_out/r2j_pytest_multi.log | cut -c1-300
(timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3) > gpurun_out/r2j_bench_2gpu.json 2> gpurun_out/r2j_bench_2gpu.err; echo "bench 2gpu rc=$?"; grep -i "warn\|error" gpurun_out/r2j_bench_2gpu.err | head -3 | cut -c1-200
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2j_bench_2gpu.json').read().strip().split('\n') if l.startswith('{')][-1])
    print(d['value'], d['ms_per_step'], d['e2e']['value'], d['per_rank_ms_per_step'], json.dumps(d['extra'])[:700])
except Exception as e: print('no bench', e)
PY
