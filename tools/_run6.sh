mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi -L | head -3
(timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/nccl_shard_check.py --n 70) > gpurun_out/r2f_shard_check.log 2>&1; echo "shard check rc=$?"; tail -2 gpurun_out/r2f_shard_check.log | cut -c1-400
(timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3) > gpurun_out/r2f_bench_2gpu.json 2> gpurun_out/r2f_bench_2gpu.err; echo "bench 2gpu rc=$?"; tail -3 gpurun_out/r2f_bench_2gpu.err | cut -c1-300
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2f_bench_2gpu.json').read().strip().split('\n') if l.startswith('{')][-1])
    print(d['value'], d['ms_per_step'], d['e2e']['value'], d['per_rank_ms_per_step'], json.dumps(d['extra'])[:900])
except Exception as e: print('no bench', e)
PY
(timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1) > gpurun_out/r2f_bench_ref_2gpu.json 2>&1; echo "ref arm rc=$?"; tail -1 gpurun_out/r2f_bench_ref_2gpu.json | cut -c1-300
