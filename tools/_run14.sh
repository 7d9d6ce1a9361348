mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
(timeout -k 5 400 python -m pytest tests -m gpu -q -x --timeout 120) > gpurun_out/r2n_pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -4 gpurun_out/r2n_pytest_all.log | cut -c1-300
(timeout 120 python tools/stream_bench.py --seconds 20) > gpurun_out/r2n_stream_default.json 2>&1; tail -1 gpurun_out/r2n_stream_default.json | cut -c1-420
(RNNT_SUB32=1 timeout 120 python tools/stream_bench.py --seconds 20) > gpurun_out/r2n_stream_sub32.json 2>&1; tail -1 gpurun_out/r2n_stream_sub32.json | cut -c1-420
(RNNT_SUB32=1 timeout -k 5 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_r2.py -m gpu -q -x --timeout 100 -k "stream_session") > gpurun_out/r2n_pytest_sub32.log 2>&1; echo "pytest sub32 rc=$?"; tail -2 gpurun_out/r2n_pytest_sub32.log | cut -c1-300
(RNNT_SUB32=1 timeout 300 python tools/batch_sweep.py --batches 64 256) > gpurun_out/r2n_sweep_sub32.json 2>&1; tail -1 gpurun_out/r2n_sweep_sub32.json | cut -c1-500
