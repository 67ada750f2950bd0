set -x
python -m pytest tests/test_gpu_gemm.py -m gpu -x -q > gpurun_out/r2c_tests_gemm.log 2>&1; tail -3 gpurun_out/r2c_tests_gemm.log
python -m pytest tests/test_gpu_mlp.py -m gpu -x -q > gpurun_out/r2c_tests_mlp.log 2>&1; tail -3 gpurun_out/r2c_tests_mlp.log
for i in 1 2; do
for s in 0 1; do
SLICED_GEMM_SCHED=$s python bench.py --steps 20 --no-sweeps --no-cpu-baseline --no-parity-check 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('SCHED=$s N=1 ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['clocks'])"
SLICED_GEMM_SCHED=$s python bench.py --steps 40 --debug-per-gpu-batch 8192 --no-sweeps --no-cpu-baseline --no-parity-check 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('SCHED=$s b8192 ms', d['ms_per_step'])"
done; done
