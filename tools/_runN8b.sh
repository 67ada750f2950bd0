TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29541 bench.py --gpus 8 --steps 40 --warmup 5 --breakdown gpurun_out/r2i_step_breakdown_n8.txt > gpurun_out/r2i_bench_n8_b.json 2> gpurun_out/r2i_bench_n8_b.err
cat gpurun_out/r2i_step_breakdown_n8.txt | head -30
SLICED_DP_NOCOMM=1 timeout 200 $TR --master-port 29542 bench.py --gpus 8 --steps 40 --warmup 5 --no-parity-check > gpurun_out/r2i_bench_n8_nocomm.json 2> /dev/null
python - <<P
import json
for f in ('gpurun_out/r2i_bench_n8_b.json','gpurun_out/r2i_bench_n8_nocomm.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, 'ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], d['clocks'])
P
