TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
i=0
for rep in 1 2; do
for d in 1 0; do
i=$((i+1))
SLICED_DP_DEFERRED=$d timeout 150 $TR --master-port $((29560+i)) bench.py --gpus 8 --steps 40 --warmup 5 --no-parity-check 2>/dev/null > gpurun_out/r2k_bench_n8_def${d}_$rep.json
python - <<P
import json
d=json.loads(open('gpurun_out/r2k_bench_n8_def${d}_$rep.json').read().strip().splitlines()[-1])
print('N=8 deferred=$d ms', round(d['ms_per_step'],4), 'e2e ms', round(d['e2e']['ms_per_step'],4), d['clocks']['sm_mhz'], d['training']['last_step_mean_loss'])
P
done; done | tee gpurun_out/r2k_deferred_ab_n8.txt
SLICED_DP_DEFERRED=1 timeout 150 $TR --master-port 29570 bench.py --gpus 8 --steps 40 --warmup 5 --breakdown gpurun_out/r2k_step_breakdown_n8_deferred.txt > gpurun_out/r2k_bench_n8.json 2>/dev/null
head -12 gpurun_out/r2k_step_breakdown_n8_deferred.txt
