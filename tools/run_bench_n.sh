# usage: bash tools/run_bench_n.sh N   — bench.py on N GPUs of one box under torchrun with an inner timeout (a hung collective must not eat the gpurun budget); line -> gpurun_out/bench_nN.json
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29531 bench.py --gpus $N --steps 40 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -c 300 gpurun_out/bench_n$N.err
python - <<P
import json
d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
print('N=$N ms', d['ms_per_step'], 'value', d['value'], 'e2e ms', d['e2e']['ms_per_step'], d['clocks'], d['parity_check'].get('dp'), d['config'].get('host_numa_node'))
P
