cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 40 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2/scale_N8d.json 2> gpurun_out/r2/scale_N8d.err
python -c "
import json
d=json.load(open('gpurun_out/r2/scale_N8d.json')); print('40 steps', d['ms_per_step'], d['latency_ms']['p50'], d['latency_ms']['max'], [x for x in d['latency_ms']['steps'] if x > 4], d['exchange'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2/scale_N8e.json 2> gpurun_out/r2/scale_N8e.err
python -c "
import json
d=json.load(open('gpurun_out/r2/scale_N8e.json')); print('10 steps', d['ms_per_step'], d['value'], d['latency_ms'], d['cpu_baseline']['parity_on_sample'], d['e2e']['value'], d['clocks'])"
