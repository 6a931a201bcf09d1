cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
export SIESTA_BENCH_DEBUG=1
for mode in NO_SAMPLER NO_POWER DEFAULT; do
export SIESTA_BENCH_NO_SAMPLER= SIESTA_BENCH_NO_POWER=
[ $mode = NO_SAMPLER ] && export SIESTA_BENCH_NO_SAMPLER=1
[ $mode = NO_POWER ] && export SIESTA_BENCH_NO_POWER=1
[ -z "$SIESTA_BENCH_NO_SAMPLER" ] && unset SIESTA_BENCH_NO_SAMPLER
[ -z "$SIESTA_BENCH_NO_POWER" ] && unset SIESTA_BENCH_NO_POWER
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 40 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2/scale_N8c.json 2> gpurun_out/r2/scale_N8c.err
python -c "
import json
d=json.load(open('gpurun_out/r2/scale_N8c.json')); print('$mode', d['ms_per_step'], d['latency_ms']['p50'], d['latency_ms']['max'], [x for x in d['latency_ms']['steps'] if x > 4])"
done
