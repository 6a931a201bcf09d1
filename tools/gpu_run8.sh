cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e > gpurun_out/r2/bench_N2b.json 2> gpurun_out/r2/bench_N2b.err; echo "bench N2 rc=$?"; python -c "
import json
d=json.load(open('gpurun_out/r2/bench_N2b.json')); print(2, d['ms_per_step'], d['value'], d['exchange'], d['roofline']['kernel_ms'], d['cpu_baseline']['parity_on_sample'])"
