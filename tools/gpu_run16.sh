cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
for n in 8 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 --no-e2e > gpurun_out/r2/scale_N$n.json 2> gpurun_out/r2/scale_N$n.err; echo "bench N$n rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/r2/scale_N$n.json')); print($n, d['ms_per_step'], d['value'], d['latency_ms'], d['exchange'], d['roofline']['kernel_ms'], d['cpu_baseline']['parity_on_sample'], d['clocks'])"
done
