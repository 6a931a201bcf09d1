cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 tests/exchange_worker.py > gpurun_out/r2/xworker8.log 2>&1; echo "worker8 rc=$?"; grep "OK\|MISMATCH\|rror" gpurun_out/r2/xworker8.log | grep "rank 0\|rank 7\|all ranks\|rror" | head -12
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2/bench_N8.json 2> gpurun_out/r2/bench_N8.err; echo "bench N8 rc=$?"; grep -v Warning gpurun_out/r2/bench_N8.err | tail -3; cat gpurun_out/r2/bench_N8.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 10 --warmup 3 --no-e2e > gpurun_out/r2/bench_N4.json 2> gpurun_out/r2/bench_N4.err; echo "bench N4 rc=$?"; python -c "
import json
for n in (4,8):
    d=json.load(open('gpurun_out/r2/bench_N%d.json'%n)); print(n, d['ms_per_step'], d['value'], d['exchange'], d['roofline']['kernel_ms'], d['cpu_baseline']['parity_on_sample'])"
