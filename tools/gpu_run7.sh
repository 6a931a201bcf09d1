cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_exchange_gpu.py tests/test_pattern_compile_kat.py -x -q -m gpu 2>&1 | tail -4
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 tests/exchange_worker.py > gpurun_out/r2/xworker4.log 2>&1; echo "worker4 rc=$?"; grep "all ranks\|MISMATCH\|rror" gpurun_out/r2/xworker4.log | head -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 10 --warmup 3 --no-e2e > gpurun_out/r2/bench_N4.json 2> gpurun_out/r2/bench_N4.err; echo "bench N4 rc=$?"; python -c "
import json
d=json.load(open('gpurun_out/r2/bench_N4.json')); print(4, d['ms_per_step'], d['value'], d['exchange'], d['roofline']['kernel_ms'], d['cpu_baseline']['parity_on_sample'])"
