cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
nvidia-smi topo -m | head -5
timeout 300 python -m pytest tests/test_exchange_gpu.py -x -q 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/exchange_worker.py > gpurun_out/r2/xworker2.log 2>&1; echo "worker rc=$?"; grep -v "^W\|^\[W" gpurun_out/r2/xworker2.log | tail -20
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2/bench_N2.json 2> gpurun_out/r2/bench_N2.err; echo "bench N2 rc=$?"; tail -3 gpurun_out/r2/bench_N2.err; cat gpurun_out/r2/bench_N2.json
