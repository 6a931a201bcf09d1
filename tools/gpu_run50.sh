cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2c
timeout 200 python -m pytest tests/test_detect_gpu.py -m gpu -x -q -k "byte_activity or pinned or chunks" 2>&1 | tail -3
N=2
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tools/bench_declare_multi.py > gpurun_out/r2c/declare_N2.json 2> gpurun_out/r2c/declare_N2.err || tail -20 gpurun_out/r2c/declare_N2.err
grep "^{" gpurun_out/r2c/declare_N2.json
timeout 300 python tools/bench_declare_multi.py > gpurun_out/r2c/declare_N1.json 2> gpurun_out/r2c/declare_N1.err || tail -20 gpurun_out/r2c/declare_N1.err
grep "^{" gpurun_out/r2c/declare_N1.json
