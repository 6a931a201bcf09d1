cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2c
( time timeout 300 python -m pytest tests/test_wnm.py -m gpu -x -q ) > gpurun_out/r2c/pytest_wnm.log 2>&1
tail -15 gpurun_out/r2c/pytest_wnm.log
timeout 300 python tools/bench_wnm.py --traces 1000000 > gpurun_out/r2c/bench_wnm.json 2> gpurun_out/r2c/bench_wnm.err || tail -20 gpurun_out/r2c/bench_wnm.err
cat gpurun_out/r2c/bench_wnm.json
