cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2b
( time timeout 900 python -m pytest tests/test_exchange_gpu.py -m gpu -x -q ) 2>&1 | tail -15
