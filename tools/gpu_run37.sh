cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2c
( time timeout 420 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2c/pytest_gpu3.log 2>&1
tail -6 gpurun_out/r2c/pytest_gpu3.log
