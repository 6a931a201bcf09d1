cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2d
( time timeout 45 python -m pytest tests/test_exchange_gpu.py -x -q -W always ) > gpurun_out/r2d/pytest_xchg2.log 2>&1
tail -6 gpurun_out/r2d/pytest_xchg2.log
( time timeout 14 python -m pytest tests/test_kleene_star_gpu.py -x -q -k "a b a" ) > gpurun_out/r2d/pytest_star.log 2>&1
tail -3 gpurun_out/r2d/pytest_star.log
