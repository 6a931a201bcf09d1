cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2d
( time timeout 22 python -m pytest "tests/test_kleene_star_gpu.py::test_one_kleene_star_state_without_constraints_closed_form[a b a*]" -x -q ) > gpurun_out/r2d/pytest_star.log 2>&1
tail -3 gpurun_out/r2d/pytest_star.log
