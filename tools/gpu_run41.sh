cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2c
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gather_kernel -s 4 -c 1 -o gpurun_out/r2c/prof_gather -f python bench.py --workload detection_gap6_4Mx50 --steps 3 --no-cpu-baseline --no-e2e > gpurun_out/r2c/ncu_gather.log 2>&1
tail -3 gpurun_out/r2c/ncu_gather.log
ls -la gpurun_out/r2c/*.ncu-rep
