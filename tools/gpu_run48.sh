cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2c
timeout 200 python -m pytest tests/test_wnm.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/bench_wnm.py --traces 1000000 > gpurun_out/r2c/bench_wnm3.json 2> gpurun_out/r2c/bench_wnm3.err || tail -20 gpurun_out/r2c/bench_wnm3.err
cat gpurun_out/r2c/bench_wnm3.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"detect_nkp|detect_kernel|detect_long|gather_|scan_chunks|scan_top|set_tail|act_range" -c 60 --csv --log-file gpurun_out/r2c/launches_default.csv python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2c/ncu_default.log 2>&1
wc -l gpurun_out/r2c/launches_default.csv
