set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
nvidia-smi --query-gpu=name,memory.total --format=csv
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest1.log 2>&1; echo "pytest rc=$?" 
tail -5 gpurun_out/r2/pytest1.log
timeout 300 python bench.py --workload detection_gap6_4Mx50 --steps 20 > gpurun_out/r2/bench_gap6_4M.json 2> gpurun_out/r2/bench_gap6_4M.err; echo "rc=$?"; cat gpurun_out/r2/bench_gap6_4M.json; tail -3 gpurun_out/r2/bench_gap6_4M.err
SIESTA_K1_CTAS_PER_SM=5 timeout 300 python bench.py --workload detection_gap6_4Mx50 --steps 20 --no-cpu-baseline --no-e2e > gpurun_out/r2/bench_gap6_4M_5cta.json 2>&1; cat gpurun_out/r2/bench_gap6_4M_5cta.json
SIESTA_NKP_TILE_BATCH=1 timeout 300 python bench.py --workload detection_gap6_4Mx50 --steps 20 --no-cpu-baseline --no-e2e > gpurun_out/r2/bench_gap6_4M_b1.json 2>&1; cat gpurun_out/r2/bench_gap6_4M_b1.json
timeout 600 python bench.py > gpurun_out/r2/bench_default.json 2> gpurun_out/r2/bench_default.err; echo "rc=$?"; cat gpurun_out/r2/bench_default.json; tail -3 gpurun_out/r2/bench_default.err
timeout 300 python bench.py --workload detection_kleene_1Mx100 --steps 20 > gpurun_out/r2/bench_kleene.json 2> gpurun_out/r2/bench_kleene.err; cat gpurun_out/r2/bench_kleene.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:detect_nkp -s 4 -c 1 -o gpurun_out/r2/prof_r02_nkp_a -f python bench.py --workload detection_gap6_4Mx50 --steps 3 --no-cpu-baseline --no-e2e > gpurun_out/r2/ncu_a.log 2>&1; tail -2 gpurun_out/r2/ncu_a.log
