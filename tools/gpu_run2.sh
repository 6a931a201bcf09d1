set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest2.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2/pytest2.log
for i in 1 2; do timeout 300 python bench.py --workload detection_gap6_4Mx50 --steps 20 --no-cpu-baseline --no-e2e > gpurun_out/r2/bench_gap6_4M_p$i.json 2>gpurun_out/r2/bench_gap6_4M_p$i.err; python -c "import json;d=json.load(open('gpurun_out/r2/bench_gap6_4M_p$i.json'));print(d['roofline']['kernel_ms'],d['roofline']['frac'],d['ms_per_step'],d['roofline']['all_kernels_ms'])"; done
SIESTA_NKP_TILE_BATCH=1 timeout 300 python bench.py --workload detection_gap6_4Mx50 --steps 20 --no-cpu-baseline --no-e2e > gpurun_out/r2/bench_gap6_4M_pb1.json 2>&1; python -c "import json;d=json.load(open('gpurun_out/r2/bench_gap6_4M_pb1.json'));print('batch1',d['roofline']['kernel_ms'],d['roofline']['frac'],d['ms_per_step'])"
SIESTA_K1_CTAS_PER_SM=5 timeout 300 python bench.py --workload detection_gap6_4Mx50 --steps 20 --no-cpu-baseline --no-e2e > gpurun_out/r2/bench_gap6_4M_p5.json 2>&1; python -c "import json;d=json.load(open('gpurun_out/r2/bench_gap6_4M_p5.json'));print('5cta',d['roofline']['kernel_ms'],d['roofline']['frac'],d['ms_per_step'])"
timeout 600 python bench.py --no-e2e > gpurun_out/r2/bench_default_p.json 2> gpurun_out/r2/bench_default_p.err; python -c "import json;d=json.load(open('gpurun_out/r2/bench_default_p.json'));print('100M',d['roofline']['kernel_ms'],d['roofline']['frac'],d['ms_per_step'],d['value'],d['cpu_baseline'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:detect_nkp -s 4 -c 1 -o gpurun_out/r2/prof_r02_nkp_b -f python bench.py --workload detection_gap6_4Mx50 --steps 3 --no-cpu-baseline --no-e2e > gpurun_out/r2/ncu_b.log 2>&1; tail -2 gpurun_out/r2/ncu_b.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2/launches_gap6_4M.csv python bench.py --workload detection_gap6_4Mx50 --steps 2 --no-cpu-baseline --no-e2e > gpurun_out/r2/ncu_l.log 2>&1; tail -1 gpurun_out/r2/ncu_l.log
