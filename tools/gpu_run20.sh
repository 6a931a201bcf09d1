cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 600 python bench.py --workload detection_gap6_4Mx50 --steps 10 --e2e-steps 5 > gpurun_out/r2/bench_gap6_4M_z.json 2> gpurun_out/r2/bench_gap6_4M_z.err; python -c "import json;d=json.load(open('gpurun_out/r2/bench_gap6_4M_z.json'));print('4M',d['roofline']['kernel_ms'],d['ms_per_step'],d['e2e'],d['cpu_baseline'])"
SIESTA_NO_TS_ZERO_COPY=1 timeout 600 python bench.py --workload detection_gap6_4Mx50 --steps 10 --e2e-steps 5 --no-cpu-baseline > gpurun_out/r2/bench_gap6_4M_nz.json 2>&1; python -c "import json;d=json.load(open('gpurun_out/r2/bench_gap6_4M_nz.json'));print('4M copy',d['e2e'])"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2/bench_reference.json 2> gpurun_out/r2/bench_reference.err; cat gpurun_out/r2/bench_reference.json | cut -c1-900
