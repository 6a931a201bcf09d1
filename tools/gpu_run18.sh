cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 600 python bench.py --no-e2e > gpurun_out/r2/bench_default_u.json 2> gpurun_out/r2/bench_default_u.err; python -c "import json;d=json.load(open('gpurun_out/r2/bench_default_u.json'));print('100M',d['roofline']['kernel_ms'],d['roofline']['frac'],d['ms_per_step'],d['value'],d['cpu_baseline']['parity_on_sample'],d['latency_ms'])"
timeout 300 python bench.py --workload detection_kleene_1Mx100 --steps 20 > gpurun_out/r2/bench_kleene_u.json 2> gpurun_out/r2/bench_kleene_u.err; python -c "import json;d=json.load(open('gpurun_out/r2/bench_kleene_u.json'));print('kleene',d['roofline']['kernel_ms'],d['roofline']['frac'],d['ms_per_step'],d['value'],d['e2e']['value'],d['cpu_baseline']['parity_on_sample'])"
