cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 600 python bench.py --workload detection_gap6_4Mx50 --steps 10 --e2e-steps 5 --no-cpu-baseline > gpurun_out/r2/bench_gap6_4M_z2.json 2> gpurun_out/r2/bench_gap6_4M_z2.err; python -c "import json;d=json.load(open('gpurun_out/r2/bench_gap6_4M_z2.json'));print('16M chunks',d['e2e']['ms_per_step'],d['e2e']['value'])"
SIESTA_CHUNK_EVENTS=33554432 timeout 600 python bench.py --workload detection_gap6_4Mx50 --steps 10 --e2e-steps 5 --no-cpu-baseline > gpurun_out/r2/bench_gap6_4M_z3.json 2>&1; python -c "import json;d=json.load(open('gpurun_out/r2/bench_gap6_4M_z3.json'));print('32M chunks',d['e2e']['ms_per_step'],d['e2e']['value'])"
SIESTA_CHUNK_EVENTS=8388608 timeout 600 python bench.py --workload detection_gap6_4Mx50 --steps 10 --e2e-steps 5 --no-cpu-baseline > gpurun_out/r2/bench_gap6_4M_z4.json 2>&1; python -c "import json;d=json.load(open('gpurun_out/r2/bench_gap6_4M_z4.json'));print('8M chunks',d['e2e']['ms_per_step'],d['e2e']['value'])"
timeout 300 python tools/bench_counting.py > gpurun_out/r2/bench_counting.jsonl 2> gpurun_out/r2/bench_counting.err; tail -5 gpurun_out/r2/bench_counting.jsonl | cut -c1-400
