cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2b
( time timeout 1200 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -6
timeout 600 python bench.py > gpurun_out/r2b/bench_default.json 2> gpurun_out/r2b/bench_default.err || tail -5 gpurun_out/r2b/bench_default.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2b/bench_default.json')); print('default',d['ms_per_step'],d['value'],d['roofline']['kernel_ms'],d['roofline']['frac'],d['cpu_baseline'].get('parity_on_sample'),d['e2e']['value'])
PY
timeout 600 python bench.py --impl reference > gpurun_out/r2b/bench_ref.json 2> gpurun_out/r2b/bench_ref.err; cut -c1-400 gpurun_out/r2b/bench_ref.json
