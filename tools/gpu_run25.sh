cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_detect_gpu.py tests/test_exchange_gpu.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -5
for w in detection_abc_kleene_gap_1Mx100; do
  timeout 600 python bench.py --workload $w --steps 10 --e2e-steps 3 > gpurun_out/r2/bench_$w.json 2> gpurun_out/r2/bench_$w.err || tail -5 gpurun_out/r2/bench_$w.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r2/bench_$w.json')); print('$w',d['ms_per_step'],d['value'],d['roofline']['kernel_ms'],d['roofline']['frac'],d['cpu_baseline'].get('parity_on_sample'),d['e2e']['value'])
PY
done
