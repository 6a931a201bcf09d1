cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2c
timeout 200 python -m pytest tests/test_counting_gpu.py tests/test_declare_templates.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/bench_counting.py > gpurun_out/r2c/counting.jsonl 2> gpurun_out/r2c/counting.err || tail -5 gpurun_out/r2c/counting.err
python - <<PY
import json
for l in open('gpurun_out/r2c/counting.jsonl'):
    l=l.strip()
    if not l.startswith('{'): continue
    d=json.loads(l)
    if d['kernel'].startswith('K3'): print(d['workload'], d['kernel'][:60], 'ms', round(d['kernel_ms'],2), 'parity', d['parity_on_sample'])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"detect_|gather_|scan_|set_tail|act_range|widen" -c 60 --csv --log-file gpurun_out/r2c/launches_default.csv python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2c/ncu_default.log 2>&1
wc -l gpurun_out/r2c/launches_default.csv
