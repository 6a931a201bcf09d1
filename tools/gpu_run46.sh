cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2c
( time timeout 420 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2c/pytest_gpu9.log 2>&1
tail -5 gpurun_out/r2c/pytest_gpu9.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
( time timeout 400 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2c/bench_default_full.json 2> gpurun_out/r2c/bench_default_full.err || tail -20 gpurun_out/r2c/bench_default_full.err
tail -4 gpurun_out/r2c/bench_default_full.err
( time timeout 400 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/r2c/bench_reference_full.json 2> gpurun_out/r2c/bench_reference_full.err || tail -20 gpurun_out/r2c/bench_reference_full.err
tail -4 gpurun_out/r2c/bench_reference_full.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2c/bench_default_full.json').read().strip().splitlines()[0])
r=json.loads(open('gpurun_out/r2c/bench_reference_full.json').read().strip().splitlines()[0])
print('ms',d['ms_per_step'],'value','%.4g'%d['value'],'k1',d['roofline']['kernel_ms'],'frac',d['roofline']['frac'], 'e2e %.4g'%d['e2e']['value'], d['e2e']['ms_per_step'], 'parity', d['cpu_baseline']['parity_on_sample'], 'launches', d['gpu_launches'], 'clocks', d['clocks'])
print('reference %.4g'%r['value'], 'ratio e2e', d['e2e']['value']/r['value'], 'ratio value', d['value']/r['value'])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c/launches_default.csv python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2c/ncu_default.log 2>&1
tail -2 gpurun_out/r2c/ncu_default.log | cut -c1-300
wc -l gpurun_out/r2c/launches_default.csv
