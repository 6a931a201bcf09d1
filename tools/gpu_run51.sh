cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2c
( time timeout 420 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2c/pytest_gpu10.log 2>&1
tail -4 gpurun_out/r2c/pytest_gpu10.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
( time timeout 400 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2c/bench_default_final.json 2> gpurun_out/r2c/bench_default_final.err || tail -20 gpurun_out/r2c/bench_default_final.err
tail -3 gpurun_out/r2c/bench_default_final.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2c/bench_default_final.json').read().strip().splitlines()[0])
print('ms',d['ms_per_step'],'value','%.4g'%d['value'],'k1',d['roofline']['kernel_ms'],'frac',d['roofline']['frac'], 'e2e %.4g'%d['e2e']['value'], d['e2e']['ms_per_step'], 'parity', d['cpu_baseline']['parity_on_sample'])
print(d.get('properties_full_size'))
PY
for w in detection_gap6_all_4Mx50 detection_kleene_1Mx100; do
timeout 200 python bench.py --workload $w --steps 10 --no-e2e > gpurun_out/r2c/props_$w.json 2> gpurun_out/r2c/props_$w.err || tail -8 gpurun_out/r2c/props_$w.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2c/props_$w.json').read().strip().splitlines()[0])
print('$w', round(d['ms_per_step'],3), d['cpu_baseline']['parity_on_sample'], d.get('properties_full_size'))
PY
done
