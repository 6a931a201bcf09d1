cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2c
( time timeout 420 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2c/pytest_gpu11.log 2>&1
tail -4 gpurun_out/r2c/pytest_gpu11.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
( time timeout 400 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2c/bench_default_final2.json 2> gpurun_out/r2c/bench_default_final2.err || tail -20 gpurun_out/r2c/bench_default_final2.err
tail -3 gpurun_out/r2c/bench_default_final2.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2c/bench_default_final2.json').read().strip().splitlines()[0])
print('ms',d['ms_per_step'],'value','%.4g'%d['value'],'k1',d['roofline']['kernel_ms'],'frac',d['roofline']['frac'], 'e2e %.4g'%d['e2e']['value'], d['e2e']['ms_per_step'], 'int32', d['e2e']['int32_column']['ms_per_step'], 'parity', d['cpu_baseline']['parity_on_sample'])
print(all(v for k,v in d['properties_full_size'].items()))
PY
