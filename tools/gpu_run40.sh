cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2c
( time timeout 420 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2c/pytest_gpu5.log 2>&1
tail -6 gpurun_out/r2c/pytest_gpu5.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 300 python tools/bench_wnm.py --traces 1000000 > gpurun_out/r2c/bench_wnm2.json 2> gpurun_out/r2c/bench_wnm2.err || tail -20 gpurun_out/r2c/bench_wnm2.err
cat gpurun_out/r2c/bench_wnm2.json
( time timeout 300 python bench.py --no-e2e ) > gpurun_out/r2c/bench_nostage.json 2> gpurun_out/r2c/bench_nostage.err || tail -20 gpurun_out/r2c/bench_nostage.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2c/bench_nostage.json').read().strip().splitlines()[0])
print('ms',d['ms_per_step'],'value',d['value'],'k1',d['roofline']['kernel_ms'],'all',d['roofline']['all_kernels_ms'],'frac',d['roofline']['frac'], 'parity', d['cpu_baseline']['parity_on_sample'])
PY
