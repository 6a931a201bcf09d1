cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2c
( time timeout 420 python -m pytest tests -m gpu -x -q --durations=12 ) > gpurun_out/r2c/pytest_gpu.log 2>&1
tail -25 gpurun_out/r2c/pytest_gpu.log
( time timeout 300 python bench.py ) > gpurun_out/r2c/bench_default.json 2> gpurun_out/r2c/bench_default.err || tail -20 gpurun_out/r2c/bench_default.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2c/bench_default.json').read().strip().splitlines()[0])
print('ms',d['ms_per_step'],'value',d['value'],'k1',d['roofline']['kernel_ms'],'frac',d['roofline']['frac'])
print('e2e',json.dumps(d['e2e'],indent=1))
print('parity',d['cpu_baseline'])
PY
tail -5 gpurun_out/r2c/bench_default.err
