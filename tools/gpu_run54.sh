cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2d
( time timeout 125 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2d/pytest_gpu12.log 2>&1
tail -4 gpurun_out/r2d/pytest_gpu12.log
timeout 40 python __graft_entry__.py smoke 2>&1 | tail -1
( time timeout 45 python bench.py --workload detection_abstar_c_1Mx100 --steps 20 --warmup 5 ) > gpurun_out/r2d/bench_abstar_c.json 2> gpurun_out/r2d/bench_abstar_c.err || tail -5 gpurun_out/r2d/bench_abstar_c.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2d/bench_abstar_c.json').read().strip().splitlines()[0])
    print('ms',d['ms_per_step'],'value','%.4g'%d['value'],'k1',d['roofline']['kernel_ms'],'frac',d['roofline']['frac'], 'parity', d['cpu_baseline']['parity_on_sample'], 'cpu %.4g'%d['cpu_baseline']['value'])
except Exception as e:
    print('no bench line', e)
PY
