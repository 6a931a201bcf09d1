cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2c
( time timeout 420 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2c/pytest_gpu8.log 2>&1
tail -5 gpurun_out/r2c/pytest_gpu8.log
show() { python - <<PY
import json
d=json.loads(open('$1').read().strip().splitlines()[0])
print('$1', 'ms',round(d['ms_per_step'],3),'value','%.4g'%d['value'],'k1',round(d['roofline']['kernel_ms'],3),'all',round(d['roofline']['all_kernels_ms'],3),'frac',round(d['roofline']['frac'],3), 'parity', (d.get('cpu_baseline') or {}).get('parity_on_sample'), 'e2e', (d.get('e2e') or {}).get('ms_per_step'), d['result'])
PY
}
timeout 300 python bench.py --workload detection_gap6_all_4Mx50 --steps 20 > gpurun_out/r2c/bench_gap6_all3.json 2> gpurun_out/r2c/bench_gap6_all3.err || tail -20 gpurun_out/r2c/bench_gap6_all3.err
show gpurun_out/r2c/bench_gap6_all3.json
timeout 300 python bench.py --workload detection_gap6_4Mx50 --steps 20 --no-e2e > gpurun_out/r2c/bench_gap6_4M.json 2>/dev/null
show gpurun_out/r2c/bench_gap6_4M.json
