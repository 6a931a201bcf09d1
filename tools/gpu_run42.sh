cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2c
( time timeout 420 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2c/pytest_gpu6.log 2>&1
tail -5 gpurun_out/r2c/pytest_gpu6.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
show() { python - <<PY
import json
d=json.loads(open('$1').read().strip().splitlines()[0])
print('$1', 'ms',round(d['ms_per_step'],3),'value','%.4g'%d['value'],'k1',round(d['roofline']['kernel_ms'],3),'all',round(d['roofline']['all_kernels_ms'],3),'frac',round(d['roofline']['frac'],3), 'parity', (d.get('cpu_baseline') or {}).get('parity_on_sample'))
PY
}
B4="python bench.py --workload detection_gap6_4Mx50 --no-e2e --no-cpu-baseline --steps 20"
$B4 > gpurun_out/r2c/g4_uniform.json 2>/dev/null; show gpurun_out/r2c/g4_uniform.json
SIESTA_NO_UNIFORM_GATHER=1 $B4 > gpurun_out/r2c/g4_general.json 2>/dev/null; show gpurun_out/r2c/g4_general.json
timeout 300 python bench.py --no-e2e > gpurun_out/r2c/bench_ugather.json 2> gpurun_out/r2c/bench_ugather.err || tail -20 gpurun_out/r2c/bench_ugather.err
show gpurun_out/r2c/bench_ugather.json
