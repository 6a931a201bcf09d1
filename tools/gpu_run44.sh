cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2c
timeout 200 python -m pytest tests/test_golden.py tests/test_wnm.py -m gpu -x -q 2>&1 | tail -3
show() { python - <<PY
import json
d=json.loads(open('$1').read().strip().splitlines()[0])
print('$1', 'ms',round(d['ms_per_step'],3),'value','%.4g'%d['value'],'k1',round(d['roofline']['kernel_ms'],3),'all',round(d['roofline']['all_kernels_ms'],3),'frac',round(d['roofline']['frac'],3), 'parity', (d.get('cpu_baseline') or {}).get('parity_on_sample'), 'e2e', (d.get('e2e') or {}).get('ms_per_step'))
PY
}
timeout 300 python bench.py --workload detection_gap6_all_4Mx50 --steps 20 > gpurun_out/r2c/bench_gap6_all.json 2> gpurun_out/r2c/bench_gap6_all.err || tail -20 gpurun_out/r2c/bench_gap6_all.err
show gpurun_out/r2c/bench_gap6_all.json
