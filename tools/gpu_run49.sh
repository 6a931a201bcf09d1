cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2c
timeout 200 python -m pytest tests/test_counting_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/bench_counting.py > gpurun_out/r2c/counting2.jsonl 2> gpurun_out/r2c/counting2.err || tail -5 gpurun_out/r2c/counting2.err
python - <<PY
import json
for l in open('gpurun_out/r2c/counting2.jsonl'):
    l=l.strip()
    if not l.startswith('{'): continue
    d=json.loads(l)
    if d['kernel'].startswith('K3'): print(d['workload'], d['kernel'][:60], 'ms', round(d['kernel_ms'],2), 'parity', d['parity_on_sample'])
PY
show() { python - <<PY
import json
d=json.loads(open('$1').read().strip().splitlines()[0])
print('$1', 'ms',round(d['ms_per_step'],3),'k1',round(d['roofline']['kernel_ms'],3),'all',round(d['roofline']['all_kernels_ms'],3))
PY
}
B="python bench.py --no-e2e --no-cpu-baseline --steps 10"
SIESTA_NKP_TILE_BATCH=8 $B > gpurun_out/r2c/tb8.json 2>/dev/null; show gpurun_out/r2c/tb8.json
SIESTA_NKP_TILE_BATCH=2 $B > gpurun_out/r2c/tb2.json 2>/dev/null; show gpurun_out/r2c/tb2.json
SIESTA_NKP_TILE_BATCH=16 $B > gpurun_out/r2c/tb16.json 2>/dev/null; show gpurun_out/r2c/tb16.json
