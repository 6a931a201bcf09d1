cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2c
( time timeout 420 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2c/pytest_gpu2.log 2>&1
tail -8 gpurun_out/r2c/pytest_gpu2.log
run() { name=$1; shift; env "$@" > gpurun_out/r2c/e2e_$name.json 2> gpurun_out/r2c/e2e_$name.err || tail -5 gpurun_out/r2c/e2e_$name.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2c/e2e_$name.json').read().strip().splitlines()[0]); e=d['e2e']
print('$name', 'e2e ms', round(e['ms_per_step'],2), 'h2d', e['h2d_bytes_per_step'], 'd2h', e['d2h_bytes_per_step'], 'int32 ms', round(e.get('int32_column',{}).get('ms_per_step',0),2), 'step ms', round(d['ms_per_step'],3))
PY
}
B="python bench.py --workload detection_gap6_4Mx50 --no-cpu-baseline --steps 5 --e2e-steps 5"
run base A=1 $B
run nocols A=1 $B --e2e-extra-flags 16
run chunk4M SIESTA_CHUNK_EVENTS=4194304 $B
run chunk64M SIESTA_CHUNK_EVENTS=67108864 $B
run nocols_chunk64M SIESTA_CHUNK_EVENTS=67108864 $B --e2e-extra-flags 16
