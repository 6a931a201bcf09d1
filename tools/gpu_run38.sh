cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2c
( time timeout 300 python -m pytest tests/test_detect_gpu.py tests/test_golden.py -m gpu -x -q ) > gpurun_out/r2c/pytest_gpu4.log 2>&1
tail -4 gpurun_out/r2c/pytest_gpu4.log
run() { name=$1; shift; env "$@" > gpurun_out/r2c/e2e_$name.json 2> gpurun_out/r2c/e2e_$name.err || tail -5 gpurun_out/r2c/e2e_$name.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2c/e2e_$name.json').read().strip().splitlines()[0]); e=d['e2e']
print('$name', 'e2e ms', round(e['ms_per_step'],2), 'int32 ms', round(e.get('int32_column',{}).get('ms_per_step',0),2), 'step ms', round(d['ms_per_step'],3))
PY
}
B="python bench.py --workload detection_gap6_4Mx50 --no-cpu-baseline --steps 5 --e2e-steps 5"
run p_default A=1 $B
run p_16M SIESTA_CHUNK_EVENTS=16777216 $B
run p_32M SIESTA_CHUNK_EVENTS=33554432 $B
run p_64M SIESTA_CHUNK_EVENTS=67108864 $B
run p_8M SIESTA_CHUNK_EVENTS=8388608 $B
