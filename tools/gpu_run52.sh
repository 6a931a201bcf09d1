cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2c
timeout 300 python -m pytest tests/test_detect_gpu.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -3
run() { name=$1; shift; env "$@" > gpurun_out/r2c/e2e_$name.json 2> gpurun_out/r2c/e2e_$name.err || tail -5 gpurun_out/r2c/e2e_$name.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2c/e2e_$name.json').read().strip().splitlines()[0]); e=d['e2e']
print('$name', 'e2e ms', round(e['ms_per_step'],2), 'int32 ms', round(e.get('int32_column',{}).get('ms_per_step',0),2))
PY
}
B="python bench.py --workload detection_gap6_4Mx50 --no-cpu-baseline --no-properties --steps 5 --e2e-steps 5"
run s_on A=1 $B
run s_off SIESTA_NO_STREAMED_RESULT=1 $B
run s_on_32M SIESTA_CHUNK_EVENTS=33554432 $B
