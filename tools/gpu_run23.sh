cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_detect_gpu.py tests/test_exchange_gpu.py -m gpu -x -q 2>&1 | tail -3
for v in 0 1; do
  if [ $v = 1 ]; then export SIESTA_NKP_NO_L2_PREFETCH=1; fi
  timeout 300 python bench.py --workload detection_gap6_4Mx50 --steps 20 --e2e-steps 0 --no-cpu-baseline > gpurun_out/r2/pf_4M_$v.json 2> gpurun_out/r2/pf_4M_$v.err
  timeout 600 python bench.py --steps 10 --e2e-steps 0 --no-cpu-baseline > gpurun_out/r2/pf_100M_$v.json 2> gpurun_out/r2/pf_100M_$v.err
  python - <<PY
import json
for w in ('4M','100M'):
    d=json.load(open('gpurun_out/r2/pf_%s_$v.json'%w)); print('nopf=$v',w,d['ms_per_step'],d['roofline'])
PY
done
