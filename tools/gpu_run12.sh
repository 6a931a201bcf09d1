cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
V=sequencedetectionqueryexecutor_b200/variants
for v in main c6 c10 c12 main; do
  L=$PWD/$V/libsiesta_$v.so; [ $v = main ] && L=$PWD/sequencedetectionqueryexecutor_b200/libsiesta_gpu.so
  SIESTA_GPU_LIB=$L timeout 300 python bench.py --workload detection_gap6_4Mx50 --steps 20 --no-cpu-baseline --no-e2e > gpurun_out/r2/var2_$v.json 2>gpurun_out/r2/var2_$v.err
  python -c "import json;d=json.load(open('gpurun_out/r2/var2_$v.json'));print('$v',d['roofline']['kernel_ms'],d['roofline']['frac'],d['ms_per_step'])"
done
