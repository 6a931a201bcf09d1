cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
V=sequencedetectionqueryexecutor_b200/variants
for v in A B C D E A; do
  SIESTA_GPU_LIB=$PWD/$V/libsiesta_$v.so timeout 300 python bench.py --workload detection_gap6_4Mx50 --steps 20 --no-cpu-baseline --no-e2e > gpurun_out/r2/var_$v.json 2>gpurun_out/r2/var_$v.err
  python -c "import json;d=json.load(open('gpurun_out/r2/var_$v.json'));print('$v',d['roofline']['kernel_ms'],d['roofline']['frac'],d['ms_per_step'])"
done
for c in 25 50 100; do
  SIESTA_NKP_CARVEOUT=$c SIESTA_GPU_LIB=$PWD/$V/libsiesta_A.so timeout 300 python bench.py --workload detection_gap6_4Mx50 --steps 20 --no-cpu-baseline --no-e2e > gpurun_out/r2/var_A_c$c.json 2>&1
  python -c "import json;d=json.load(open('gpurun_out/r2/var_A_c$c.json'));print('A carveout $c',d['roofline']['kernel_ms'],d['roofline']['frac'],d['ms_per_step'])"
done
timeout 300 python -m pytest tests/test_exchange_gpu.py -x -q 2>&1 | tail -3
