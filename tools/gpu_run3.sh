set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_exchange_gpu.py -x -q > gpurun_out/r2/pytest3x.log 2>&1; echo "pytest exchange rc=$?"
tail -15 gpurun_out/r2/pytest3x.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest3.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2/pytest3.log
for i in 1 2; do timeout 300 python bench.py --workload detection_gap6_4Mx50 --steps 20 --no-cpu-baseline --no-e2e > gpurun_out/r2/bench_gap6_4M_q$i.json 2>gpurun_out/r2/bench_gap6_4M_q$i.err; python -c "import json;d=json.load(open('gpurun_out/r2/bench_gap6_4M_q$i.json'));print(d['roofline']['kernel_ms'],d['roofline']['frac'],d['ms_per_step'],d['roofline']['all_kernels_ms'])"; done
