cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_counting_gpu.py tests/test_declare_templates.py tests/test_exchange_gpu.py -m gpu -x -q 2>&1 | tail -5
timeout 600 python tools/bench_counting.py > gpurun_out/r2/bench_counting2.jsonl 2> gpurun_out/r2/bench_counting2.err; grep -h "K3" gpurun_out/r2/bench_counting2.jsonl | cut -c1-420; tail -3 gpurun_out/r2/bench_counting2.err
