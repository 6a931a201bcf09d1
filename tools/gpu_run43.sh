cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2c
( time timeout 300 python -m pytest tests/test_exchange_gpu.py -m gpu -x -q ) > gpurun_out/r2c/pytest_xchg.log 2>&1
tail -4 gpurun_out/r2c/pytest_xchg.log
N=2
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e > gpurun_out/r2c/bench_N2.json 2> gpurun_out/r2c/bench_N2.err || tail -20 gpurun_out/r2c/bench_N2.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2c/bench_N2.json').read().strip().splitlines()[0]); x=d['exchange']
print('N=2 ms',round(d['ms_per_step'],3),'value','%.4g'%d['value'],'k1',round(d['roofline']['kernel_ms'],3),'scan+place',round(x['scan_and_place_ms'],3),'wait',round(x['wait_for_slowest_rank_ms'],3),'pull',round(x['pull_and_decode_ms'],3),'parity',(d.get('cpu_baseline') or {}).get('parity_on_sample'))
PY
