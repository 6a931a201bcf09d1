cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
for ctas in 0 148 296 1184; do
export SIESTA_XCHG_DECODE_CTAS=$ctas; [ $ctas = 0 ] && unset SIESTA_XCHG_DECODE_CTAS
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2/bench_N2c.json 2> gpurun_out/r2/bench_N2c.err; python -c "
import json
d=json.load(open('gpurun_out/r2/bench_N2c.json')); print('ctas=$ctas', d['ms_per_step'], d['exchange']['ms'], d['exchange']['host_sizes_and_alloc_ms'], d['exchange']['pull_and_decode_ms'], d['exchange']['scan_and_place_ms'], d['roofline']['kernel_ms'])"
done
