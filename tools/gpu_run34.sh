cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2b
N=${1:-8}
run() { # name, blocks, extra args, env...
  name=$1; B=$2; extra=$3; shift 3
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e $extra --blocks $B > gpurun_out/r2b/y${N}_$name.json 2> gpurun_out/r2b/y${N}_$name.err || tail -20 gpurun_out/r2b/y${N}_$name.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r2b/y${N}_$name.json')); x=d['exchange']; print('$name N=$N blocks=$B ms',round(d['ms_per_step'],3),'p50',round(d['p50_latency_ms'],3),'k1',round(d['roofline']['kernel_ms'],3),'scan+place',round(x['scan_and_place_ms'],3),'wait',round(x['wait_for_slowest_rank_ms'],3),'pull',round(x['pull_and_decode_ms'],3),'join',round(x['join_stream_ms'],3),'eager',x['eager_allocation'],'parity',(d.get('cpu_baseline') or {}).get('parity_on_sample'))
PY
}
run b1 1 "" A=1
run b1_c149 1 --no-cpu-baseline SIESTA_XCHG_DECODE_CTAS=149
run b3 3 "" A=1
run b5 5 --no-cpu-baseline A=1
