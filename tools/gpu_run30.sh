cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2b
N=${1:-2}
run() { # name, blocks, env...
  name=$1; B=$2; shift 2
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --blocks $B > gpurun_out/r2b/x_$name.json 2> gpurun_out/r2b/x_$name.err || tail -20 gpurun_out/r2b/x_$name.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r2b/x_$name.json')); x=d['exchange']; print('$name N=$N blocks=$B ms',round(d['ms_per_step'],3),'k1',round(d['roofline']['kernel_ms'],3),'scan+place',round(x['scan_and_place_ms'],3),'tail',round(x['pull_and_decode_ms'],3),'join',round(x['join_stream_ms'],3),'eager',x['eager_allocation'])
PY
}
run deferred10 10 SIESTA_XCHG_EAGER_MAX_BYTES=1000
run deferred4 4 SIESTA_XCHG_EAGER_MAX_BYTES=1000
run eager4_c74 4 SIESTA_XCHG_DECODE_CTAS=74
run eager4_c37 4 SIESTA_XCHG_DECODE_CTAS=37
run eager10_c74 10 SIESTA_XCHG_DECODE_CTAS=74
